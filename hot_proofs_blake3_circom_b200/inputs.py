"""inputs.py -- deterministic synthetic inputs for the batched witness generator.

lcg_compression_inputs : BASELINE config 2, distribution A -- the i-th successive `genRandomChunk(lcg)`
                         call on one shared `new LCG(6429)` (reference test/utils.ts:4-56,
                         test/witness_gen.test.ts:26,36); instance 0 is the reference's golden input.
splitmix_compression_inputs : distribution B -- counter-based (splitmix64 keyed by the instance index):
                         random h, t, d in 0..15, b = 4*(r mod 17) with the words beyond b/4 zeroed.
Rows are u32 in circuit declaration order h[8] m[16] t[2] b d.
"""
import numpy as np

IV = np.array([0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19],
              np.uint32)


def lcg_stream(seed, count):
    """`count` successive LCG.next() values (a=1664525, c=1013904223, m=2^32; test/utils.ts:10-20), vectorised
    by doubling: x_{k+s} = A_s x_k + C_s."""
    out = np.empty(count, np.uint64)
    if count == 0:
        return out.astype(np.uint32)
    a, c, m = 1664525, 1013904223, (1 << 32) - 1
    out[0] = (a * seed + c) & m
    filled, A, Cc = 1, a, c              # (A, Cc) = the map advancing by `filled` steps
    while filled < count:
        n = min(filled, count - filled)
        out[filled:filled + n] = (out[:n] * np.uint64(A) + np.uint64(Cc)) & np.uint64(m)
        Cc = (A * Cc + Cc) & m
        A = (A * A) & m
        filled += n
    return out.astype(np.uint32)


def lcg_compression_inputs(n, seed=6429, first=0):
    """Instances [first, first+n) of the genRandomChunk(lcg) sequence (b=64, d=0, t=[0,0], h=IV)."""
    draws = lcg_stream(seed, 17 * (first + n))[17 * first:].reshape(n, 17)
    rows = np.zeros((n, 28), np.uint32)
    rows[:, 0:8] = IV
    rows[:, 8:24] = draws[:, 1:]          # the first draw of every call is discarded (test/utils.ts:45)
    rows[:, 26] = 64
    return rows


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def splitmix_words(seed, index, nwords):
    """u32 words j = 0..nwords-1 for each instance index: low half of splitmix64(seed ^ (index << 8 | j))."""
    idx = np.asarray(index, np.uint64)[:, None]
    j = np.arange(nwords, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        return (splitmix64(np.uint64(seed) ^ ((idx << np.uint64(8)) | j)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def splitmix_compression_inputs(n, seed=0xB3B30001, first=0):
    w = splitmix_words(seed, np.arange(first, first + n, dtype=np.uint64), 30)
    rows = np.zeros((n, 28), np.uint32)
    rows[:, 0:8] = w[:, 0:8]
    b_words = (w[:, 28] % 17).astype(np.int64)            # b = 4 * (r mod 17)
    m = w[:, 8:24].copy()
    m[np.arange(16)[None, :] >= b_words[:, None]] = 0
    rows[:, 8:24] = m
    rows[:, 24:26] = w[:, 24:26]
    rows[:, 26] = (4 * b_words).astype(np.uint32)
    rows[:, 27] = w[:, 29] % 16
    return rows


def splitmix_nova_inputs(n, seed=0xB3B30004, first=0):
    """Independent blake3_nova step inputs: leaf_depth = total_depth in [1,64],
    depth in [0, leaf_depth), n_blocks in [1,16], block_count in [0, n_blocks), 64-bit chunk_idx, random u32 h / m,
    b in [0,64].  Rows are u32 in circuit declaration order (circuits/blake3_nova.circom:173-191):
    n_blocks block_count h[8] chunk_idx_low chunk_idx_high leaf_depth total_depth depth m[16] b."""
    w = splitmix_words(seed, np.arange(first, first + n, dtype=np.uint64), 40)
    rows = np.zeros((n, 32), np.uint32)
    n_blocks = w[:, 32] % 16 + 1
    leaf_depth = w[:, 34] % 64 + 1
    rows[:, 0] = n_blocks
    rows[:, 1] = w[:, 33] % n_blocks
    rows[:, 2:10] = w[:, 0:8]
    rows[:, 10:12] = w[:, 8:10]
    rows[:, 12] = leaf_depth
    rows[:, 13] = leaf_depth
    rows[:, 14] = w[:, 35] % leaf_depth
    rows[:, 15:31] = w[:, 10:26]
    rows[:, 31] = w[:, 36] % 65
    return rows


def parallel_rows(fn, n, first=0, threads=8, slice_len=1 << 18, **kw):
    """fn(count, first=...) over [first, first + n) in slices on a thread pool (numpy releases the GIL in the big array
    operations): the 2^24-instance inputs of BASELINE config 5 in seconds instead of half a minute."""
    from concurrent.futures import ThreadPoolExecutor
    starts = list(range(0, n, slice_len))
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        parts = list(ex.map(lambda s: fn(min(slice_len, n - s), first=first + s, **kw), starts))
    return np.concatenate(parts) if parts else fn(0, first=first, **kw)


def witness_checksums(wit, ws):
    """numpy statement of the per-instance witness checksum of include/blake3wit.h (b3w_checksum_device, b3w_batch_extras.sums):
    sum over slots s, 64-bit limbs j of (limb + 1) * (4 s + j + 1) * 0x9E3779B97F4A7C15 mod 2^64."""
    w = np.ascontiguousarray(wit).reshape(-1, ws * 32).view(np.uint64)
    e = np.arange(ws * 4, dtype=np.uint64)
    with np.errstate(over="ignore"):
        mix = (e + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        return ((w + np.uint64(1)) * mix[None, :]).sum(axis=1, dtype=np.uint64)
