"""CPU tests: the restatement of the reference's Nova step driver (oracle/nova_chain_ref.py -- what tests/test_gpu_chain.py
holds b3w_nova_chain to) pinned against the reference's own Rust tests (rust_fold/src/main.rs:414-539: the final h_out of
every proved chunk == blake3::hash(data); the hashes those tests quote in comments) and against the circuit itself
(Oracle B evaluates every restated step row: outputs of step i are the inputs of step i + 1, the last ones are the hash)."""
import numpy as np
import pytest

from oracle import nova_chain_ref as ncr
from oracle import port

blake3 = pytest.importorskip("blake3")


def rng_bytes(n, seed):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8).tobytes()


# hashes quoted in the comments of the reference's tests (main.rs:512-514 "real ...", :521-523, :497/:507)
COMMENT_KATS = [(bytes(1024), "d6fd9de5bccf223f523b316c9cd1cf9a9d87ea42473d68e011dad13f09bf8917"),
                (bytes(68), "155e0c74d6aa369966999c8a972e3d92e6266656fd74087fa46531db452965f5"),
                (bytes(1028), "3c94b113d1a2f4e9b90058740c2843f45306e1dfdc3c69be25dd97cdfec89cab")]


@pytest.mark.parametrize("data,hexhash", COMMENT_KATS, ids=["1024-zeros", "68-zeros", "1028-zeros"])
def test_hashes_quoted_in_the_reference_tests(data, hexhash):
    assert blake3.blake3(data).hexdigest() == hexhash
    rows, step_off, finals = ncr.chain_rows(data)
    assert all(f.hex() == hexhash for f in finals)                       # every chunk's path folds to it
    # main.rs:512-514 also quotes the hash as little-endian words ("Hash bytes")
    if len(data) == 1024:
        words = ["%08x" % int.from_bytes(finals[0][4 * i:4 * i + 4], "little") for i in range(8)]
        assert words == ["e59dfdd6", "3f22cfbc", "6c313b52", "9acfd19c", "42ea879d", "e0683d47", "3fd1da11", "1789bf09"]


def rust_test_inputs():
    """the inputs of rust_fold/src/main.rs:414-539, StdRng replaced by numpy's generator (the byte values are free)"""
    cases = [("one_block", bytes(4)), ("one_block_nonempty", bytes([117]) * 17), ("two_blocks", bytes(68)),
             ("full_blocks", bytes(1024)), ("simple_path", bytes(1024 + 4)), ("middle_path", bytes(1024 * 3 + 5))]
    cases += [("random_chunk_%d" % n, rng_bytes(n, n)) for n in (1, 63, 64, 65, 777, 1023, 1024)]
    cases += [("full_bin_tree_%d" % c, rng_bytes(1024 * c, c)) for c in (2, 4, 8)]             # n_levels 2..4
    return cases


@pytest.mark.parametrize("name,data", rust_test_inputs(), ids=[c[0] for c in rust_test_inputs()])
def test_every_chunk_folds_to_blake3_of_the_file(name, data):
    rows, step_off, finals = ncr.chain_rows(data)
    want = blake3.blake3(data).digest()
    n_chunks = max(1, (len(data) + 1023) // 1024)
    assert len(finals) == n_chunks == len(step_off) - 1 and all(f == want for f in finals)
    # main.rs:94: n_blocks + total_depth - 1 steps per chunk
    for c in range(n_chunks):
        r0 = rows[step_off[c]]
        assert step_off[c + 1] - step_off[c] == r0[0] + r0[13] - 1
    # on perfect trees the reference's literal sibling rule (blake3_hash.rs:60-78) is the same path
    assert ncr.chain_rows(data, reference_siblings=True)[0] == rows


@pytest.mark.parametrize("variant", ["nova_pasta_o2", "nova_bn_o2", "nova_bn_o1"])
def test_restated_rows_through_the_circuit(built, variant):
    """Oracle B (the circuit, pinned to the reference wasm) on every restated step row of a 4-chunk file with a short last
    chunk: no assert fails, z_{i+1} (witness slots 1..15) is the next row's z_i, the last h_out is the file's hash."""
    data = rng_bytes(1024 * 3 + 5, 99)
    rows, step_off, finals = ncr.chain_rows(data)
    arr = np.array(rows, np.uint64).astype(np.uint32)
    wit, _, status = port.witness_batch(variant, arr, want="both")
    assert not status.any()
    ws = port.witness_size(variant)
    z_next = wit.view(np.uint32).reshape(len(rows), ws, 8)[:, 1:16, :]
    assert not z_next[:, :, 1:].any()                                    # the outputs are u32-valued
    # output order (circuits/blake3_nova.circom:192-202): n_blocks, block_count, h[8], total_depth, depth, chunk_idx lo / hi,
    # leaf_depth -- the inputs declare chunk_idx, leaf_depth, total_depth, depth (:173-184)
    z_next = z_next[:, :, 0][:, list(range(10)) + [12, 13, 14, 10, 11]]
    for c in range(len(finals)):
        lo, hi = step_off[c], step_off[c + 1]
        assert np.array_equal(z_next[lo:hi - 1], arr[lo + 1:hi, :15])    # the prove_step loop's hand-over (main.rs:166-171)
        assert z_next[hi - 1, 2:10].astype("<u4").tobytes() == blake3.blake3(data).digest() == finals[c]


def test_imperfect_trees_the_two_sibling_rules_differ():
    """3, 5, 6, 7 chunks: the literal reference rule and the true-sibling rule part ways; which chunks
    still fold to blake3(file) is pinned chunk by chunk in tests/test_gpu_chain.py -- here only that the restatement of
    both rules is deterministic and differs exactly where a right-spine subtree is shallower than the tree."""
    for n_chunks in (3, 5, 6, 7):
        data = rng_bytes(1024 * n_chunks, 7 * n_chunks)
        true_rows, _, true_finals = ncr.chain_rows(data)
        ref_rows, _, ref_finals = ncr.chain_rows(data, reference_siblings=True)
        want = blake3.blake3(data).digest()
        assert len(true_rows) == len(ref_rows)
        assert true_finals[0] == want and ref_finals[0] == want          # the left-most chunk descends a full-depth path
        assert ncr.chain_rows(data)[2] == true_finals
