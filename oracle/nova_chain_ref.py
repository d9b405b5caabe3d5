"""nova_chain_ref.py -- CPU restatement of the reference's Nova step driver.  TEST INFRASTRUCTURE ONLY.

Follows rust_fold/src/main.rs:71-94,130-142,166-171 (z0, number of steps, the prove_step loop),
rust_fold/src/blake3_circuit.rs:160-290 (new / update_for_step / format_input) and
rust_fold/src/blake3_hash.rs:17-93 (parent path = sibling chaining values, root first; the reference takes them from
`bao` 0.12.1 slice extraction -- not vendored -- whose published tree layout is BLAKE3's: the left subtree holds the
largest power of two of chunks that leaves at least one chunk on the right).  The per-step outputs z_{i+1} are
computed with the plain compression of oracle/blake3_ref.py using the flag / message rules of
circuits/blake3_nova.circom:86-167,229-266.
"""
from . import blake3_ref as b3

CHUNK_START, CHUNK_END, PARENT, ROOT = 1, 2, 4, 8


def _words(b):
    b = b + b"\0" * (64 - len(b))
    return [int.from_bytes(b[4 * i:4 * i + 4], "little") for i in range(16)]


def chunk_cv(chunk, idx):
    h = list(b3.IV)
    nb = max(1, (len(chunk) + 63) // 64)
    for k in range(nb):
        blk = chunk[64 * k:64 * k + 64]
        flags = (CHUNK_START if k == 0 else 0) | (CHUNK_END if k == nb - 1 else 0)
        h = b3.compress(h, _words(blk), idx & 0xFFFFFFFF, idx >> 32, len(blk), flags)[:8]
    return h


def tree_paths(cvs, reference_siblings=False):
    """-> per chunk: list of sibling CVs, root first (= Blake3HashProof.parent_path).
    Default: the TRUE sibling of every parent on the chunk's path (the child the path does not descend into).
    reference_siblings=True restates rust_fold/src/blake3_hash.rs:60-78 literally: the bao slice holds, root first, the
    64-byte parent nodes (left CV || right CV) on the path; for parent i of par_len the reference computes
    `mask = 1 << (par_len - i - 1)`, direction Left iff `leaf & mask == 0`, and keeps bytes 32..64 (the right child) when
    descending left, bytes 0..32 (the left child) otherwise -- whatever the path really does at that node."""
    paths = [[] for _ in cvs]

    def rec(first, n, stack):
        if n == 1:
            if reference_siblings:
                par_len = len(stack)
                paths[first] = [(rcv if first & (1 << (par_len - i - 1)) == 0 else lcv) for i, (lcv, rcv, _) in enumerate(stack)]
            else:
                paths[first] = [(rcv if went_left else lcv) for lcv, rcv, went_left in stack]
            return cvs[first]
        left = 1
        while left * 2 < n:
            left *= 2
        # hash children first (need both CVs before descending with the sibling known)
        lcv = subtree_cv(first, left)
        rcv = subtree_cv(first + left, n - left)
        rec(first, left, stack + [(lcv, rcv, True)])
        rec(first + left, n - left, stack + [(lcv, rcv, False)])
        return b3.compress(b3.IV, lcv + rcv, 0, 0, 64, PARENT)[:8]

    memo = {}

    def subtree_cv(first, n):
        if (first, n) in memo:
            return memo[(first, n)]
        if n == 1:
            r = cvs[first]
        else:
            left = 1
            while left * 2 < n:
                left *= 2
            r = b3.compress(b3.IV, subtree_cv(first, left) + subtree_cv(first + left, n - left), 0, 0, 64, PARENT)[:8]
        memo[(first, n)] = r
        return r

    rec(0, len(cvs), [])
    return paths


def chain_rows(data, reference_siblings=False):
    """-> (rows: list of 32-int step inputs in circuit declaration order, step_off, final h_out per chunk)."""
    chunks = [data[i:i + 1024] for i in range(0, len(data), 1024)] or [b""]
    cvs = [chunk_cv(c, i) for i, c in enumerate(chunks)]
    paths = tree_paths(cvs, reference_siblings)
    rows, step_off, finals = [], [0], []
    for c, chunk in enumerate(chunks):
        parent_path = paths[c]
        total_depth = leaf_depth = len(parent_path) + 1                 # blake3_circuit.rs:169, main.rs:71
        n_blocks = max(1, (len(chunk) + 63) // 64)                      # utils.rs:112-114 (0 bytes: one empty block)
        h, block_count, depth = list(b3.IV), 0, leaf_depth - 1          # z0 (main.rs:130-142)
        current_block, current_depth = 0, total_depth - 1
        for _ in range(n_blocks + total_depth - 1):                     # main.rs:94
            if current_block < n_blocks:                                # format_input :203-224
                blk = chunk[64 * current_block:64 * current_block + 64]
                m, b = _words(blk), len(blk)
            else:                                                       # :225-247
                m, b = parent_path[current_depth] + [0] * 8, 64
            rows.append([n_blocks, block_count] + h + [c & 0xFFFFFFFF, c >> 32, leaf_depth, total_depth, depth] + m + [b])
            # the circuit's outputs (circuits/blake3_nova.circom)
            is_parent, is_root = depth < leaf_depth - 1, depth == 0
            last = block_count == n_blocks - 1
            if is_parent:
                left = ((c >> (total_depth - 2 - depth)) & 1) == 0
                mm = h + m[:8] if left else m[:8] + h
                h = b3.compress(b3.IV, mm, 0, 0, b, PARENT | (ROOT if is_root else 0))[:8]
            else:
                flags = (CHUNK_START if block_count == 0 else 0) | (CHUNK_END if last else 0) | (ROOT if last and is_root else 0)
                h = b3.compress(h, m, c & 0xFFFFFFFF, c >> 32, b, flags)[:8]
                block_count += 1
            if (is_parent or last) and not is_root:
                depth -= 1
            # update_for_step (:185-195)
            if current_block < n_blocks:
                current_block += 1
            if current_block == n_blocks and current_depth > 0:
                current_depth -= 1
        step_off.append(len(rows))
        finals.append(b"".join(x.to_bytes(4, "little") for x in h))
    return rows, step_off, finals
