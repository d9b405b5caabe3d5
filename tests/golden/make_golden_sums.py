#!/usr/bin/env python3
"""make_golden_sums.py [log2_n [variant]] -- regenerates tests/golden/compression_sums_2p20.npz (default) or, with 24, compression_sums_2p24.npz
(run anywhere: needs the oracles only; 2^20 takes 3.5 min on 8 cores, 2^24 an hour).

The fixture pins the per-instance witness checksums (the b3w_checksum_device definition) of the 2^20
blake3_compression instances splitmix_compression_inputs(2^20, first=0) -- BASELINE configs[3]'s count on the
compression circuit, SURVEY.md section 8(d) "a checksum of checksums" -- WITHOUT shipping 8 MiB of sums:

  block_digest[b] = first 8 bytes (little endian) of sha256(sums[4096 b : 4096 (b + 1)].tobytes()),  b = 0 .. 255
  sha256          = sha256 of all 2^20 sums (u64 little endian, instance order)

The sums come from Oracle B (oracle/circuit_oracle.c through oracle/port.py), the restatement that tests/test_oracle.py
pins to the reference wasm (Oracle A); this script re-checks a spread of 256 instances (one per block) against Oracle A
itself when oracle/_ref is built.  A GPU test that matches every block digest has matched every one of the 2^20 sums;
it does so without the 150 s (16 host cores) of oracle time the full comparison cost inside the GPU suite, which still
compares a random 2^15-instance subset with Oracle B live (B3W_FULL_ORACLE=1 brings the full live comparison back).
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import port, ref_wasm  # noqa: E402
from hot_proofs_blake3_circom_b200.inputs import splitmix_compression_inputs, splitmix_nova_inputs  # noqa: E402

LOG2_N, BLOCK = 20, 4096


def block_digests(sums):
    """u64[n] -> u64[n / BLOCK]: the first 8 bytes of each block's sha256 (shared with tests/test_gpu_extras.py)."""
    sums = np.ascontiguousarray(sums, "<u8")
    assert sums.size % BLOCK == 0
    return np.array([int.from_bytes(hashlib.sha256(sums[b:b + BLOCK].tobytes()).digest()[:8], "little")
                     for b in range(0, sums.size, BLOCK)], np.uint64)


def main(log2_n=LOG2_N, variant="compression"):
    n = 1 << log2_n
    rows_fn = splitmix_compression_inputs if variant == "compression" else splitmix_nova_inputs
    t = time.time()
    step = 1 << 18                                                       # bounded memory: 2^18 instances at a time
    sums = np.empty(n, np.uint64)
    for lo in range(0, n, step):
        rows = rows_fn(min(step, n - lo), first=lo)
        sums[lo:lo + rows.shape[0]], status = port.witness_batch(variant, rows, want="sums+status")
        assert not status.any(), "the fixture is defined on inputs no instance of which asserts (a GPU batch reports sums = 0 for those)"
        if log2_n > 20:
            print("  %d / %d  (%.0f s)" % (lo + rows.shape[0], n, time.time() - t), flush=True)
    print("oracle B: %d sums in %.0f s" % (n, time.time() - t))
    if ref_wasm.available(variant):
        idx = (np.arange(0, n, BLOCK) + (np.arange(n // BLOCK) * 2654435761 % BLOCK))[::max(1, n // BLOCK // 256)]   # <= 256 spread instances
        rows = np.concatenate([rows_fn(1, first=int(i)) for i in idx])
        wit, status, _ = ref_wasm.RefWasm(variant).batch_u32(rows, nthreads=os.cpu_count() or 1)
        assert (status == 0).all()
        wit_b, sums_b, _ = port.witness_batch(variant, rows, want="both")
        assert np.array_equal(wit, wit_b) and np.array_equal(sums_b, sums[idx])
        print("oracle A == oracle B on %d spread instances (every byte)" % idx.size)
    else:
        print("oracle/_ref not built: Oracle A cross-check skipped")
    sha = hashlib.sha256(np.ascontiguousarray(sums, "<u8").tobytes()).hexdigest()
    name = "%s_sums_2p%d.npz" % (variant, log2_n)
    np.savez_compressed(os.path.join(HERE, name), log2_n=np.uint32(log2_n), block=np.uint32(BLOCK),
                        block_digest=block_digests(sums), sha256=np.frombuffer(sha.encode(), np.uint8),
                        first16=sums[:16].copy(), xor=np.bitwise_xor.reduce(sums))
    print("%s: sha256 %s" % (name, sha))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else LOG2_N, sys.argv[2] if len(sys.argv) > 2 else "compression")
