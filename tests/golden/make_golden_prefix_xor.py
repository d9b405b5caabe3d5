#!/usr/bin/env python3
"""make_golden_prefix_xor.py -- tests/golden/compression_sums_prefix_xor.npz: the XOR of Oracle B's witness checksums over the
first 2^k instances of the splitmix blake3_compression sequence, k = 16 .. 23 (26 min of Oracle B on 8 cores); with argument
nova_pasta_o2: nova_pasta_o2_sums_prefix_xor.npz, k = 16 .. 19, for config 4's shards.

bench.py prints `sums_xor_rank0` for its streamed runs; rank 0's shard is always a prefix of the sequence (contiguous index
ranges), so these values are what the committed multi-GPU bench lines under profiles/ can be held to after the fact
(tests/test_bench_contract.py): 2^24 / N instances for config 5, 2^22 / N for the byte-check run.  Self-check: the block
digests of the sums computed here equal the first blocks of compression_sums_2p24.npz."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle import port  # noqa: E402
from hot_proofs_blake3_circom_b200.inputs import splitmix_compression_inputs, splitmix_nova_inputs  # noqa: E402
import make_golden_sums as mk  # noqa: E402

LOG2_MAX = 23


def main(variant="compression", log2_max=LOG2_MAX, digests="compression_sums_2p24.npz"):
    g = np.load(os.path.join(HERE, digests))
    rows_fn = splitmix_compression_inputs if variant == "compression" else splitmix_nova_inputs
    step, acc, t = 1 << 16, np.uint64(0), time.time()
    ks, xors = [], []
    for lo in range(0, 1 << log2_max, step):
        sums, status = port.witness_batch(variant, rows_fn(step, first=lo), want="sums+status")
        assert not status.any()
        assert np.array_equal(mk.block_digests(sums), g["block_digest"][lo // mk.BLOCK:(lo + step) // mk.BLOCK])
        acc ^= np.bitwise_xor.reduce(sums)
        done = lo + step
        if done & (done - 1) == 0:
            ks.append(done.bit_length() - 1)
            xors.append(acc)
            print("2^%d: %d  (%.0f s)" % (ks[-1], int(acc), time.time() - t), flush=True)
    np.savez_compressed(os.path.join(HERE, "%s_sums_prefix_xor.npz" % variant), log2_n=np.array(ks, np.uint32), xor=np.array(xors, np.uint64))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "nova_pasta_o2":          # config 4's shards: 2^20 / N steps per rank
        main("nova_pasta_o2", 19, "nova_pasta_o2_sums_2p20.npz")
    else:
        main()
