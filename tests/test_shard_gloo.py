"""World-size-2 gloo test of the multi-GPU host logic (CPU): index-range sharding + the summary reduction.
The witnesses of each shard come from the C oracle here (no GPU in this container); on the GPU box the same plan
drives libblake3wit (bench.py, N > 1)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, n_total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from hot_proofs_blake3_circom_b200 import inputs as gen
    from hot_proofs_blake3_circom_b200.shard import shard_range, local_summary, reduce_summary
    from oracle import port
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port_no, rank=rank, world_size=world)
    first, count = shard_range(n_total, rank, world)
    rows = gen.lcg_compression_inputs(count, first=first)
    sums = port.witness_batch("compression", rows, nthreads=2, want="sums")
    total = reduce_summary(local_summary(np.zeros(count, np.uint8), sums), dist)
    q.put((rank, first, count, [int(x) for x in total]))
    dist.destroy_process_group()


def test_shard_ranges_partition_exactly():
    from hot_proofs_blake3_circom_b200.shard import shard_range
    for n in (0, 1, 7, 65536, 2 ** 24 + 3):
        for world in (1, 2, 4, 8):
            pos = 0
            for r in range(world):
                first, count = shard_range(n, r, world)
                assert first == pos and count in (n // world, n // world + 1)
                pos += count
            assert pos == n


def test_two_rank_gloo_sharded_batch_matches_single(built):
    from hot_proofs_blake3_circom_b200 import inputs as gen
    from hot_proofs_blake3_circom_b200.shard import local_summary
    from oracle import port
    n_total, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 19), (19, 18)]
    single = local_summary(np.zeros(n_total, np.uint8),
                           port.witness_batch("compression", gen.lcg_compression_inputs(n_total), nthreads=4, want="sums"))
    for r in res:
        assert r[3] == [int(x) for x in single]          # every rank holds the same reduced summary
