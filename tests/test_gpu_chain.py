"""GPU tests of the chained-chunk driver b3w_nova_chain (BASELINE config 3): the batched form of the reference's
rust_fold prove_step loop.  Checker: oracle/nova_chain_ref.py (restatement of the Rust driver) + Oracle B / A for the
witness bytes + the blake3 package for the final hash (what rust_fold's tests assert, main.rs:392,410,439,474)."""
import os

import numpy as np
import pytest
import torch

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import nova_chain_ref, port, ref_wasm

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1
blake3 = pytest.importorskip("blake3")


def synth(n, seed=0xB3B30003):
    w = gen.splitmix_words(seed, np.arange((n + 3) // 4, dtype=np.uint64), 1)[:, 0]
    return w.tobytes()[:n]


@pytest.fixture(scope="module")
def wc(built):
    w = pkg.builder("blake3_nova", device=0)
    yield w
    w.close()


@pytest.mark.parametrize("n", [0, 4, 17, 64, 68, 1023, 1024, 1025, 2048, 3000, 4096, 5123, 7168, 8192, 16 * 1024 + 1])
def test_rows_match_the_rust_driver_restatement(wc, n):
    data = synth(n)
    res = wc.novaChain(data)
    rows, off, finals = nova_chain_ref.chain_rows(data)
    assert res["total_steps"] == len(rows) and list(res["step_off"]) == off
    assert np.array_equal(res["rows"], np.array(rows, np.uint32))
    assert (res["status"] == 0).all()
    assert res["root"] == finals[0]
    # h_out of every chunk's last step = that chunk's final value in the restatement
    for c in range(res["n_chunks"]):
        last = int(res["step_off"][c + 1]) - 1
        assert res["pub"][last, 2:10].tobytes() == finals[c]
    nchunks = max(1, (n + 1023) // 1024)
    if nchunks & (nchunks - 1) == 0:                      # perfect tree: every chunk folds to the real BLAKE3 hash
        assert all(f == blake3.blake3(data).digest() for f in finals)


@pytest.mark.parametrize("name,variant", [("blake3_nova", "nova_bn_o2"), ("blake3_nova_pasta", "nova_pasta_o2"),
                                          ("blake3_nova_o1", "nova_bn_o1")])
def test_witness_bytes_of_a_small_file(built, name, variant):
    w = pkg.builder(name, device=0, chunk=16)             # 16 steps per ring slot: exercises the ring
    data = synth(3000)
    res = w.novaChain(data, want_witness=True)
    want, _, st = port.witness_batch(variant, res["rows"], nthreads=NCPU, want="both")
    assert (st == 0).all()
    assert np.array_equal(res["witness"], want)
    z = want.view(np.uint32).reshape(res["total_steps"], w.witnessSize, 8)[:, 1:16, 0]
    assert np.array_equal(res["pub"], z)
    w.close()


def test_config3_one_mebibyte(wc):
    data = synth(1 << 20)
    res = wc.novaChain(data)
    assert res["n_chunks"] == 1024 and res["total_steps"] == 26624       # 1024 x (16 + 10), SURVEY 8(a) A13
    assert (res["status"] == 0).all()
    digest = blake3.blake3(data).digest()
    assert res["root"] == digest
    rows, pub, off = res["rows"], res["pub"], res["step_off"]
    last = (off[1:] - 1).astype(np.int64)
    assert all(pub[i, 2:10].tobytes() == digest for i in last)           # every chunk folds to blake3(file)
    # z chaining: outputs of step i are the public inputs of step i+1 inside a chunk
    inner = np.ones(res["total_steps"], bool)
    inner[last] = False
    i = np.nonzero(inner)[0]
    nxt = rows[i + 1]
    z = pub[i]
    assert np.array_equal(z[:, 0], nxt[:, 0]) and np.array_equal(z[:, 1], nxt[:, 1])          # n_blocks, block_count
    assert np.array_equal(z[:, 2:10], nxt[:, 2:10])                                           # h
    assert np.array_equal(z[:, 10], nxt[:, 13]) and np.array_equal(z[:, 11], nxt[:, 14])      # total_depth, depth
    assert np.array_equal(z[:, 12:14], nxt[:, 10:12]) and np.array_equal(z[:, 14], nxt[:, 12])  # chunk_idx, leaf_depth
    # witness bytes of a sample of chunks against the reference wasm (Oracle A) / the C oracle
    sel = np.concatenate([np.arange(off[c], off[c + 1]) for c in (0, 517, 1023)]).astype(np.int64)
    got = wc.calculateWitnessBatch(rows[sel])["witness"]
    assert np.array_equal(got, port.witness_batch("nova_bn_o2", rows[sel], nthreads=NCPU))
    if ref_wasm.available("nova_bn_o2"):
        ref = ref_wasm.RefWasm("nova_bn_o2")
        want, st, _ = ref.batch_u32(rows[sel[:26]], nthreads=min(NCPU, 26))
        assert (st == 0).all() and np.array_equal(got[:26], want)


@pytest.mark.parametrize("n", [0, 1000, 5123, 64 * 1024 + 77])
def test_multi_gpu_chain_equals_single(wc, n):
    """b3w_multi_nova_chain: chunks sharded over device slots (balanced by steps), every slot hashes the whole tree.
    With one GPU the same device is listed three times; on a multi-GPU box all visible devices are used too."""
    import torch
    data = synth(n)
    one = wc.novaChain(data, want_witness=n <= 5123)
    configs = [[0, 0, 0]] + ([None] if torch.cuda.device_count() > 1 else [])
    for devices in configs:
        m = pkg.MultiGpuCalculator("blake3_nova", devices=devices, chunk=64)
        res = m.novaChain(data, want_witness=n <= 5123)
        for k in ("rows", "status", "pub", "step_off"):
            assert np.array_equal(res[k], one[k]), k
        assert res["root"] == one["root"] and res["total_steps"] == one["total_steps"]
        if n <= 5123:
            assert np.array_equal(res["witness"], one["witness"])
        m.close()


def test_chain_called_twice_with_shrinking_input(wc):
    # the grow-only device scratch is reused: stale bytes of a longer earlier input must not leak into a shorter one
    big, small = synth(9000), synth(1500, seed=7)
    wc.novaChain(big)
    res = wc.novaChain(small)
    rows, off, finals = nova_chain_ref.chain_rows(small)
    assert np.array_equal(res["rows"], np.array(rows, np.uint32)) and res["root"] == finals[0]


# ---- non-power-of-two chunk counts: which chunks fold to blake3(file), under both sibling rules --------------------------
# The circuit orders (h, sibling) by a bit of the chunk index, which is the real direction only under a perfect subtree.
# Pinned so the behaviour cannot drift: for 3, 5, 6, 7, 9 chunks exactly the chunks listed here do NOT reach the file's hash,
# with the true BLAKE3 sibling (default) and with the reference's rule (rust_fold/src/blake3_hash.rs:60-78) alike.
NOT_FOLDING = {2: [], 3: [2], 4: [], 5: [4], 6: [4, 5], 7: [6], 8: [], 9: [8]}


@pytest.mark.parametrize("reference_siblings", [False, True], ids=["true-sibling", "reference-rule"])
def test_imperfect_trees_are_pinned(built, reference_siblings):
    w = pkg.builder("blake3_nova", device=0, reference_siblings=reference_siblings)
    differs = 0
    for nch, bad in NOT_FOLDING.items():
        data = synth(nch * 1024 - 5)
        res = w.novaChain(data)
        rows, off, finals = nova_chain_ref.chain_rows(data, reference_siblings)
        assert res["total_steps"] == len(rows) and list(res["step_off"]) == off
        assert np.array_equal(res["rows"], np.array(rows, np.uint32))          # the restatement of the chosen rule, row by row
        digest = blake3.blake3(data).digest()
        got_bad = [c for c in range(nch) if res["pub"][int(res["step_off"][c + 1]) - 1, 2:10].tobytes() != digest]
        assert got_bad == bad, (nch, got_bad)
        other, _, _ = nova_chain_ref.chain_rows(data, not reference_siblings)
        differs += other != rows
    assert differs >= 4            # the two rules really feed different siblings on the imperfect trees ...
    w.close()


def test_reference_rule_equals_default_on_perfect_trees(built):
    a = pkg.builder("blake3_nova", device=0)
    b = pkg.builder("blake3_nova", device=0, reference_siblings=True)
    for n in (1024, 2048, 4096, 8192, 16384):
        data = synth(n)
        ra, rb = a.novaChain(data), b.novaChain(data)
        assert np.array_equal(ra["rows"], rb["rows"]) and ra["root"] == rb["root"] == blake3.blake3(data).digest()
    a.close()
    b.close()


@pytest.mark.parametrize("name,variant", [("blake3_nova", "nova_bn_o2"), ("blake3_nova_pasta", "nova_pasta_o2")])
def test_chain_device_form(built, name, variant):
    """b3w_nova_chain_device: step witnesses left in device memory (one launch), what a prover on the same GPU consumes"""
    import torch
    w = pkg.builder(name, device=0)
    data = synth(5000)
    host = w.novaChain(data, want_witness=True)
    ns, ws = host["total_steps"], w.witnessSize
    d_out = torch.zeros(ns * ws * 32, dtype=torch.uint8, device="cuda")
    d_st = torch.full((ns,), 255, dtype=torch.uint8, device="cuda")
    d_pub = torch.zeros(ns * 15, dtype=torch.int32, device="cuda")
    d_rows = torch.zeros(ns * 32, dtype=torch.int32, device="cuda")
    res = w.novaChainDevice(data, d_out.data_ptr(), d_st.data_ptr(), d_pub.data_ptr(), d_rows.data_ptr())
    torch.cuda.synchronize()
    rows, off, finals = nova_chain_ref.chain_rows(data)
    assert res["total_steps"] == len(rows) and list(res["step_off"]) == off and res["root"] == finals[0]
    got_rows = d_rows.cpu().numpy().view(np.uint32).reshape(ns, 32)
    assert np.array_equal(got_rows, np.array(rows, np.uint32))
    want, _, st = port.witness_batch(variant, got_rows, nthreads=NCPU, want="both")
    assert (st == 0).all() and int(d_st.max()) == 0
    assert np.array_equal(d_out.cpu().numpy().reshape(ns, ws * 32), want)
    assert np.array_equal(d_pub.cpu().numpy().view(np.uint32).reshape(ns, 15), host["pub"])
    # without the optional buffers
    d_out.zero_()
    res2 = w.novaChainDevice(data, d_out.data_ptr())
    assert res2["root"] == res["root"] and np.array_equal(d_out.cpu().numpy().reshape(ns, ws * 32), want)
    t = w.lastTiming()
    assert t["launches"] == 1 and t["instances"] == ns and t["d2h_bytes"] == 0
    w.close()


@pytest.mark.parametrize("name", ["blake3_nova", "blake3_nova_o1"])
def test_chain_with_the_byte_check_flag(built, name):
    """B3W_FLAG_BYTE_CHECK covers the chain driver too: every step witness is re-read from the ring (or from the caller's
    device buffer) and all rows are evaluated on its bytes; the chain's results are unchanged, a faulty witness is flagged."""
    data = synth(40000)
    plain = pkg.builder(name, device=0, chunk=64)
    ref = plain.novaChain(data)
    plain.close()
    w = pkg.builder(name, device=0, chunk=64, byte_check=True)
    res = w.novaChain(data)
    assert (res["status"] == 0).all() and np.array_equal(res["pub"], ref["pub"]) and res["root"] == ref["root"]
    chunks = (res["total_steps"] + 63) // 64
    assert w.lastTiming()["launches"] == 2 * chunks                 # generator + checker per ring chunk
    ns = res["total_steps"]
    d_out = torch.zeros(ns * w.witnessSize * 32, dtype=torch.uint8, device="cuda")
    d_st = torch.full((ns,), 255, dtype=torch.uint8, device="cuda")
    w.novaChainDevice(data, d_out.data_ptr(), d_st.data_ptr())
    torch.cuda.synchronize()
    assert int(d_st.max()) == 0
    w.close()
    bad = pkg.builder(name, device=0, chunk=64, fused_check=True, byte_check=True)
    bad.inject_fault(48 + 3, 1)                                     # the stored witnesses are faulty
    res = bad.novaChain(data)
    assert (res["status"] == _lib.B3W_R1CS_VIOLATION).all()
    bad.close()
