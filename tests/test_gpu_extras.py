"""GPU tests of the extra batch results (b3w_batch_extras, BASELINE config 5's compact-result contract):
per-instance witness checksums computed by the expansion warps, full witnesses for a <= 1 024-instance sample out of the
streamed HBM ring, first violated row; single GPU and the multi-GPU entry point.  Checker: Oracle B."""
import hashlib
import os
import sys
import threading

import numpy as np
import pytest
import torch

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)
import make_golden_sums  # noqa: E402  (block_digests: the fixture's own definition)

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1
WS = 24093


def test_streamed_2_20_sums_and_samples(built):
    """config 5's shape on one GPU: nothing but compact results leaves the device, yet every witness is accounted for by
    its checksum and 1 024 of them are looked at byte by byte."""
    n = 1 << 20
    rows = gen.splitmix_compression_inputs(n, first=0)
    wc = pkg.builder("blake3_compression", device=0, chunk=4096, fused_check=True)
    rng = np.random.default_rng(11)
    sample = np.unique(np.concatenate([[0, 1, 4095, 4096, n - 1], rng.integers(0, n, 1019)]))[:1024]
    rng.shuffle(sample)                                         # any order
    res = wc.calculateWitnessBatch(rows, want_witness=False, sums=True, samples=sample, first_bad=True)
    assert res["witness"] is None and not res["status"].any()
    assert (res["first_bad"] == _lib.B3W_NO_ROW).all()
    # every one of the 2^20 sums against Oracle B: through the committed per-4096-block digests of Oracle B's sums
    # (tests/golden/make_golden_sums.py), and live on a random 2^15 of them (B3W_FULL_ORACLE=1: live on all, 150 s)
    g = np.load(os.path.join(GOLDEN, "compression_sums_2p20.npz"))
    assert int(g["log2_n"]) == 20 and np.array_equal(res["sums"][:16], g["first16"])
    got = make_golden_sums.block_digests(res["sums"])
    bad = np.nonzero(got != g["block_digest"])[0]
    assert bad.size == 0, "blocks of 4096 instances whose checksums differ from Oracle B's: %s" % bad[:8]
    assert hashlib.sha256(np.ascontiguousarray(res["sums"], "<u8").tobytes()).hexdigest() == bytes(g["sha256"]).decode()
    live = np.arange(n) if os.environ.get("B3W_FULL_ORACLE") == "1" else np.sort(rng.choice(n, 1 << 15, replace=False))
    want_sums = port.witness_batch("compression", rows[live], nthreads=NCPU, want="sums")
    assert np.array_equal(res["sums"][live], want_sums)
    want = port.witness_batch("compression", rows[sample], nthreads=NCPU)
    assert np.array_equal(res["samples"], want)
    t = wc.lastTiming()
    # kernel_ms sums the launches' device times; launches alternate between the two ring streams and may overlap, so the
    # sum is bounded by twice the call's wall clock, not by the wall clock itself
    assert t["instances"] == n and t["launches"] == n // 4096 and 0 < t["kernel_ms"] <= 2 * t["total_ms"]
    assert t["d2h_bytes"] == n * (1 + 64 + 8 + 4) + len(sample) * WS * 32
    wc.close()


def test_streamed_2_24_every_checksum_against_oracle_b(built):
    """BASELINE configs[4]'s size on one GPU (north_star: "bit-exact blake3_compression witnesses for 2^24 synthetic inputs"):
    12.9 TB of witnesses stream through the HBM ring with the fused check; the checksum of EVERY one of them is held to
    Oracle B's through the committed per-block digests (tests/golden/make_golden_sums.py 24: an hour of Oracle B, not
    repeatable inside the suite), and a random 2^14 of them are re-derived live."""
    n = 1 << 24
    rows = gen.parallel_rows(gen.splitmix_compression_inputs, n, first=0, threads=NCPU)
    wc = pkg.builder("blake3_compression", device=0, chunk=16384, fused_check=True)
    res = wc.calculateWitnessBatch(rows, want_witness=False, sums=True, first_bad=True)
    assert not res["status"].any() and (res["first_bad"] == _lib.B3W_NO_ROW).all()
    g = np.load(os.path.join(GOLDEN, "compression_sums_2p24.npz"))
    assert int(g["log2_n"]) == 24 and np.array_equal(res["sums"][:16], g["first16"])
    bad = np.nonzero(make_golden_sums.block_digests(res["sums"]) != g["block_digest"])[0]
    assert bad.size == 0, "blocks of 4096 instances whose checksums differ from Oracle B's: %s" % bad[:8]
    assert int(np.bitwise_xor.reduce(res["sums"])) == int(g["xor"])
    live = np.sort(np.random.default_rng(24).choice(n, 1 << 14, replace=False))
    assert np.array_equal(res["sums"][live], port.witness_batch("compression", rows[live], nthreads=NCPU, want="sums"))
    # out[16] of every instance is the plain BLAKE3 compression of its inputs?  checked at 2^16 in test_gpu_parity.py; here the
    # public outputs of the live sample against the oracle's witnesses
    w = port.witness_batch("compression", rows[live[:256]], nthreads=NCPU).view(np.uint32).reshape(256, WS, 8)
    assert np.array_equal(res["pub"][live[:256]], w[:, 1:17, 0])
    wc.close()


def test_streamed_nova_pasta_2_20_every_checksum_against_oracle_b(built):
    """BASELINE configs[3] (2^20 blake3_nova_pasta steps, Pallas scalar field, streamed): every checksum against Oracle B's
    committed digests (tests/golden/make_golden_sums.py 20 nova_pasta_o2), a random 2^13 re-derived live."""
    n = 1 << 20
    rows = gen.parallel_rows(gen.splitmix_nova_inputs, n, first=0, threads=NCPU)
    wc = pkg.builder("blake3_nova_pasta", device=0, chunk=16384)
    res = wc.calculateWitnessBatch(rows, want_witness=False, sums=True)
    assert not res["status"].any()
    g = np.load(os.path.join(GOLDEN, "nova_pasta_o2_sums_2p20.npz"))
    assert int(g["log2_n"]) == 20 and np.array_equal(res["sums"][:16], g["first16"])
    bad = np.nonzero(make_golden_sums.block_digests(res["sums"]) != g["block_digest"])[0]
    assert bad.size == 0, "blocks of 4096 instances whose checksums differ from Oracle B's: %s" % bad[:8]
    live = np.sort(np.random.default_rng(20).choice(n, 1 << 13, replace=False))
    want, want_sums, st = port.witness_batch("nova_pasta_o2", rows[live[:512]], nthreads=NCPU, want="both")
    assert not st.any() and np.array_equal(res["sums"][live[:512]], want_sums)
    assert np.array_equal(res["pub"][live[:512]], want.view(np.uint32).reshape(512, wc.witnessSize, 8)[:, 1:16, 0])       # z_{i+1}
    assert np.array_equal(res["sums"][live], port.witness_batch("nova_pasta_o2", rows[live], nthreads=NCPU, want="sums"))
    wc.close()


@pytest.mark.parametrize("name,variant,rows_fn", [("blake3_nova_pasta", "nova_pasta_o2", gen.splitmix_nova_inputs),
                                                  ("blake3_nova_o1", "nova_bn_o1", gen.splitmix_nova_inputs)])
def test_streamed_nova_sums_and_samples(built, name, variant, rows_fn):
    n = 20000
    rows = rows_fn(n, first=77)
    rows[5, 14] = rows[5, 12]                                   # asserts
    wc = pkg.builder(name, device=0, chunk=2048)
    sample = np.array([0, 6, 2047, 2048, 19999, 12345], np.uint64)
    res = wc.calculateWitnessBatch(rows, want_witness=False, sums=True, samples=sample)
    want, want_sums, st = port.witness_batch(variant, rows, nthreads=NCPU, want="both")
    ok = st == 0
    assert np.array_equal(res["status"] == 0, ok) and res["status"][5] == _lib.B3W_CIRCOM_ASSERT
    assert np.array_equal(res["sums"][ok], want_sums[ok]) and res["sums"][5] == 0
    assert np.array_equal(res["samples"], want[sample.astype(np.int64)])
    wc.close()


def test_argument_checks(built):
    wc = pkg.builder("blake3_compression", device=0)
    rows = gen.lcg_compression_inputs(4)
    with pytest.raises(pkg.B3WError) as e:
        wc.calculateWitnessBatch(rows, samples=[4])             # instance 4 of 4
    assert e.value.code == _lib.B3W_ERR_INVALID
    with pytest.raises(pkg.B3WError):
        wc.calculateWitnessBatch(np.repeat(rows, 300, 0), samples=np.arange(1025))   # > B3W_MAX_SAMPLES
    res = wc.calculateWitnessBatch(rows, first_bad=True)        # no check configured: "no violated row"
    assert (res["first_bad"] == _lib.B3W_NO_ROW).all()
    wc.close()


def test_first_bad_with_injected_fault(built):
    wc = pkg.builder("blake3_compression", device=0, fused_check=True, chunk=64)
    rows = gen.lcg_compression_inputs(200)
    wc.inject_fault(48 + 3, 1)
    res = wc.calculateWitnessBatch(rows, want_witness=False, first_bad=True, sums=True)
    assert (res["status"] == _lib.B3W_R1CS_VIOLATION).all() and (res["first_bad"] != _lib.B3W_NO_ROW).all()
    # the checksum follows the bytes that were stored (the faulty witness), not the clean one
    clean = port.witness_batch("compression", rows, nthreads=NCPU, want="sums")
    assert (res["sums"] != clean).all()
    wc.close()


def test_wide_message_words_keep_their_checksum(built):
    """the wide-domain kernel rewrites the m slots after the expansion: the fused checksum must follow"""
    wc = pkg.builder("blake3_compression", device=0)
    n = 64
    rows = gen.splitmix_compression_inputs(n, first=3)
    P = wc.prime
    vals = [[int(x) for x in r] for r in rows]
    for i in range(n):
        vals[i][8 + i % 16] = [2**32, P - 1, 2**33 + 5, P - 2**32][i % 4] if i % 3 else vals[i][8 + i % 16]
    res = wc.calculateWitnessBatchFr(vals, sums=True)
    ok = res["status"] == 0
    assert ok.sum() > n // 2
    from conftest import checksum_np
    assert np.array_equal(res["sums"][ok], checksum_np(res["witness"][ok], WS))
    for i in np.nonzero(ok)[0][:24]:
        rc, w = port.witness_fr("compression", vals[i])
        assert rc == 0 and np.array_equal(res["witness"][i], w)
    for i in np.nonzero(~ok)[0]:
        assert port.witness_fr("compression", vals[i])[0] == 4 and res["sums"][i] == 0
    wc.close()


def test_multi_gpu_extras(built):
    devices = None if torch.cuda.device_count() > 1 else [0, 0, 0]
    m = pkg.MultiGpuCalculator("blake3_compression", devices=devices, chunk=256, fused_check=True)
    n = 5001
    rows = gen.splitmix_compression_inputs(n, first=9)
    sample = np.array([5000, 0, 1666, 1667, 3333, 3334, 17], np.uint64)
    res = m.calculateWitnessBatch(rows, want_witness=False, sums=True, samples=sample, first_bad=True)
    want, want_sums, st = port.witness_batch("compression", rows, nthreads=NCPU, want="both")
    assert not res["status"].any() and np.array_equal(res["sums"], want_sums)
    assert np.array_equal(res["samples"], want[sample.astype(np.int64)])
    assert (res["first_bad"] == _lib.B3W_NO_ROW).all()
    m.close()


def test_one_context_serialises_concurrent_callers(built):
    """overlapping calls on one context (what Promise.all over one calculator does through the N-API worker
    threads) must not see each other's ring slots."""
    wc = pkg.builder("blake3_compression", device=0, chunk=32)
    batches = [gen.splitmix_compression_inputs(150, first=1000 * k) for k in range(6)]
    want = [port.witness_batch("compression", b, nthreads=4) for b in batches]
    got = [None] * len(batches)

    def work(k):
        got[k] = wc.calculateWitnessBatch(batches[k])["witness"]
    th = [threading.Thread(target=work, args=(k,)) for k in range(len(batches))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for k in range(len(batches)):
        assert np.array_equal(got[k], want[k]), k
    wc.close()


def test_more_device_launches_in_flight_than_counter_sets(built):
    """> 32 launches in flight on different streams must not corrupt each other's work-item counters."""
    wc = pkg.builder("blake3_compression", device=0)
    n, k = 96, 80
    rows = gen.splitmix_compression_inputs(n, first=5)
    want = port.witness_batch("compression", rows, nthreads=NCPU)
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    outs = [torch.zeros(n * WS * 32, dtype=torch.uint8, device="cuda") for _ in range(k)]
    streams = [torch.cuda.Stream() for _ in range(k)]
    torch.cuda.synchronize()
    for o, s in zip(outs, streams):
        wc.witness_batch_device(d_in.data_ptr(), n, o.data_ptr(), 0, 0, s.cuda_stream)
    torch.cuda.synchronize()
    w = torch.from_numpy(want.reshape(-1)).cuda()
    for j, o in enumerate(outs):
        assert torch.equal(o, w), j
    wc.close()


def test_entry_points_restore_the_callers_device(built):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    m = pkg.MultiGpuCalculator("blake3_compression", devices=None, chunk=64)
    from cuda.bindings import runtime as cudart
    assert cudart.cudaGetDevice()[1] == 0
    m.calculateWitnessBatch(gen.lcg_compression_inputs(64))
    assert cudart.cudaGetDevice()[1] == 0
    wc = pkg.builder("blake3_compression", device=1)
    wc.calculateWitnessBatch(gen.lcg_compression_inputs(2))
    assert cudart.cudaGetDevice()[1] == 0
    wc.close()
    m.close()


def test_byte_check_flag_checks_what_lies_in_the_ring(built):
    """B3W_FLAG_BYTE_CHECK: every chunk is read back from the HBM ring and all rows of the constraint system are evaluated on
    its bytes (the consumer's check, rust_fold/src/utils.rs:78-85) -- valid witnesses pass, "Assert Failed." stays."""
    n = 5000
    rows = gen.splitmix_nova_inputs(n, first=3)
    rows[5, 14] = rows[5, 12]                                   # asserts
    wc = pkg.builder("blake3_nova_pasta", device=0, chunk=1024, byte_check=True)
    res = wc.calculateWitnessBatch(rows, want_witness=False, sums=True, first_bad=True)
    _, want_sums, st = port.witness_batch("nova_pasta_o2", rows, nthreads=NCPU, want="both")
    ok = st == 0
    assert np.array_equal(res["status"] == 0, ok) and res["status"][5] == _lib.B3W_CIRCOM_ASSERT
    assert (res["first_bad"] == _lib.B3W_NO_ROW).all()
    assert np.array_equal(res["sums"][ok], want_sums[ok])
    assert wc.lastTiming()["launches"] == 2 * ((n + 1023) // 1024)            # generator + checker per chunk
    wc.close()


def test_byte_check_reports_the_rows_of_the_standalone_checker(built):
    wc = pkg.builder("blake3_compression", device=0, fused_check=True, byte_check=True, chunk=64)
    rows = gen.lcg_compression_inputs(200)
    wc.inject_fault(48 + 3, 1)                                  # the stored witnesses are faulty
    res = wc.calculateWitnessBatch(rows, want_witness=True, first_bad=True)
    assert (res["status"] == _lib.B3W_R1CS_VIOLATION).all()
    # the same bytes through b3w_r1cs_check_device: same verdict, same row numbering
    d = torch.from_numpy(np.ascontiguousarray(res["witness"])).cuda()
    d_st = torch.full((200,), 255, dtype=torch.uint8, device="cuda")
    d_bad = torch.zeros(200, dtype=torch.int32, device="cuda")
    wc.r1cs_check_device(d.data_ptr(), 200, d_st.data_ptr(), d_bad.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (d_st.cpu().numpy() == _lib.B3W_R1CS_VIOLATION).all()
    assert np.array_equal(d_bad.cpu().numpy().view(np.uint32), res["first_bad"])
    wc.close()


def test_ring_grows_with_the_batch(built):
    """the HBM ring is sized for the call at hand (min(chunk, n) rounded up to a power of two) and re-created when a larger
    batch arrives: same bytes whatever the order of the calls"""
    wc = pkg.builder("blake3_compression", device=0)            # default chunk
    rows = gen.lcg_compression_inputs(300)
    want = port.witness_batch("compression", rows, nthreads=NCPU)
    for n in (1, 70, 300, 5, 129):
        res = wc.calculateWitnessBatch(rows[:n], sums=True)
        assert not res["status"].any() and np.array_equal(res["witness"], want[:n]), n
    # the Fr256 staging follows the ring: field-element rows before and after a growth
    import ctypes as C
    L = pkg.lib()
    big = gen.lcg_compression_inputs(3000)
    want_big = port.witness_batch("compression", big, nthreads=NCPU, want="sums")
    for n in (200, 3000):
        fr = np.zeros((n, 28, 32), np.uint8)
        fr[:, :, :4] = big[:n].view(np.uint8).reshape(n, 28, 4)
        st, pub, sums = np.zeros(n, np.uint8), np.zeros((n, 16), np.uint32), np.zeros(n, np.uint64)
        ex = _lib.BatchExtras()
        ex.sums = sums.ctypes.data
        _lib.check(L.b3w_witness_batch_fr_ex(wc._h, fr.ctypes.data, n, None, st.ctypes.data, pub.ctypes.data, C.byref(ex)))
        assert not st.any() and np.array_equal(sums, want_big[:n]), n
    wc.close()
