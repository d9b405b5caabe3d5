"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI of
libblake3wit.so (via the witness_calculator mirror); the oracles are only the checkers.
Bar: bit-exact (integer / byte work)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import blake3_ref, port, ref_wasm
from conftest import checksum_np

pytestmark = pytest.mark.gpu
WS = 24093
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NCPU = os.cpu_count() or 1


@pytest.fixture(scope="module")
def wc(built):
    assert torch.cuda.is_available(), "these tests need the B200"
    return pkg.builder("blake3_compression", device=0)


def as_input(row):
    row = [int(x) for x in row]
    return {"h": row[0:8], "m": row[8:24], "t": row[24:26], "b": row[26], "d": row[27]}


# ---- config 1: the reference's golden vector through the single-witness API -----------------------
def test_golden_wtns_bytes(wc, golden):
    buff = wc.calculateWTNSBin(as_input(golden["row"]), 0)
    assert buff.tobytes() == golden["wtns"].tobytes()


def test_calculate_witness_and_bin_witness(wc, golden):
    w = wc.calculateWitness(as_input(golden["row"]), 0)
    assert len(w) == WS and w[0] == 1
    assert w[1:17] == [int(x) for x in golden["public"]]          # public.json = main.out
    b = wc.calculateBinWitness(as_input(golden["row"]), 0)
    assert b.tobytes() == golden["wtns"].tobytes()[76:]


def test_cli_generate_witness(built, golden, tmp_path):
    inp = tmp_path / "testInp.json"
    out = tmp_path / "witness.wtns"
    inp.write_text(json.dumps(as_input(golden["row"])))
    subprocess.run([sys.executable, "-m", "hot_proofs_blake3_circom_b200.generate_witness", "blake3_compression",
                    str(inp), str(out)], check=True, cwd=ROOT)
    assert out.read_bytes() == golden["wtns"].tobytes()
    r = subprocess.run([sys.executable, "-m", "hot_proofs_blake3_circom_b200.generate_witness"], cwd=ROOT,
                       capture_output=True, text=True)
    assert r.stdout.startswith("Usage:")                              # generate_witness.js:4-5


# ---- fixtures made with the reference wasm (edge cases: empty / ragged blocks, extreme words) ------
def test_reference_cases_fixture(wc, cases):
    res = wc.calculateWitnessBatch(cases["rows"])
    assert (res["status"] == 0).all()
    assert np.array_equal(res["witness"], cases["witness"])


# ---- live against Oracle A (the reference's own witness program) ----------------------------------
@pytest.mark.skipif(not ref_wasm.available("compression"), reason="oracle/_ref not shipped")
def test_against_reference_wasm_live(wc):
    rows = np.concatenate([gen.lcg_compression_inputs(16, first=65520), gen.splitmix_compression_inputs(32, first=7)])
    ref = ref_wasm.RefWasm("compression")
    want, st, _ = ref.batch_u32(rows, nthreads=min(NCPU, 48))
    assert (st == 0).all()
    got = wc.calculateWitnessBatch(rows)["witness"]
    assert np.array_equal(got, want)


# ---- against Oracle B on every byte of a few thousand instances ------------------------------------
def test_against_port_4096_full_bytes(wc):
    rows = np.concatenate([gen.lcg_compression_inputs(2048), gen.splitmix_compression_inputs(2048)])
    want = port.witness_batch("compression", rows, nthreads=NCPU)
    res = wc.calculateWitnessBatch(rows)
    assert np.array_equal(res["witness"], want)
    outs = want.view(np.uint32).reshape(-1, WS, 8)[:, 1:17, 0]
    assert np.array_equal(res["pub"], outs)


# ---- ragged batch sizes, ring-slot boundaries, idempotence -----------------------------------------
@pytest.mark.parametrize("n", [0, 1, 7, 9, 33])
def test_ragged_batches_and_ring_boundaries(built, n):
    small = pkg.builder("blake3_compression", device=0, chunk=4)       # 4 instances per ring slot
    rows = gen.splitmix_compression_inputs(n, first=99)
    res = small.calculateWitnessBatch(rows)
    assert res["witness"].shape == (n, WS * 32)
    if n:
        assert np.array_equal(res["witness"], port.witness_batch("compression", rows, nthreads=NCPU))
        again = small.calculateWitnessBatch(rows)
        assert np.array_equal(again["witness"], res["witness"])
        one = small.calculateBinWitness(as_input(rows[n - 1]))
        assert np.array_equal(one, res["witness"][n - 1])
    small.close()


def test_compact_mode_no_witness_copy(wc):
    rows = gen.splitmix_compression_inputs(3000, first=5)
    res = wc.calculateWitnessBatch(rows, want_witness=False)
    assert res["witness"] is None and (res["status"] == 0).all()
    want = np.array([blake3_ref.compress(*_split(r)) for r in rows[:64]], np.uint32)
    assert np.array_equal(res["pub"][:64], want)


def _split(r):
    r = [int(x) for x in r]
    return r[0:8], r[8:24], r[24], r[25], r[26], r[27]


def test_device_entry_point_argument_checks(wc):
    d_in = torch.zeros(28, dtype=torch.int32, device="cuda")
    d_out = torch.zeros(WS * 32 + 64, dtype=torch.uint8, device="cuda")
    with pytest.raises(pkg.B3WError) as e:
        wc.witness_batch_device(d_in.data_ptr(), 1, d_out.data_ptr() + 8)
    assert e.value.code == _lib.B3W_ERR_INVALID and "aligned" in str(e.value)


# ---- BASELINE config 2 at full size: 2^16 instances resident in HBM --------------------------------
def test_full_2p16_checksums_and_properties(wc):
    n = 1 << 16
    rows = gen.lcg_compression_inputs(n)                               # instance 0 = the golden input
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_out = torch.empty(n * WS * 32, dtype=torch.uint8, device="cuda")
    d_st = torch.full((n,), 255, dtype=torch.uint8, device="cuda")
    d_pub = torch.empty(n * 16, dtype=torch.int32, device="cuda")
    d_sum = torch.empty(n, dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    wc.witness_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), d_pub.data_ptr(), s)
    wc.checksum_device(d_out.data_ptr(), n, d_sum.data_ptr(), s)       # re-reads what was written to HBM
    torch.cuda.synchronize()
    assert int(d_st.max()) == 0
    sums = d_sum.cpu().numpy().view(np.uint64)
    # every instance: checksum of all 770 976 bytes vs the C oracle
    want = port.witness_batch("compression", rows, nthreads=NCPU, want="sums")
    assert np.array_equal(sums, want)
    # a checksum of checksums, for the record
    with np.errstate(over="ignore"):
        total = np.bitwise_xor.reduce(sums)
    print("xor of 2^16 witness checksums: 0x%016x" % int(total))
    # size-independent properties read straight from HBM
    w = d_out.view(n, WS, 32)
    assert bool((w[:, 0, 0] == 1).all()) and int(w[:, 0, 1:].max()) == 0        # slot 0 == 1
    assert int(w[:, :, 8:].max()) == 0                                          # no compression slot exceeds 64 bits
    pub = d_pub.cpu().numpy().view(np.uint32).reshape(n, 16)
    idx = np.random.default_rng(1).integers(0, n, 256)
    for i in idx:
        assert list(pub[i]) == blake3_ref.compress(*_split(rows[i]))           # out == plain BLAKE3 compress
    # first / last / middle instances byte-for-byte
    sel = [0, 1, n // 2, n - 2, n - 1]
    got = w[sel].cpu().numpy().reshape(len(sel), WS * 32)
    assert np.array_equal(got, port.witness_batch("compression", rows[sel], nthreads=4))
    assert (checksum_np(got, WS) == sums[sel]).all()


def test_multi_gpu_entry_point_matches_single(built):
    """b3w_multi_witness_batch: contiguous shards, one host thread + context per device slot.  With one GPU in the box
    the same device is listed three times, which exercises the sharding, the threads and the slice arithmetic; on a
    multi-GPU box every visible device is used as well."""
    import torch
    rows = gen.splitmix_compression_inputs(301)
    want = port.witness_batch("compression", rows, nthreads=4)
    configs = [[0, 0, 0]]
    if torch.cuda.device_count() > 1:
        configs.append(None)
    for devices in configs:
        m = pkg.MultiGpuCalculator("blake3_compression", devices=devices, chunk=64)
        assert m.nDevices == (len(devices) if devices else torch.cuda.device_count())
        pos = 0
        for g in range(m.nDevices):
            first, count = m.shard(301, g)
            assert first == pos
            pos += count
        assert pos == 301
        res = m.calculateWitnessBatch(rows)
        assert not res["status"].any()
        assert np.array_equal(res["witness"], want)
        assert np.array_equal(res["pub"], want.view(np.uint32).reshape(301, WS, 8)[:, 1:17, 0])
        small = m.calculateWitnessBatch(rows[:2])            # fewer instances than device slots: empty shards
        assert np.array_equal(small["witness"], want[:2])
        m.close()


def test_field_element_inputs_entry_point(wc):
    """b3w_witness_batch_fr: Fr256 inputs (as rust_fold holds them), reduced mod p like normalize()."""
    rows = gen.splitmix_compression_inputs(9)
    P = wc.prime
    fr = np.zeros((9, 28, 32), np.uint8)
    for i in range(9):
        for k in range(28):
            v = int(rows[i, k]) + (P if (i + k) % 3 == 0 else 0)          # some values non-canonical (x + p)
            fr[i, k] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
    out = np.zeros((9, WS * 32), np.uint8)
    st = np.ones(9, np.uint8)
    _lib.check(pkg.lib().b3w_witness_batch_fr(wc._h, fr.ctypes.data, 9, out.ctypes.data, st.ctypes.data, None))
    assert not st.any()
    assert np.array_equal(out, port.witness_batch("compression", rows, nthreads=2))


# ---- compressible device memory (b3w_device_alloc): same bytes, hardware-compressed between L2 and HBM --------------
def test_compressible_output_buffer_and_ring(wc):
    n = 4096
    rows = gen.splitmix_compression_inputs(n, first=21)
    w_o, _, _ = port.witness_batch("compression", rows, nthreads=NCPU, want="both")        # the checker is Oracle B, not another GPU run
    want = {"witness": w_o, "pub": w_o.view(np.uint32).reshape(n, WS, 8)[:, 1:17, 0]}
    ptr, granted = wc.device_alloc(n * WS * 32, compressible=True)
    assert ptr and ptr % 32 == 0
    assert granted, "B200 grants CU_MEM_ALLOCATION_COMP_GENERIC"
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_st = torch.full((n,), 255, dtype=torch.uint8, device="cuda")
    d_sums = torch.zeros(n, dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    wc.witness_batch_device(d_in.data_ptr(), n, ptr, d_st.data_ptr(), 0, s)
    wc.checksum_device(ptr, n, d_sums.data_ptr(), s)
    torch.cuda.synchronize()
    assert int(d_st.max()) == 0
    assert np.array_equal(d_sums.cpu().numpy().view(np.uint64), checksum_np(want["witness"], WS))
    # bytes come back identical through an ordinary device-to-host copy
    from cuda.bindings import runtime as cudart
    host = np.empty(64 * WS * 32, np.uint8)
    err, = cudart.cudaMemcpy(host.ctypes.data, ptr + 1000 * WS * 32, host.nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    assert int(err) == 0
    assert np.array_equal(host.reshape(64, WS * 32), want["witness"][1000:1064])
    wc.device_free(ptr)
    with pytest.raises(pkg.B3WError):
        wc.device_free(ptr)                                            # not (any more) one of this context's blocks
    plain, g2 = wc.device_alloc(1 << 20, compressible=False)
    assert plain and not g2
    wc.device_free(plain)
    # the host-buffer calls with the ring in compressible memory
    wcr = pkg.builder("blake3_compression", device=0, chunk=1024, compressible_ring=True)
    got = wcr.calculateWitnessBatch(rows)
    assert np.array_equal(got["witness"], want["witness"]) and np.array_equal(got["pub"], want["pub"])
    wcr.close()
