"""GPU tests of the full input domain of blake3_compression: message words outside u32 (valid witnesses in the reference,
SURVEY.md 8(a) A8) and inputs on which the reference asserts.  Checkers: the fixture made with the reference's own
witness program (tests/golden/compression_wide_cases.npz), Oracle B on fresh random inputs, Oracle A live on a few.
Bar: bit-exact witnesses, identical status, identical "Assert Failed." text."""
import os
import random

import numpy as np
import pytest
import torch

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import port, ref_wasm

pytestmark = pytest.mark.gpu
P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
WS = 24093
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def wc(built):
    assert torch.cuda.is_available(), "these tests need the B200"
    return pkg.builder("blake3_compression", device=0)


@pytest.fixture(scope="module")
def wide_cases():
    return np.load(os.path.join(GOLDEN, "compression_wide_cases.npz"))


def as_input(v):
    v = [int(x) for x in v]
    return {"h": v[0:8], "m": v[8:24], "t": v[24:26], "b": v[26], "d": v[27]}


def random_wide_rows(n, seed, p_assert=0.2):
    """mostly VALID wide instances: message words a little above 2^32 or a little below 0"""
    rnd = random.Random(seed)
    out = []
    for i in range(n):
        v = [rnd.randrange(2**32) for _ in range(28)]
        v[26], v[27] = rnd.randrange(65), rnd.randrange(16)
        kind = rnd.random()
        if kind < 0.15:
            pass                                              # a plain u32 instance inside the wide batch
        elif kind < 0.15 + p_assert:
            v[rnd.randrange(28)] = rnd.choice([2**34 + rnd.randrange(2**40), P - 2**33 - 1 - rnd.randrange(2**40), rnd.randrange(P)])
        else:
            for j in rnd.sample(range(16), rnd.randrange(1, 9)):
                v[8 + j] = rnd.choice([2**32 + rnd.randrange(2**30), P - 1 - rnd.randrange(2**30), 2**32, P - 1,
                                       rnd.randrange(2**32, 2**33 + 2**31)])
        out.append(v)
    return out


def oracle_b(vals_rows):
    status, wit = [], []
    for v in vals_rows:
        rc, w = port.witness_fr("compression", [x % P for x in v])
        status.append(rc)
        wit.append(w)
    return np.array(status), wit


def test_reference_fixture(wc, wide_cases):
    fr, status, valid, want = wide_cases["fr"], wide_cases["status"], wide_cases["valid"], wide_cases["witness"]
    res = wc.calculateWitnessBatchFr(fr)
    assert np.array_equal(res["status"], status.astype(np.uint8))
    assert np.array_equal(res["witness"][valid], want)
    assert (res["pub"][status == 4] == 0).all()
    # out[16] = witness slots 1..16
    assert np.array_equal(res["pub"][valid], want.view(np.uint32).reshape(len(valid), WS, 8)[:, 1:17, 0])


def test_surveyed_cases_through_the_single_witness_api(wc, golden, wide_cases):
    """m[0] = 2^32 and m[0] = p - 1 give valid witnesses, b = 2^33 asserts"""
    fr, valid, want = wide_cases["fr"], list(wide_cases["valid"]), wide_cases["witness"]
    row = [int(x) for x in golden["row"]]
    for i, x in ((0, 2**32), (1, -1)):                        # fixture cases 0 and 1; -1 is normalised to p - 1
        v = list(row)
        v[8] = x
        assert np.array_equal(fr[i, 8], np.frombuffer(int(x % P).to_bytes(32, "little"), np.uint8))
        w = wc.calculateWitness(as_input(v), 0)
        assert w[25] == x % P and w[0] == 1
        b = wc.calculateBinWitness(as_input(v), 0)
        assert np.array_equal(b, want[valid.index(i)])
    v = list(row)
    v[26] = 2**33
    with pytest.raises(RuntimeError) as e:
        wc.calculateWTNSBin(as_input(v), 0)
    assert str(e.value) == "Error: Assert Failed.\n" + bytes(wide_cases["text"][2]).decode()
    assert "RotXorWordBits_5 line: 62\nError in template HalfFunG_18 line: 91" in str(e.value)


def test_against_oracle_b_every_byte(wc):
    vals = random_wide_rows(384, 11)
    status, wit = oracle_b(vals)
    assert (status == 0).sum() > 150 and (status == 4).sum() > 40
    res = wc.calculateWitnessBatch([as_input(v) for v in vals])
    assert np.array_equal(res["status"], status.astype(np.uint8))
    for i in np.nonzero(status == 0)[0]:
        assert np.array_equal(res["witness"][i], wit[i]), i
    assert (res["pub"][status == 4] == 0).all()


@pytest.mark.skipif(not ref_wasm.available("compression"), reason="oracle/_ref not shipped")
def test_against_reference_wasm_live(wc):
    vals = random_wide_rows(24, 12, p_assert=0.1)
    ref = ref_wasm.RefWasm("compression")
    res = wc.calculateWitnessBatchFr(vals)
    n_ok = 0
    for i, v in enumerate(vals):
        rc, w = ref.calculate(as_input(v))
        assert res["status"][i] == rc, i
        if rc == 0:
            n_ok += 1
            assert np.array_equal(res["witness"][i], w), i
        else:
            assert wc.assertTraceFr(v) == ref.err_msg()
    assert n_ok >= 12


def device_run(wc, rows, ext, check):
    n = rows.shape[0]
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_ext = torch.from_numpy(ext).cuda()
    d_out = torch.zeros(n * WS * 32, dtype=torch.uint8, device="cuda")
    d_st = torch.full((n,), 255, dtype=torch.uint8, device="cuda")
    d_pub = torch.zeros((n, 16), dtype=torch.int32, device="cuda")
    d_bad = torch.zeros(n, dtype=torch.int32, device="cuda")
    wc.witness_batch_device_wide(d_in.data_ptr(), d_ext.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), d_pub.data_ptr(),
                                 d_bad.data_ptr() if check else 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_out, d_st.cpu().numpy(), d_bad.cpu().numpy().view(np.uint32)


def convert(vals):
    import ctypes as C
    n = len(vals)
    fr = np.frombuffer(b"".join(int(x % P).to_bytes(32, "little") for v in vals for x in v), np.uint8).copy()
    rows, ext = np.zeros((n, 28), np.uint32), np.zeros((n, 16), np.int8)
    _lib.check(pkg.lib().b3w_inputs_from_fr_wide(0, fr.ctypes.data, n, rows.ctypes.data, ext.ctypes.data, None))
    return rows, ext


def test_fused_check_and_hbm_check_on_wide_witnesses(wc):
    vals = random_wide_rows(256, 13)
    status, wit = oracle_b(vals)
    rows, ext = convert(vals)
    d_plain, st_plain, _ = device_run(wc, rows, ext, check=False)
    d_chk, st_chk, bad = device_run(wc, rows, ext, check=True)
    assert np.array_equal(st_plain, status.astype(np.uint8)) and np.array_equal(st_chk, st_plain)
    assert (bad == _lib.B3W_NO_ROW).all()
    ok = np.nonzero(status == 0)[0]
    got = d_chk.cpu().numpy().reshape(len(vals), WS * 32)
    assert np.array_equal(got[ok], d_plain.cpu().numpy().reshape(len(vals), WS * 32)[ok])
    assert np.array_equal(got[ok[0]], wit[ok[0]])
    # the stand-alone check re-reads the witnesses from HBM: negative message words are p - k there
    d_ok = torch.from_numpy(np.ascontiguousarray(got[ok])).cuda()
    d_st = torch.full((len(ok),), 255, dtype=torch.uint8, device="cuda")
    wc.r1cs_check_device(d_ok.data_ptr(), len(ok), d_st.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(d_st.max()) == 0
    # a fault in the trace is still caught with wide message words
    wc.inject_fault(48 + 3, 1)
    try:
        _, st_f, bad_f = device_run(wc, rows, ext, check=True)
    finally:
        wc.inject_fault()
    assert (st_f[ok] == _lib.B3W_R1CS_VIOLATION).all() and (bad_f[ok] != _lib.B3W_NO_ROW).all()
    assert (st_f[status == 4] == 4).all()


def test_any_work_item_split(wc):
    """the m slots may straddle work items: 32-slot items put m[0..6] and m[7..15] into different ones"""
    vals = random_wide_rows(64, 14, p_assert=0.0)
    status, wit = oracle_b(vals)
    rows, ext = convert(vals)
    for parts in (753, 5, 1):
        wc.set_launch(0, parts)
        try:
            d_out, st, _ = device_run(wc, rows, ext, check=False)
        finally:
            wc.set_launch(0, 0)
        got = d_out.cpu().numpy().reshape(len(vals), WS * 32)
        assert np.array_equal(st, status.astype(np.uint8))
        for i in np.nonzero(status == 0)[0]:
            assert np.array_equal(got[i], wit[i]), (parts, i)


def test_u32_rows_with_zero_ext_equal_the_plain_kernel(wc):
    rows = gen.splitmix_compression_inputs(512, first=3)
    ext = np.zeros((512, 16), np.int8)
    d_out, st, _ = device_run(wc, rows, ext, check=False)
    want = wc.calculateWitnessBatch(rows)["witness"]
    assert (st == 0).all() and np.array_equal(d_out.cpu().numpy().reshape(512, WS * 32), want)


def test_fused_flag_context_and_large_mixed_batch(built):
    """b3w_witness_batch_fr through the HBM ring (3 ring chunks) with B3W_FLAG_FUSED_CHECK: status only, checksummed"""
    wcf = pkg.builder("blake3_compression", device=0, chunk=512, fused_check=True)
    vals = random_wide_rows(1200, 15)
    status, _ = oracle_b(vals[:200])
    res = wcf.calculateWitnessBatchFr(vals, want_witness=False)
    assert np.array_equal(res["status"][:200], status.astype(np.uint8))
    assert set(np.unique(res["status"])) <= {0, 4}
    ok = np.nonzero(res["status"] == 0)[0]
    # out[16] of a wide instance = plain BLAKE3 compression of the low words (the carries do not reach the state)
    from oracle import blake3_ref
    rows, _ = convert(vals)
    for i in ok[:64]:
        r = [int(x) for x in rows[i]]
        assert list(res["pub"][i]) == blake3_ref.compress(r[0:8], r[8:24], r[24], r[25], r[26], r[27])
    wcf.close()
