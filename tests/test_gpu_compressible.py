"""GPU parity on the path the headline numbers are measured on: witnesses written into COMPRESSIBLE device memory
(b3w_device_alloc, the library's default HBM ring), where launch_witness splits a witness into 12 work items instead of
24.  All four circuits x {plain, fused check}; EVERY byte against Oracle B (the C restatement), a sample against Oracle A
(the reference's own wasm), never against another GPU run.  Bar: bit-exact."""
import os

import numpy as np
import pytest
import torch

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import port, ref_wasm
from conftest import checksum_np

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1
N = 4096
CASES = [("compression", "blake3_compression"), ("nova_bn_o2", "blake3_nova"), ("nova_pasta_o2", "blake3_nova_pasta"),
         ("nova_bn_o1", "blake3_nova_o1")]


def rows_for(variant, n, first):
    if variant == "compression":
        return np.concatenate([gen.lcg_compression_inputs(n // 2, first=first), gen.splitmix_compression_inputs(n - n // 2, first=first)])
    rows = gen.splitmix_nova_inputs(n, first=first)
    rows[7, 14] = rows[7, 12]                      # one instance that fails CheckDepth: "Assert Failed."
    return rows


def d2h(ptr, nbytes):
    from cuda.bindings import runtime as cudart
    host = np.empty(nbytes, np.uint8)
    err, = cudart.cudaMemcpy(host.ctypes.data, ptr, nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    assert int(err) == 0
    return host


@pytest.fixture(scope="module")
def oracle_cache():
    return {}


def oracle(cache, variant, rows):
    key = (variant, rows.shape[0])
    if key not in cache:
        cache[key] = port.witness_batch(variant, rows, nthreads=NCPU, want="both")
    return cache[key]


@pytest.mark.parametrize("checked", [False, True], ids=["plain", "checked"])
@pytest.mark.parametrize("variant,name", CASES, ids=[c[0] for c in CASES])
def test_every_byte_in_a_compressible_buffer(built, oracle_cache, variant, name, checked):
    wc = pkg.builder(name, device=0)
    ws, npub = wc.witnessSize, wc.nPublic
    rows = rows_for(variant, N, 31)
    want, want_sums, want_st = oracle(oracle_cache, variant, rows)
    ptr, granted = wc.device_alloc(N * ws * 32, compressible=True)
    assert granted, "B200 grants CU_MEM_ALLOCATION_COMP_GENERIC"
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_st = torch.full((N,), 255, dtype=torch.uint8, device="cuda")
    d_pub = torch.zeros(N * npub, dtype=torch.int32, device="cuda")
    d_bad = torch.zeros(N, dtype=torch.int32, device="cuda")
    d_sums = torch.full((N,), -1, dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    wc.witness_batch_device_ex(d_in.data_ptr(), N, ptr, d_st.data_ptr(), d_pub.data_ptr(), d_bad.data_ptr() if checked else 0,
                               d_sums.data_ptr(), check=checked, stream=s)
    torch.cuda.synchronize()
    st = d_st.cpu().numpy()
    ok = want_st == 0
    assert np.array_equal(st == 0, ok) and (st[~ok] == _lib.B3W_CIRCOM_ASSERT).all()
    got = d2h(ptr, N * ws * 32).reshape(N, ws * 32)
    assert np.array_equal(got[ok], want[ok])                                   # every byte vs Oracle B
    sums = d_sums.cpu().numpy().view(np.uint64)
    assert np.array_equal(sums[ok], want_sums[ok]) and (sums[~ok] == 0).all()  # the fused checksum = the oracle's
    assert np.array_equal(sums[ok], checksum_np(got[ok], ws))                  # ... = the checksum of the bytes in HBM
    pub = d_pub.cpu().numpy().view(np.uint32).reshape(N, npub)
    assert np.array_equal(pub[ok], want.view(np.uint32).reshape(N, ws, 8)[ok][:, 1:1 + npub, 0])
    if checked:
        assert (d_bad.cpu().numpy().view(np.uint32) == _lib.B3W_NO_ROW).all()
    # a sample against Oracle A, the reference's own witness program
    if ref_wasm.available(variant):
        sel = np.array([0, 1, 2, N // 2, N - 2, N - 1, 1000, 3001])
        ref = ref_wasm.RefWasm(variant)
        a_out, a_st, _ = ref.batch_u32(rows[sel], nthreads=min(NCPU, 8))
        assert (a_st == 0).all() and np.array_equal(got[sel], a_out)
    wc.device_free(ptr)
    wc.close()


@pytest.mark.parametrize("fused", [False, True], ids=["plain", "checked"])
@pytest.mark.parametrize("variant,name", CASES, ids=[c[0] for c in CASES])
def test_every_byte_through_the_default_ring(built, oracle_cache, variant, name, fused):
    """the host-buffer call as a user makes it: the ring is compressible by default (no flag)"""
    wc = pkg.builder(name, device=0, chunk=1024, fused_check=fused)
    rows = rows_for(variant, N, 31)
    want, want_sums, want_st = oracle(oracle_cache, variant, rows)
    ok = want_st == 0
    res = wc.calculateWitnessBatch(rows, sums=True, first_bad=True)
    assert np.array_equal(res["status"] == 0, ok)
    assert np.array_equal(res["witness"][ok], want[ok])
    assert np.array_equal(res["sums"][ok], want_sums[ok])
    assert (res["first_bad"] == _lib.B3W_NO_ROW).all()
    wc.close()


def test_plain_ring_flag_gives_the_same_bytes(built, oracle_cache):
    rows = rows_for("compression", N, 31)
    want, _, _ = oracle(oracle_cache, "compression", rows)
    wc = pkg.builder("blake3_compression", device=0, chunk=512, compressible_ring=False)
    assert np.array_equal(wc.calculateWitnessBatch(rows)["witness"], want)
    wc.close()


def test_tma_store_mode_gives_the_same_bytes(built, oracle_cache):
    """b3w_debug_set_store_mode(1): expanded tiles staged in shared memory and written by the TMA engine (the store path
    BASELINE's north_star sketches; an experiment switch, see profiles/r02_store_mode.jsonl) -- every byte vs Oracle B"""
    rows = rows_for("compression", N, 31)
    want, _, _ = oracle(oracle_cache, "compression", rows)
    wc = pkg.builder("blake3_compression", device=0)
    ws = wc.witnessSize
    wc.set_store_mode(1)
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_st = torch.full((N,), 255, dtype=torch.uint8, device="cuda")
    d_pub = torch.zeros(N * 16, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for compressible in (False, True):
        for parts in (0, 7):                                      # 7 items per witness: ragged tile tails
            wc.set_launch(0, parts)
            ptr, granted = wc.device_alloc(N * ws * 32, compressible=compressible)
            wc.witness_batch_device(d_in.data_ptr(), N, ptr, d_st.data_ptr(), d_pub.data_ptr(), s)
            torch.cuda.synchronize()
            assert int(d_st.max()) == 0
            assert np.array_equal(d2h(ptr, N * ws * 32).reshape(N, ws * 32), want), (compressible, parts)
            wc.device_free(ptr)
    wc.close()
