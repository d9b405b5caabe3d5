"""GPU tests of the host-side expansion of packed witnesses (b3w_unpack_host) and the hybrid export built on it
(b3w_witness_batch_hybrid: only the 3.8 / 5.3 KB packed records cross PCIe, the 200x expansion to .wtns bodies runs on host
threads -- the .wtns writer of witness_calculator.js:208-272 as a lazy, multithreaded export).  Checker: Oracle B."""
import os

import numpy as np
import pytest

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import port

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1
CASES = [("blake3_compression", "compression", gen.splitmix_compression_inputs), ("blake3_nova", "nova_bn_o2", gen.splitmix_nova_inputs),
         ("blake3_nova_pasta", "nova_pasta_o2", gen.splitmix_nova_inputs), ("blake3_nova_o1", "nova_bn_o1", gen.splitmix_nova_inputs)]


@pytest.mark.parametrize("name,variant,rows_fn", CASES, ids=[c[1] for c in CASES])
def test_unpack_host_equals_the_oracle(built, name, variant, rows_fn):
    wc = pkg.builder(name, device=0)
    rows = rows_fn(700, first=13)
    want = port.witness_batch(variant, rows, nthreads=NCPU)
    pk = wc.calculateWitnessBatchPacked(rows)
    assert not pk["status"].any()
    for threads in (1, 0):
        assert np.array_equal(wc.unpackWitnessesHost(pk["packed"], threads=threads), want)
    # an unaligned destination takes the plain-store path: same bytes
    buf = np.zeros(700 * wc.witnessSize * 32 + 8, np.uint8)
    out = buf[8:].reshape(700, -1)
    _lib.check(pkg.lib().b3w_unpack_host(wc._h, pk["packed"].ctypes.data, 700, out.ctypes.data, 3))
    assert np.array_equal(out, want)
    wc.close()


@pytest.mark.parametrize("name,variant,rows_fn", CASES[:2], ids=[c[1] for c in CASES[:2]])
def test_hybrid_batch(built, name, variant, rows_fn):
    wc = pkg.builder(name, device=0)
    n = 3000
    rows = rows_fn(n, first=2)
    if variant != "compression":
        rows[9, 14] = rows[9, 12]                     # asserts: status 4; its output slot is not a witness
    want, _, st = port.witness_batch(variant, rows, nthreads=NCPU, want="both")
    res = wc.calculateWitnessBatchHybrid(rows)
    ok = st == 0
    assert np.array_equal(res["status"] == 0, ok)
    assert np.array_equal(res["witness"][ok], want[ok])
    assert np.array_equal(res["pub"][ok], want.view(np.uint32).reshape(n, wc.witnessSize, 8)[ok][:, 1:1 + wc.nPublic, 0])
    t = wc.lastTiming()
    assert t["host_ms"] > 0 and t["d2h_bytes"] < n * wc.witnessSize * 32 / 100
    wc.close()


def test_hybrid_spans_several_packed_chunks(built):
    wc = pkg.builder("blake3_compression", device=0)
    n = (1 << 15) * 2 + 77
    rows = gen.splitmix_compression_inputs(n, first=0)
    res = wc.calculateWitnessBatchHybrid(rows)
    assert not res["status"].any()
    sums = port.witness_batch("compression", rows, nthreads=NCPU, want="sums")
    from conftest import checksum_np
    for lo in range(0, n, 8192):
        assert np.array_equal(checksum_np(res["witness"][lo:lo + 8192], wc.witnessSize), sums[lo:lo + 8192])
    wc.close()
