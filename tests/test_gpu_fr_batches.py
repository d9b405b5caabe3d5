"""GPU tests of MIXED-DOMAIN batches through b3w_witness_batch_fr: Fr256 rows are converted on the device, u32 instances run
on the hot kernels and only the instances that hold a field-valued input take the general path, into the same outputs
(reference: witness_calculator.js:319-323 accepts any field element per input; rust_fold holds Vec<F>,
rust_fold/src/blake3_circuit.rs:197-289).  Checker: Oracle B, every byte of the u32 instances via its batch entry point and
every field-valued instance via its Fr entry point; checksums for both."""
import os
import sys

import numpy as np
import pytest

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import port
from conftest import checksum_np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1


def fr_rows(rows):
    """(n, k) u32 -> (n, k, 32) u8 little-endian field elements"""
    n, k = rows.shape
    fr = np.zeros((n, k, 32), np.uint8)
    fr[:, :, 0:4] = np.ascontiguousarray(rows).view(np.uint8).reshape(n, k, 4)
    return fr


def put(fr, i, vals):
    for k, v in enumerate(vals):
        fr[i, k] = np.frombuffer(int(v).to_bytes(32, "little"), np.uint8)


@pytest.mark.parametrize("fused", [False, True], ids=["plain", "checked"])
@pytest.mark.parametrize("name,variant", [("blake3_nova", "nova_bn_o2"), ("blake3_nova_pasta", "nova_pasta_o2"), ("blake3_nova_o1", "nova_bn_o1")])
def test_nova_one_percent_field_valued(built, name, variant, fused):
    import test_gpu_nova_wide as tn
    wc = pkg.builder(name, device=0, chunk=1024, fused_check=fused)
    n = 12000
    rows = gen.splitmix_nova_inputs(n, first=300)
    fr = fr_rows(rows)
    wide_idx = np.arange(17, n, 100)                                  # 1 % of the batch, spread over every ring chunk
    wide_vals = tn.random_rows(len(wide_idx), wc.prime, 5)
    for i, v in zip(wide_idx, wide_vals):
        put(fr, i, [x % wc.prime for x in v])
    res = wc.calculateWitnessBatchFr(fr, sums=True, first_bad=fused)
    # the u32 instances: every byte and every checksum vs Oracle B's batch entry point
    u32 = np.ones(n, bool)
    u32[wide_idx] = False
    want, want_sums, st = port.witness_batch(variant, rows[u32], nthreads=NCPU, want="both")
    assert (st == 0).all() and not res["status"][u32].any()
    assert np.array_equal(res["witness"][u32], want)
    assert np.array_equal(res["sums"][u32], want_sums)
    # the field-valued ones, one by one vs Oracle B's Fr entry point
    n_ok = 0
    for i, v in zip(wide_idx, wide_vals):
        rc, w = port.witness_fr(variant, [x % wc.prime for x in v])
        assert res["status"][i] == rc, (i, v)
        if rc == 0:
            n_ok += 1
            assert np.array_equal(res["witness"][i], w), (i, v)
            assert res["sums"][i] == checksum_np(w, wc.witnessSize)[0]
            assert np.array_equal(res["pub"][i], w.view(np.uint32).reshape(wc.witnessSize, 8)[1:16, 0])
        else:
            assert res["sums"][i] == 0 and not res["pub"][i].any()
    assert n_ok > 40
    if fused:
        assert (res["first_bad"] == _lib.B3W_NO_ROW).all()
    wc.close()


@pytest.mark.parametrize("fused", [False, True], ids=["plain", "checked"])
def test_compression_mixed_wide_message_words(built, fused):
    import test_gpu_wide as tw
    wc = pkg.builder("blake3_compression", device=0, chunk=512, fused_check=fused)
    n = 6000
    rows = gen.splitmix_compression_inputs(n, first=50)
    fr = fr_rows(rows)
    wide_idx = np.arange(3, n, 50)
    wide_vals = tw.random_wide_rows(len(wide_idx), 9, 1.0)
    for i, v in zip(wide_idx, wide_vals):
        put(fr, i, [x % wc.prime for x in v])
    res = wc.calculateWitnessBatchFr(fr, sums=True)
    u32 = np.ones(n, bool)
    u32[wide_idx] = False
    want, want_sums, st = port.witness_batch("compression", rows[u32], nthreads=NCPU, want="both")
    assert not res["status"][u32].any() and np.array_equal(res["witness"][u32], want) and np.array_equal(res["sums"][u32], want_sums)
    n_ok = 0
    for i, v in zip(wide_idx, wide_vals):
        rc, w = port.witness_fr("compression", [x % wc.prime for x in v])
        assert res["status"][i] == rc, (i, v)
        if rc == 0:
            n_ok += 1
            assert np.array_equal(res["witness"][i], w), (i, v)
            assert res["sums"][i] == checksum_np(w, wc.witnessSize)[0]
    assert n_ok > 20
    wc.close()


def test_non_canonical_values_are_reduced_on_the_device(built):
    """normalize(): BigInt(n) % p -- a value x + p, x + 2p is the same input (witness_calculator.js:319-323)"""
    for name, variant, rows in (("blake3_compression", "compression", gen.splitmix_compression_inputs(40)),
                                ("blake3_nova_pasta", "nova_pasta_o2", gen.splitmix_nova_inputs(40))):
        wc = pkg.builder(name, device=0)
        fr = fr_rows(rows)
        for i in range(40):
            for k in range(rows.shape[1]):
                if (i + k) % 3 == 0:
                    v = int(rows[i, k]) + wc.prime * (1 + (i + k) % 3)
                    if v < 2**256:
                        fr[i, k] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
        res = wc.calculateWitnessBatchFr(fr)
        assert np.array_equal(res["witness"], port.witness_batch(variant, rows, nthreads=4))
        wc.close()
