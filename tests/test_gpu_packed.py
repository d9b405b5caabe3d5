"""GPU tests of the compact witness format: packed generation + on-device unpack must reproduce
the expanded witnesses bit for bit (checker: the C oracle and the expanding kernels)."""
import numpy as np
import pytest
import torch

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import port

pytestmark = pytest.mark.gpu

VARIANTS = [("blake3_compression", "compression"), ("blake3_nova", "nova_bn_o2"), ("blake3_nova_pasta", "nova_pasta_o2"),
            ("blake3_nova_o1", "nova_bn_o1")]


def rows_for(name, n):
    if name == "blake3_compression":
        return np.concatenate([gen.lcg_compression_inputs(n // 2), gen.splitmix_compression_inputs(n - n // 2)])
    rows = gen.splitmix_nova_inputs(n)
    rows[5, 14] = rows[5, 12]                 # depth == leaf_depth: "Assert Failed."
    return rows


@pytest.mark.parametrize("name,variant", VARIANTS)
def test_packed_then_unpack_equals_oracle(built, name, variant):
    wc = pkg.builder(name, device=0)
    n = 257
    rows = rows_for(name, n)
    res = wc.calculateWitnessBatchPacked(rows)
    assert res["packed"].shape == (n, wc.packedWords) and wc.packedWords * 4 < wc.witnessSize * 32 // 100
    want_status = np.zeros(n, np.uint8)
    if name != "blake3_compression":
        want_status[5] = 4
    assert np.array_equal(res["status"], want_status)
    ok = want_status == 0
    assert (res["packed"][ok, 1] == 1).all() and (res["packed"][~ok, 1] == 0).all()
    got = wc.unpackWitnesses(res["packed"][ok])
    want = port.witness_batch(variant, rows[ok], nthreads=4)
    assert np.array_equal(got, want)
    full = wc.calculateWitnessBatch(rows[ok])
    assert np.array_equal(res["pub"][ok], full["pub"])
    wc.close()


def test_packed_device_path_matches_expanding_kernel_2p14(built):
    wc = pkg.builder("blake3_compression", device=0)
    n = 1 << 14
    rows = gen.splitmix_compression_inputs(n)
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_pk = torch.empty((n, wc.packedWords), dtype=torch.int32, device="cuda")
    d_a = torch.empty(n * wc.witnessSize * 32, dtype=torch.uint8, device="cuda")
    d_b = torch.zeros(n * wc.witnessSize * 32, dtype=torch.uint8, device="cuda")
    d_st = torch.ones(n, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    wc.witness_batch_device(d_in.data_ptr(), n, d_a.data_ptr(), 0, 0, s)
    wc.witness_batch_packed_device(d_in.data_ptr(), n, d_pk.data_ptr(), d_st.data_ptr(), 0, s)
    wc.unpack_device(d_pk.data_ptr(), n, d_b.data_ptr(), s)
    torch.cuda.synchronize()
    assert int(d_st.max()) == 0
    assert torch.equal(d_a, d_b)
    wc.close()
