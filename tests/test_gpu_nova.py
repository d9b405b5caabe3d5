"""GPU parity tests for the blake3_nova / blake3_nova_pasta step circuits (three committed builds)."""
import os

import numpy as np
import pytest
import torch

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import port, ref_wasm
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1
VARIANTS = [("nova_bn_o2", "blake3_nova", 23291), ("nova_pasta_o2", "blake3_nova_pasta", 23291),
            ("nova_bn_o1", "blake3_nova_o1", 24614)]
NAMES = [("n_blocks", 0, 1), ("block_count", 1, 1), ("h", 2, 8), ("chunk_idx_low", 10, 1), ("chunk_idx_high", 11, 1),
         ("leaf_depth", 12, 1), ("total_depth", 13, 1), ("depth", 14, 1), ("m", 15, 16), ("b", 31, 1)]


def as_input(row):
    row = [int(x) for x in row]
    return {k: (row[o] if n == 1 else row[o:o + n]) for k, o, n in NAMES}


@pytest.fixture(scope="module", params=VARIANTS, ids=[v[0] for v in VARIANTS])
def nv(request, built):
    variant, name, ws = request.param
    wc = pkg.builder(name, device=0)
    assert wc.witnessSize == ws and wc.nInputs == 32 and wc.nPublic == 15
    yield variant, wc
    wc.close()


def test_reference_fixture_cases(nv):
    variant, wc = nv
    fx = np.load(os.path.join(GOLDEN, "%s_cases.npz" % variant))
    res = wc.calculateWitnessBatch(fx["rows"])
    assert list(res["status"]) == list(fx["status"])          # incl. "Assert Failed." (code 4) for the two bad inputs
    ok = fx["status"] == 0
    assert np.array_equal(res["witness"][ok], fx["witness"][ok])


def test_against_port_full_bytes(nv):
    variant, wc = nv
    rows = gen.splitmix_nova_inputs(1536, first=11)
    want, _, st = port.witness_batch(variant, rows, nthreads=NCPU, want="both")
    assert (st == 0).all()
    res = wc.calculateWitnessBatch(rows)
    assert (res["status"] == 0).all()
    assert np.array_equal(res["witness"], want)
    # pub = z_{i+1} = witness slots 1..15 (low 32 bits)
    z = want.view(np.uint32).reshape(len(rows), wc.witnessSize, 8)[:, 1:16, 0]
    assert np.array_equal(res["pub"], z)


@pytest.mark.skipif(not ref_wasm.available("nova_bn_o2"), reason="oracle/_ref not shipped")
def test_against_reference_wasm_live(nv):
    variant, wc = nv
    rows = gen.splitmix_nova_inputs(24, first=5000)
    rows[3, 14] = rows[3, 12]                                  # depth == leaf_depth -> assert
    rows[5, 14], rows[5, 12], rows[5, 13] = 4000000123, 4000000124, 77      # generic (non-table) inverses
    ref = ref_wasm.RefWasm(variant)
    want, st, _ = ref.batch_u32(rows, nthreads=min(NCPU, 24))
    res = wc.calculateWitnessBatch(rows)
    assert list(res["status"]) == [int(x) for x in st]
    ok = st == 0
    assert ok.sum() == 23
    assert np.array_equal(res["witness"][ok], want[ok])


def test_single_witness_api_and_assert(nv, capsys):
    variant, wc = nv
    row = gen.splitmix_nova_inputs(1, first=3)[0]
    w = wc.calculateBinWitness(as_input(row))
    assert "D_FLAGS:  0" in capsys.readouterr().out            # circuits/blake3_nova.circom:166 via console.log
    assert np.array_equal(w, port.witness_batch(variant, row[None, :], nthreads=1)[0])
    bad = as_input(row)
    bad["depth"] = bad["leaf_depth"]
    with pytest.raises(RuntimeError, match="Assert Failed"):
        wc.calculateWitness(bad)
    extra = as_input(row)
    extra["override_h_to_IV"] = 1        # rust_fold passes it (blake3_circuit.rs:260-265); no committed wasm declares it
    with pytest.raises(RuntimeError, match="Too many values for input signal override_h_to_IV"):
        wc.calculateWitness(extra)


def test_2p15_resident_checksums(nv):
    variant, wc = nv
    n = 1 << 15
    ws = wc.witnessSize
    rows = gen.splitmix_nova_inputs(n)
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_out = torch.empty(n * ws * 32, dtype=torch.uint8, device="cuda")
    d_st = torch.full((n,), 255, dtype=torch.uint8, device="cuda")
    d_sum = torch.empty(n, dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    wc.witness_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), 0, s)
    wc.checksum_device(d_out.data_ptr(), n, d_sum.data_ptr(), s)
    torch.cuda.synchronize()
    assert int(d_st.max()) == 0
    want = port.witness_batch(variant, rows, nthreads=NCPU, want="sums")
    assert np.array_equal(d_sum.cpu().numpy().view(np.uint64), want)
