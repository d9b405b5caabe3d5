"""tools/export_r1cs.py: the regenerated .r1cs / .sym artefacts (absent from the reference tree).
The exporter itself asserts that witnesses of the reference's own wasm satisfy every exported row; here the files are
read back and checked against the reference's GOLDEN witness and the documented counts."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import ref_wasm  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_wasm.available("compression"), reason="oracle/_ref not built")
BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


@needs_ref
def test_compression_r1cs_roundtrip_and_golden_witness(tmp_path, golden):
    import export_r1cs as ex
    path, rows = ex.export("compression", str(tmp_path), trials=1, verbose=False)
    r = ex.read_r1cs(path)
    # 24 544 constraints at O1 (23 376 quadratic + 1 168 linear); nPublic = 16 (groth16_vkey.json)
    assert (r["prime"], r["n_wires"], r["n_pub_out"], r["n_pub_in"], r["n_prv_in"]) == (BN254_R, 24093, 16, 0, 28)
    assert len(r["rows"]) == 24544 and sum(1 for A, B, C in r["rows"] if A) == 23376
    assert r["n_labels"] == 69381 and len(r["wire2label"]) == 24093 and r["wire2label"][0] == 0
    # the reference's golden witness (build/blake3_compression/testInp/witness.wtns) satisfies every row read back
    body = golden["wtns"][76:].tobytes()
    w = [int.from_bytes(body[32 * i:32 * i + 32], "little") for i in range(24093)]
    assert ex.check_rows(r["rows"], w, BN254_R) is None
    # ... and a single corrupted slot does not
    for slot in (1, 44, 1665, 24092):
        w2 = list(w)
        w2[slot] = (w2[slot] + 1) % BN254_R
        assert ex.check_rows(r["rows"], w2, BN254_R) is not None, slot
    # .sym: same label -> wire pairs as the reference's committed file wherever that file assigns a wire
    sym = {}
    with open(os.path.join(str(tmp_path), "blake3_compression.sym")) as f:
        for line in f:
            lab, wire, _, name = line.rstrip("\n").split(",", 3)
            sym[int(lab)] = (int(wire), name)
    assert len(sym) == 69380
    assert sym[1][1] == "main.out[0]" and sym[1][0] == 1


@needs_ref
def test_nova_o2_r1cs_counts(tmp_path):
    import export_r1cs as ex
    if not ref_wasm.available("nova_pasta_o2"):
        pytest.skip("oracle/_ref not built")
    path, rows = ex.export("nova_pasta_o2", str(tmp_path), trials=1, verbose=False)
    r = ex.read_r1cs(path)
    assert r["prime"] == 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001
    # 1 + 15 outputs + 12 public inputs = 28 = num_inputs on the Rust side (rust_fold/src/utils.rs:33)
    assert (r["n_wires"], r["n_pub_out"], r["n_pub_in"], r["n_prv_in"]) == (23291, 15, 12, 20)
    # O1 has 25 064 rows over 24 614 wires; the O2 pass removes one linear row per dropped wire (1 323)
    assert len(r["rows"]) == 23743 and sum(1 for A, B, C in r["rows"] if not A) == 3
