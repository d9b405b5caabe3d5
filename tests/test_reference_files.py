"""Checks that need the reference tree itself (run only where /root/reference exists, i.e. in the build container)."""
import os

import pytest

import hot_proofs_blake3_circom_b200 as pkg

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "build")), reason="reference tree not present")


@pytest.mark.parametrize("path,cid", [
    ("build/blake3_compression/blake3_compression_js/blake3_compression.wasm", 0),
    ("build/blake3_nova_js/blake3_nova.wasm", 1),
    ("build/blake3_nova_pasta_js/blake3_nova_pasta.wasm", 2),
    ("build/blake3_nova/blake3_nova_js/blake3_nova.wasm", 3),
    ("build/blake3_nova_pasta/blake3_nova_pasta_js/blake3_nova_pasta.wasm", 3),
])
def test_builder_identifies_the_reference_wasm_files(path, cid):
    with open(os.path.join(REF, path), "rb") as f:
        assert pkg.circuit_from_wasm(f.read()) == cid


def test_generated_tables_are_up_to_date(built):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # re-derives all four tables from the reference wasm files and the circuit model (validating the model
    # signal-by-signal against the wasm memory on the way) and compares with the committed headers
    subprocess.run([sys.executable, os.path.join(root, "tools", "gen_tables.py"), "--check-only", "--trials", "2"],
                   check=True)


def test_exported_sym_equals_the_committed_sym(built, tmp_path):
    """tools/export_r1cs.py's .sym for blake3_compression against build/blake3_compression/blake3_compression.sym:
    every one of the 69 380 rows has the same label, wire and name (the template index column is not re-derived)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import export_r1cs as ex
    ex.export("compression", str(tmp_path), trials=1, verbose=False)

    def rows(path):
        out = []
        with open(path) as f:
            for line in f:
                lab, wire, _, name = line.rstrip("\n").split(",", 3)
                out.append((int(lab), int(wire), name))
        return out
    mine = rows(os.path.join(str(tmp_path), "blake3_compression.sym"))
    ref = rows(os.path.join(REF, "build/blake3_compression/blake3_compression.sym"))
    assert len(ref) == 69380 and mine == ref
