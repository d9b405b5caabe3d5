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
