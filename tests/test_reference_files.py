"""Checks that need the reference tree itself (run only where /root/reference exists, i.e. in the build container)."""
import os

import pytest

import hot_proofs_blake3_circom_b200 as pkg

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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


def test_js_drop_in_keeps_the_reference_surface():
    """integration/js cannot be executed here (no node): hold it to the reference's files textually -- the methods of the
    reference's WitnessCalculator, its error message templates (witness_calculator.js:143-167) and the CLI's usage line
    (generate_witness.js:5) appear in the drop-in, and the same messages in the Python mirror the tests execute."""
    import re
    ref = open(os.path.join(REF, "blake3_nova_js", "witness_calculator.js")).read()
    ours = open(os.path.join(ROOT, "integration", "js", "witness_calculator.js")).read()
    py = open(os.path.join(ROOT, "hot_proofs_blake3_circom_b200", "witness_calculator.py")).read()
    # the three generated copies of witness_calculator.js in the reference tree are one file
    for other in ("build/blake3_nova_js", "build/blake3_nova_pasta_js", "build/blake3_compression/blake3_compression_js"):
        assert open(os.path.join(REF, other, "witness_calculator.js")).read() == ref
    assert "module.exports = async function builder(code, options)" in ref and "module.exports = async function builder(code, options)" in ours
    methods = re.findall(r"^    (?:async )?(\w+)\(", ref[:ref.index("function toArray32")], flags=re.M)   # the class, not the helpers
    public = [m for m in methods if not m.startswith("_") and m != "constructor"]
    assert public == ["circom_version", "calculateWitness", "calculateBinWitness", "calculateWTNSBin"]
    for m in public:
        assert re.search(r"^    (?:async )?%s\(" % m, ours, flags=re.M), m
        assert re.search(r"^    def %s\(" % m, py, flags=re.M), m
    for field in ("version", "n32", "prime", "witnessSize", "sanityCheck", "instance"):
        assert "this.%s = " % field in ref and "this.%s = " % field in ours, field
    msgs = re.findall(r"throw new Error\(`([^`]*)`\)", ref)
    assert len(msgs) == 4
    for msg in msgs:
        head = msg.split("${")[0]
        assert ("`" + head) in ours, head
        assert ('"' + head) in py, head
    usage = re.search(r'console\.log\("(Usage: [^"]*)"\)', open(os.path.join(REF, "blake3_nova_js", "generate_witness.js")).read()).group(1)
    assert usage in open(os.path.join(ROOT, "integration", "js", "generate_witness.js")).read()
    # "Assert Failed." and the other exceptionHandler texts (witness_calculator.js:21-36)
    for code_msg in re.findall(r'errStr = "([^"]*)"', ref):
        assert code_msg in ours or code_msg in open(os.path.join(ROOT, "integration", "js", "addon", "blake3wit_napi.cc")).read() \
            or code_msg in open(os.path.join(ROOT, "hot_proofs_blake3_circom_b200", "csrc", "blake3wit.cu")).read(), code_msg
