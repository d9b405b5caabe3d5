"""CPU test: the three bindings of include/blake3wit.h that cannot all be executed here (ctypes: executed; Rust FFI and JS:
no cargo / node in the image) are held to the header mechanically -- struct layouts as gcc lays them out, the values of the
#defines and enum members, and the argument counts of every function the Rust extern block declares."""
import ctypes as C
import os
import re
import subprocess

from hot_proofs_blake3_circom_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "blake3wit.h")
STRUCTS = {"b3w_config": (_lib.Config, "B3wConfig"), "b3w_info": (_lib.Info, "B3wInfo"),
           "b3w_batch_extras": (_lib.BatchExtras, "B3wBatchExtras"), "b3w_timing": (_lib.Timing, None)}
MACROS = ["B3W_VERSION", "B3W_OK", "B3W_ERR_INVALID", "B3W_ERR_CUDA", "B3W_ERR_NOMEM", "B3W_ERR_DOMAIN", "B3W_ERR_UNSUPPORTED",
          "B3W_CIRCOM_ASSERT", "B3W_R1CS_VIOLATION", "B3W_NO_ROW", "B3W_FLAG_FUSED_CHECK", "B3W_FLAG_COMPRESSIBLE_RING",
          "B3W_FLAG_PLAIN_RING", "B3W_FLAG_REFERENCE_SIBLINGS", "B3W_FLAG_BYTE_CHECK", "B3W_MEM_COMPRESSIBLE", "B3W_MAX_SAMPLES",
          "B3W_COMPRESSION", "B3W_NOVA_BN_O2", "B3W_NOVA_PASTA_O2", "B3W_NOVA_BN_O1"]


def c_facts(tmp_path):
    """sizeof / offsetof of every struct field and the value of every macro, from gcc"""
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "blake3wit.h"', 'int main(void) {']
    for cname, (ct, _) in STRUCTS.items():
        lines.append('  printf("sizeof %s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f, _t in ct._fields_:
            lines.append('  printf("offsetof %s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
    for m in MACROS:
        lines.append('  printf("value %s %%lld\\n", (long long)%s);' % (m, m))
    lines += ['  return 0;', '}']
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines) + "\n")
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    facts = {}
    for ln in out.splitlines():
        kind, name, val = ln.split()
        facts[(kind, name)] = int(val)
    return facts


def test_ctypes_structs_have_the_layout_gcc_gives_the_header(tmp_path):
    facts = c_facts(tmp_path)
    for cname, (ct, _) in STRUCTS.items():
        assert C.sizeof(ct) == facts[("sizeof", cname)], cname
        for f, _t in ct._fields_:
            assert getattr(ct, f).offset == facts[("offsetof", "%s.%s" % (cname, f))], (cname, f)
    # every field of the header's structs is bound (a field added to the header and forgotten in _lib.py changes sizeof)
    for m in MACROS:
        if hasattr(_lib, m):
            assert getattr(_lib, m) == facts[("value", m)], m
    assert _lib.B3W_VERSION == facts[("value", "B3W_VERSION")] == 0x000200


RUST_TYPES = {"u32": 4, "i32": 4, "u64": 8, "*mut u64": 8, "*const u64": 8, "*mut u8": 8, "*mut u32": 8, "[u32; 3]": 12, "[u8; 32]": 32}


def test_rust_ffi_follows_the_header(tmp_path):
    facts = c_facts(tmp_path)
    rs = open(os.path.join(ROOT, "integration", "rust", "blake3wit_ffi.rs")).read()
    hdr = open(HDR).read()
    # constants
    for name, val in re.findall(r"pub const (B3W_[A-Z0-9_]+): u32 = (\d+);", rs):
        assert facts[("value", name)] == int(val), name
    # #[repr(C)] structs: same field names in the same order, natural alignment gives the same offsets
    for cname, (ct, rname) in STRUCTS.items():
        if rname is None:
            continue
        body = re.search(r"pub struct %s \{([^}]*)\}" % rname, rs).group(1)
        fields = [(n, t.strip()) for n, t in re.findall(r"pub (\w+): ([^,]+?)(?:,|$)", body.strip())]
        assert [n for n, _ in fields] == [f for f, _ in ct._fields_], rname
        off = 0
        for n, t in fields:
            size = RUST_TYPES[t]
            align = 1 if t == "[u8; 32]" else 4 if t == "[u32; 3]" else size
            off = (off + align - 1) // align * align
            assert off == facts[("offsetof", "%s.%s" % (cname, n))], (rname, n)
            off += size
    # extern block: every function exists in the header with the same number of parameters
    ext = rs[rs.index('extern "C" {'):]
    ext = ext[:ext.index("\n}\n")]
    decls = re.findall(r"pub fn (b3w_\w+)\(([^)]*)\)", ext)
    assert len(decls) >= 14
    flat = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    for name, params in decls:
        m = re.search(r"\b%s\(([^)]*)\)\s*;" % name, flat)
        assert m, "%s is not declared in blake3wit.h" % name
        c_params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        r_params = [p for p in params.split(",") if p.strip()]
        assert len(c_params) == len(r_params), name


def test_js_layer_uses_the_headers_values(tmp_path):
    facts = c_facts(tmp_path)
    js = open(os.path.join(ROOT, "integration", "js", "witness_calculator.js")).read()
    m = re.search(r"const flags = \(options\.fusedCheck \? (\d+) : 0\) \| \(options\.byteCheck \? (\d+) : 0\);", js)
    assert m and int(m.group(1)) == facts[("value", "B3W_FLAG_FUSED_CHECK")] and int(m.group(2)) == facts[("value", "B3W_FLAG_BYTE_CHECK")]
    # the circuit ids behind the sha256 table = the enum, and the hashes = the ones the Python host layer routes on
    from hot_proofs_blake3_circom_b200 import witness_calculator as wcpy
    table = dict((h, int(i)) for h, i in re.findall(r'"([0-9a-f]{64})": (\d),', js))
    assert len(table) == 4 and sorted(table.values()) == [facts[("value", n)] for n in
                                                           ("B3W_COMPRESSION", "B3W_NOVA_BN_O2", "B3W_NOVA_PASTA_O2", "B3W_NOVA_BN_O1")]
    assert {k: v[0] for k, v in wcpy.CIRCUITS.items()} == table
