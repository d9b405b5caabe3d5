"""CPU tests of the drop-in boundary: the C ABI library loads, exports what include/blake3wit.h declares,
needs no GPU for metadata, and refuses loudly to compute without one."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "blake3wit.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b3w_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built):
    L = pkg.lib()
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), "libblake3wit.so does not export %s" % s
    assert set(syms) == set(_lib.EXPORTS)
    assert L.b3w_version() == _lib.B3W_VERSION == 0x000200


def test_no_torch_types_in_header():
    hdr = open(os.path.join(ROOT, "include", "blake3wit.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", hdr, flags=re.S).lower()
    assert "at::" not in hdr and "c10" not in hdr


def test_circuit_info_matches_reference_constants(built):
    L = pkg.lib()
    info = _lib.Info()
    assert L.b3w_circuit_info(0, C.byref(info)) == 0
    assert (info.witness_size, info.n_inputs, info.n32, info.n_public) == (24093, 28, 8, 16)
    assert list(info.version) == [2, 1, 6]
    # test/blake3_hash.test.ts:10-12
    assert int.from_bytes(bytes(info.prime), "little") == \
        21888242871839275222246405745257275088548364400416034343698204186575808495617
    assert L.b3w_circuit_info(99, C.byref(info)) == _lib.B3W_ERR_UNSUPPORTED


def test_wtns_header_equals_reference_golden(golden, built):
    hdr = np.zeros(76, np.uint8)
    assert pkg.lib().b3w_wtns_header(0, hdr.ctypes.data) == 0
    assert hdr.tobytes() == golden["wtns"].tobytes()[:76]


def test_inputs_from_fr_reduces_and_refuses(built):
    L = pkg.lib()
    P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    vals = [5, P + 7, 2 * P + 0xFFFFFFFF, 0] + list(range(24))
    fr = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), np.uint8).copy()
    rows = np.zeros(28, np.uint32)
    assert L.b3w_inputs_from_fr(0, fr.ctypes.data, 1, rows.ctypes.data) == 0
    assert list(rows) == [5, 7, 0xFFFFFFFF, 0] + list(range(24))
    vals[9] = 2 ** 32                                            # m[1]
    fr = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), np.uint8).copy()
    assert L.b3w_inputs_from_fr(0, fr.ctypes.data, 1, rows.ctypes.data) == _lib.B3W_ERR_DOMAIN
    assert b"input m[1]" in L.b3w_last_error()
    vals[9] = P - 1                                              # "-1": legal for the wasm, outside the u32 domain here
    fr = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), np.uint8).copy()
    assert L.b3w_inputs_from_fr(0, fr.ctypes.data, 1, rows.ctypes.data) == _lib.B3W_ERR_DOMAIN


def test_packed_sizes(built):
    L = pkg.lib()
    w = C.c_uint32()
    for cid, words in ((0, 944), (1, 1324), (2, 1324), (3, 1324)):
        assert L.b3w_packed_words(cid, C.byref(w)) == 0 and w.value == words
    assert L.b3w_packed_words(9, C.byref(w)) == _lib.B3W_ERR_UNSUPPORTED


def test_input_signal_table(built):
    L = pkg.lib()
    off, size = C.c_uint32(), C.c_uint32()
    want = {"h": (0, 8), "m": (8, 16), "t": (24, 2), "b": (26, 1), "d": (27, 1)}
    for k, v in want.items():
        assert L.b3w_input_signal(0, k.encode(), C.byref(off), C.byref(size)) == 0
        assert (off.value, size.value) == v
    assert L.b3w_input_signal(0, b"nope", C.byref(off), C.byref(size)) == _lib.B3W_ERR_INVALID
    assert b"nope" in L.b3w_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_compute_fails_loudly_without_gpu(built):
    with pytest.raises(pkg.B3WError) as e:
        pkg.builder("blake3_compression")
    assert e.value.code == _lib.B3W_ERR_CUDA
    assert "no CPU path" in str(e.value)
    wc = pkg.builder("blake3_compression", lazy=True)
    with pytest.raises(pkg.B3WError):
        wc.calculateWitness({"h": [0] * 8, "m": [0] * 16, "t": [0, 0], "b": 0, "d": 0})


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_multi_gpu_entry_fails_loudly_without_gpu(built):
    with pytest.raises(pkg.B3WError) as e:
        pkg.MultiGpuCalculator("blake3_compression")
    assert e.value.code == _lib.B3W_ERR_CUDA and "no CPU path" in str(e.value)


def test_shard_range_equals_the_python_plan(built):
    # the C ABI's index-range split (b3w_multi_witness_batch) and shard.py (bench.py / torch.distributed ranks) agree
    from hot_proofs_blake3_circom_b200.shard import shard_range
    L = pkg.lib()
    first, count = C.c_uint64(), C.c_uint64()
    for n in (0, 1, 7, 37, 65536, 2 ** 24 + 3):
        for world in (1, 2, 3, 4, 8):
            for r in range(world):
                assert L.b3w_shard_range(n, r, world, C.byref(first), C.byref(count)) == 0
                assert (first.value, count.value) == shard_range(n, r, world)
    assert L.b3w_shard_range(5, 2, 2, C.byref(first), C.byref(count)) == _lib.B3W_ERR_INVALID


def test_product_does_not_import_oracle():
    # the product path must never route through oracle/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hot_proofs_blake3_circom_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp", ".c")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src.replace("tools/", ""), f


def test_sass_streaming_stores_are_wide(built):
    """ptxas 12.9 once turned a v8.b32 store with eight computed operands into a 32-bit STG inside a non-inlined
    function (only limb 0 of nova's field slots reached memory).  Every no-allocate store in the library must be a
    256-bit or 128-bit STG; the hot path must have the 256-bit form."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", pkg.lib_path()], capture_output=True, text=True, check=True).stdout
    na = [l for l in sass.splitlines() if "STG.E.NA" in l]
    assert len(na) > 10
    assert all(".256" in l or ".128" in l for l in na), [l for l in na if ".256" not in l and ".128" not in l][:3]
    assert sum(".256" in l for l in na) > 10


def test_header_is_plain_c_and_links_from_c(built, tmp_path):
    """the boundary is a C ABI: include/blake3wit.h compiles as strict C99 and a C program links against the library"""
    import subprocess
    src = tmp_path / "c_abi.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "blake3wit.h"\n'
                   'int main(void) {\n'
                   '  b3w_info i; unsigned char hdr[76]; uint32_t off, size;\n'
                   '  if (b3w_version() != B3W_VERSION) return 1;\n'
                   '  if (b3w_circuit_info(B3W_COMPRESSION, &i) != B3W_OK || i.witness_size != 24093u) return 2;\n'
                   '  if (b3w_wtns_header(B3W_NOVA_PASTA_O2, hdr) != B3W_OK || memcmp(hdr, "wtns", 4) != 0) return 3;\n'
                   '  if (b3w_input_signal(B3W_NOVA_BN_O2, "m", &off, &size) != B3W_OK || off != 15u || size != 16u) return 4;\n'
                   '  if (b3w_input_signal(B3W_COMPRESSION, "nope", &off, &size) != B3W_ERR_INVALID) return 5;\n'
                   '  printf("%s\\n", b3w_last_error());\n  return 0;\n}\n')
    exe = tmp_path / "c_abi"
    libdir = os.path.dirname(pkg.lib_path())
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-L", libdir, "-lblake3wit", "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.returncode
    assert "Signal nope not found" in r.stdout


def test_builtin_systems_compile_completely(built):
    """fp_compile (host-only): every row of the four built-in constraint systems is covered by the compiled program of the
    stand-alone checker -- booleanity masks, XOR runs, row tiles -- so the general evaluator has nothing left to do"""
    L = pkg.lib()
    want = {0: (24544, 464), 1: (23743, 812), 2: (23743, 812), 3: (25064, 464)}      # O2: XOR rows over virtual bits split runs
    for circuit, (rows, xor_runs) in want.items():
        v = [C.c_uint32() for _ in range(5)]
        assert L.b3w_r1cs_compile_stats(circuit, *[C.byref(x) for x in v]) == 0
        n_rows, n_compiled, n_xor, n_tiles, n_items = (x.value for x in v)
        assert (n_rows, n_compiled, n_xor) == (rows, rows, xor_runs), (circuit, n_rows, n_compiled, n_xor)
        assert 15 < n_tiles < 100 and n_items < 20000
        ex = (C.c_uint32 * 9)()
        assert L.b3w_r1cs_compile_stats_ex(circuit, ex, 9) == 0
        assert list(ex[:5]) == [n_rows, n_compiled, n_xor, n_tiles, n_items]
        n_virtual, n_fast, plain_tiles, plain_items = ex[5:9]
        if circuit in (1, 2):
            # the O2-form system: every bit circom substituted by a linear combination is a virtual bit; the rows that carried
            # them are booleanity / XOR / short rows now, and nearly every tile sums in 64 bits; the plainly compiled program
            # (the fall-back for witnesses whose virtual bits are not bits) has 2.6 times the tiles
            assert n_virtual == 701 and plain_tiles == 55 and plain_items > 1.5 * n_items and n_fast >= n_tiles - 8
        else:
            assert (n_virtual, plain_tiles, plain_items) == (0, 0, 0) and n_fast >= n_tiles - 8


def test_extras_and_timing_structs_match_the_header():
    hdr = open(os.path.join(ROOT, "include", "blake3wit.h")).read()
    assert "uint64_t *sums;" in hdr and "const uint64_t *sample_idx;" in hdr and "uint32_t *first_bad;" in hdr
    assert C.sizeof(_lib.BatchExtras) == 40 and C.sizeof(_lib.Timing) == 56
    assert _lib.B3W_MAX_SAMPLES == 1024
