"""CPU tests of the full input domain of blake3_compression (host side: include/blake3wit.h b3w_inputs_from_fr_wide,
b3w_assert_trace_fr).  The reference reduces every input mod p and lets the circuit decide (witness_calculator.js:
319-323); tests/golden/compression_wide_cases.npz holds what its own witness program does with 80
such inputs (make_golden_wide.py), including the text it prints on "Assert Failed."."""
import ctypes as C
import os
import random

import numpy as np
import pytest

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from oracle import ref_wasm

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def wide_cases():
    return np.load(os.path.join(GOLDEN, "compression_wide_cases.npz"))


def fr_bytes(vals):
    return np.frombuffer(b"".join(int(v % P).to_bytes(32, "little") for v in vals), np.uint8).copy()


def convert(vals_rows):
    L = pkg.lib()
    n = len(vals_rows)
    fr = fr_bytes([x for v in vals_rows for x in v])
    rows, ext, nw = np.zeros((n, 28), np.uint32), np.zeros((n, 16), np.int8), C.c_uint64()
    assert L.b3w_inputs_from_fr_wide(0, fr.ctypes.data, n, rows.ctypes.data, ext.ctypes.data, C.byref(nw)) == 0
    return rows, ext, nw.value


def trace_of(vals):
    buf = C.create_string_buffer(1024)
    fr = fr_bytes(vals)                                # keep the array alive across the call (.ctypes.data is a bare address)
    rc = pkg.lib().b3w_assert_trace_fr(0, fr.ctypes.data, buf, len(buf))
    return rc, buf.value.decode()


def test_conversion_of_message_words(built):
    base = list(range(100, 128))
    cases = {2**32: (0, 1), 2**32 + 9: (9, 1), 2**34 - 1: (0xFFFFFFFF, 3), P - 1: (0xFFFFFFFF, -1), P - 2**32: (0, -1),
             P - 2**32 - 1: (0xFFFFFFFF, -2), P - 2**33: (0, -2), 2 * P + 5: (5, 0), 7: (7, 0)}
    for j, (x, (lo, e)) in enumerate(cases.items()):
        v = list(base)
        v[8 + j] = x
        rows, ext, nw = convert([v])
        want = list(base)
        want[8 + j] = lo
        assert list(rows[0]) == want, hex(x)
        assert ext[0, j] == e and int(np.abs(ext[0]).sum()) == abs(e)
        assert nw == (1 if e else 0)


def test_conversion_marks_hopeless_instances(built):
    base = list(range(100, 128))
    dead = []
    for k, x in ((0, 2**32), (7, P - 1), (24, 2**40), (25, 2**32), (26, 2**33), (27, P - 3), (8, 2**34), (23, P - 2**33 - 1),
                 (10, 2**200)):
        v = list(base)
        v[k] = x
        dead.append(v)
    rows, ext, nw = convert(dead + [base])
    assert nw == len(dead)
    assert (ext[:len(dead), 0] == _lib.B3W_EXT_ASSERT).all() and (ext[:len(dead), 1:] == 0).all()
    assert (ext[len(dead)] == 0).all() and list(rows[len(dead)]) == base


def test_wide_conversion_is_compression_only(built):
    L = pkg.lib()
    fr = np.zeros(32 * 32, np.uint8)
    rows, ext = np.zeros(32, np.uint32), np.zeros(16, np.int8)
    for cid in (1, 2, 3):
        assert L.b3w_inputs_from_fr_wide(cid, fr.ctypes.data, 1, rows.ctypes.data, ext.ctypes.data, None) == _lib.B3W_ERR_UNSUPPORTED
    assert L.b3w_inputs_from_fr_wide(9, fr.ctypes.data, 1, rows.ctypes.data, ext.ctypes.data, None) == _lib.B3W_ERR_UNSUPPORTED


def test_assert_trace_equals_reference_fixture(built, wide_cases):
    """status and the printErrorMessage lines of the reference's wasm, case by case"""
    fr, status, text = wide_cases["fr"], wide_cases["status"], wide_cases["text"]
    assert len(status) >= 80 and (status == 4).sum() >= 30 and (status == 0).sum() >= 30
    buf = C.create_string_buffer(1024)
    for i in range(len(status)):
        row = np.ascontiguousarray(fr[i])
        rc = pkg.lib().b3w_assert_trace_fr(0, row.ctypes.data, buf, len(buf))
        assert rc == status[i], i
        assert buf.value == bytes(text[i]), i


def test_assert_trace_names_the_surveyed_case(built, golden):
    """b = 2^33 on the golden input"""
    v = [int(x) for x in golden["row"]]
    v[26] = 2**33
    rc, txt = trace_of(v)
    assert rc == 4
    assert txt.startswith("Error in template ToBits_3 line: 153\nError in template RotXorWordBits_5 line: 62\n"
                          "Error in template HalfFunG_18 line: 91\n")
    assert txt.endswith("Error in template Blake3Compression_40 line: 194\n")
    v[26] = 64
    assert trace_of(v) == (0, "")


def test_assert_trace_nova_check_depth(built):
    L = pkg.lib()
    buf = C.create_string_buffer(1024)
    vals = [1, 0] + [0] * 8 + [0, 0, 3, 3, 5] + [0] * 16 + [64]          # depth 5 >= leaf_depth 3: CheckDepth asserts
    fr = fr_bytes(vals)
    assert L.b3w_assert_trace_fr(1, fr.ctypes.data, buf, len(buf)) == 4
    assert b"Blake3NovaTreePath_CheckDepth_5" in buf.value


@pytest.mark.skipif(not ref_wasm.available("compression"), reason="oracle/_ref not shipped")
def test_assert_trace_against_reference_wasm_live(built):
    """fresh random inputs (several wide values at once, cancelling h / m pairs) against the reference's own error path"""
    ref = ref_wasm.RefWasm("compression")
    rnd = random.Random(20261017)

    def wide_val():
        k = rnd.randrange(6)
        return [2**32 + rnd.randrange(2**32), P - 1 - rnd.randrange(2**20), P - rnd.randrange(1, 2**33),
                rnd.randrange(2**32, 2**34), rnd.randrange(P), 2**34 - 1 - rnd.randrange(2**31)][k]
    n_assert = 0
    for it in range(48):
        v = [rnd.randrange(2**32) for _ in range(28)]
        mode = it % 4
        if mode == 0:
            for k in rnd.sample(range(28), rnd.randrange(2, 6)):
                v[k] = wide_val()
        elif mode == 1:
            g = rnd.randrange(4)
            x = rnd.randrange(P)
            v[g], v[8 + 2 * g] = x, (P - x + rnd.randrange(2**32)) % P
        elif mode == 2:
            v[rnd.choice(list(range(8)) + [24, 25, 26, 27])] = wide_val()
        else:
            for j in rnd.sample(range(16), 3):
                v[8 + j] = wide_val()
        rc, _ = ref.calculate({"h": v[0:8], "m": v[8:24], "t": v[24:26], "b": v[26], "d": v[27]})
        if rc == 0:
            continue          # valid witnesses cost the wasm 0.3 s each; the GPU tests cover them
        n_assert += 1
        assert trace_of(v) == (rc, ref.err_msg()), v
    assert n_assert >= 20


def test_host_mirror_row_format_is_u32(built):
    wc = pkg.builder("blake3_nova", lazy=True)
    inp = {"n_blocks": 1, "block_count": 0, "h": [0] * 8, "chunk_idx_low": 0, "chunk_idx_high": 0, "leaf_depth": 1,
           "total_depth": 1, "depth": 0, "m": [2**32] + [0] * 15, "b": 64}
    with pytest.raises(pkg.B3WError) as e:
        wc._row(inp)
    assert e.value.code == _lib.B3W_ERR_DOMAIN and "m[0]" in str(e.value)
    wc0 = pkg.builder("blake3_compression", lazy=True)
    vals = wc0._values({"h": [0] * 8, "m": [-1] + [0] * 15, "t": [0, 0], "b": 64, "d": 0})
    assert vals[8] == P - 1                                  # normalize(): negatives wrap (witness_calculator.js:319-323)


# ---- nova step circuits: any 32 field elements ------------------------------------------------------------------------
NOVA = (("nova_bn_o2", 1), ("nova_pasta_o2", 2), ("nova_bn_o1", 3))


@pytest.fixture(scope="module")
def nova_wide_cases():
    return np.load(os.path.join(GOLDEN, "nova_wide_cases.npz"))


@pytest.mark.parametrize("variant,cid", NOVA)
def test_nova_assert_trace_equals_reference_fixture(built, nova_wide_cases, variant, cid):
    fr, status, text = (nova_wide_cases[variant + k] for k in ("_fr", "_status", "_text"))
    assert len(status) >= 30 and (status == 4).sum() >= 10 and (status == 0).sum() >= 10
    buf = C.create_string_buffer(1024)
    for i in range(len(status)):
        row = np.ascontiguousarray(fr[i])
        rc = pkg.lib().b3w_assert_trace_fr(cid, row.ctypes.data, buf, len(buf))
        assert rc == status[i], i
        assert buf.value == bytes(text[i]), i
    kinds = {bytes(t).split(b"\n")[0] for t in text if len(bytes(t))}
    # CheckDepth's Num2Bits(9), exceed_depth, Num2Bits(65), and the embedded compression's Bits34 / ToBits
    assert {b"Error in template Num2Bits_2 line: 38", b"Error in template Blake3NovaTreePath_CheckDepth_5 line: 38",
            b"Error in template Num2Bits_11 line: 38", b"Error in template ToBits_16 line: 153"} <= kinds


@pytest.mark.parametrize("variant,cid", NOVA)
def test_nova_assert_trace_against_reference_wasm_live(built, variant, cid):
    if not ref_wasm.available(variant):
        pytest.skip("oracle/_ref not shipped")
    from oracle import port
    ref = ref_wasm.RefWasm(variant)
    p = ref.prime
    rnd = random.Random(77 + cid)
    buf = C.create_string_buffer(1024)
    n_assert = 0
    for it in range(150):
        leaf = rnd.randrange(1, 65)
        nb = rnd.randrange(1, 17)
        v = [nb, rnd.randrange(nb)] + [rnd.randrange(2**32) for _ in range(8)] + [rnd.randrange(2**32), rnd.randrange(2**32), leaf, leaf,
             rnd.randrange(leaf)] + [rnd.randrange(2**32) for _ in range(16)] + [rnd.randrange(65)]
        X = rnd.randrange(p)
        mode = it % 6
        if mode == 0:
            v[14], v[12], v[13] = X, (X + rnd.choice([0, 1, 2, 256, 257, 300])) % p, (X + rnd.randrange(70)) % p
        elif mode == 1:
            v[10], v[11] = rnd.choice([(2**64 + 5, 0), ((p - 3 * 2**32) % p, 3), (7, 2**33), (X, 0)])
        elif mode == 2:
            v[2 + rnd.randrange(8)] = rnd.choice([2**32 + 5, X, p - 1])
        elif mode == 3:
            v[15 + rnd.randrange(16)] = rnd.choice([2**32 + 5, X, p - 1, 2**34 - 1, p - 2**31])
        elif mode == 4:
            v[rnd.choice([0, 1, 31])] = rnd.choice([2**32, X, p - 2])
        else:
            for k in rnd.sample(range(32), 3):
                v[k] = rnd.choice([X, 2**32 + 1, p - 2])
        frb = np.frombuffer(b"".join(int(x % p).to_bytes(32, "little") for x in v), np.uint8).copy()
        rc = pkg.lib().b3w_assert_trace_fr(cid, frb.ctypes.data, buf, len(buf))
        rcb, _ = port.witness_fr(variant, [x % p for x in v])           # Oracle B decides quickly whether it asserts at all
        assert rc == rcb, (it, v)
        if rcb == 0:
            continue
        n_assert += 1
        rca, _ = ref.calculate({"n_blocks": v[0], "block_count": v[1], "h": v[2:10], "chunk_idx_low": v[10], "chunk_idx_high": v[11],
                                "leaf_depth": v[12], "total_depth": v[13], "depth": v[14], "m": v[15:31], "b": v[31]})
        assert (rca, ref.err_msg()) == (rc, buf.value.decode()), (it, v)
    assert n_assert >= 30


# ---- properties (hypothesis) ---------------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402

_word = st.one_of(st.integers(0, 2**32 - 1), st.integers(2**32, 2**34 + 5), st.integers(1, 2**33 + 5).map(lambda k: P - k),
                  st.integers(0, 2**256 - 1))


@settings(max_examples=200, deadline=None)
@given(st.lists(_word, min_size=16, max_size=16), st.integers(0, 15))
def test_conversion_reconstructs_every_message_word(built_once, m, j):
    """ext * 2^32 + lo == m (mod p) for every instance the conversion lets through; refused ones are outside the window"""
    v = list(range(100, 128))
    v[8:24] = m
    rows, ext, nw = convert([v])
    if ext[0, 0] == _lib.B3W_EXT_ASSERT:
        assert any(not (x % P < 2**34 or P - (x % P) <= 2**33) for x in m)
        return
    for k in range(16):
        assert (int(ext[0, k]) * 2**32 + int(rows[0, 8 + k])) % P == m[k] % P
        assert -2 <= int(ext[0, k]) <= 3
    assert nw == (1 if any(int(e) for e in ext[0]) else 0)


@settings(max_examples=150, deadline=None)
@given(st.binary(min_size=32 * 32, max_size=32 * 32), st.integers(0, 3))
def test_assert_trace_total_on_arbitrary_bytes(built_once, blob, cid):
    """any 256-bit values (also >= p) are inputs: the replay answers 0 or 4 with a well-formed text, never anything else"""
    buf = C.create_string_buffer(1024)
    data = np.frombuffer(blob, np.uint8).copy()
    rc = pkg.lib().b3w_assert_trace_fr(cid, data.ctypes.data, buf, len(buf))
    assert rc in (0, 4)
    txt = buf.value.decode()
    assert (txt == "") == (rc == 0)
    if rc == 4:
        assert txt.endswith("\n") and all(line.startswith("Error in template ") for line in txt.strip().split("\n"))
        assert ("Blake3Nova_54" in txt) == (cid != 0)


@pytest.fixture(scope="module")
def built_once(built):
    return built
