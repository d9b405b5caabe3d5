"""CPU tests of the host layer that mirrors witness_calculator.js:131-169 (input handling and errors)."""
import numpy as np
import pytest

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import blake3_ref

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


@pytest.fixture(scope="module")
def wc(built):
    return pkg.builder("blake3_compression", lazy=True)


def good():
    return {"h": list(blake3_ref.IV), "m": list(range(16)), "t": [0, 0], "b": 64, "d": 0}


def test_constructor_fields(wc):
    assert (wc.version, wc.n32, wc.prime, wc.witnessSize) == (2, 8, P, 24093)
    assert wc.circom_version() == 2


def test_row_layout_and_value_forms(wc):
    inp = good()
    inp["m"] = [[str(i) for i in range(8)], [hex(i) for i in range(8, 16)]]     # nested arrays are flattened (:303-317)
    inp["b"] = "64"
    inp["t"] = [P, -P]                                                          # reduced mod p (:319-323)
    inp["d"] = -(P - 3)                                                         # negative -> wrapped
    row = wc._row(inp)
    assert list(row) == list(blake3_ref.IV) + list(range(16)) + [0, 0, 64, 3]


def test_error_messages_match_reference(wc):
    inp = good(); inp["m"] = list(range(15))
    with pytest.raises(RuntimeError, match="Not enough values for input signal m"):
        wc._row(inp)
    inp = good(); inp["h"] = list(range(9))
    with pytest.raises(RuntimeError, match="Too many values for input signal h"):
        wc._row(inp)
    inp = good(); inp["bogus"] = 1          # the wasm reports size 0 for unknown names
    with pytest.raises(RuntimeError, match="Too many values for input signal bogus"):
        wc._row(inp)
    inp = good(); del inp["d"]
    with pytest.raises(RuntimeError, match="Not all inputs have been set. Only 27 out of 28"):
        wc._row(inp)


def test_out_of_domain_is_loud(wc):
    inp = good(); inp["m"][0] = 2 ** 32
    with pytest.raises(pkg.B3WError) as e:
        wc._row(inp)
    assert e.value.code == _lib.B3W_ERR_DOMAIN


def test_wasm_identification():
    with pytest.raises(pkg.B3WError):
        pkg.circuit_from_wasm(b"\0asm\x01\0\0\0")
    assert len(pkg.CIRCUITS) == 4


def test_input_generators():
    a = gen.lcg_compression_inputs(5)
    lcg = blake3_ref.LCG(6429)
    for i in range(5):
        c = blake3_ref.gen_random_chunk(lcg)
        assert list(a[i]) == c["h"] + c["m"] + c["t"] + [c["b"], c["d"]]
    assert (gen.lcg_compression_inputs(3, first=2) == a[2:]).all()
    b = gen.splitmix_compression_inputs(4096)
    assert (b[:, 26] % 4 == 0).all() and b[:, 26].max() == 64 and b[:, 26].min() == 0
    assert b[:, 27].max() == 15
    nz = (np.arange(16)[None, :] >= (b[:, 26] // 4)[:, None])
    assert not b[:, 8:24][nz].any()
    assert (gen.splitmix_compression_inputs(10, first=100) == gen.splitmix_compression_inputs(110)[100:]).all()


# ---- the assert trace text of the nova circuits (witness_calculator.js:21-43), against the reference's own wasm ----
@pytest.mark.parametrize("variant,cid", [("nova_bn_o2", 1), ("nova_pasta_o2", 2), ("nova_bn_o1", 3)])
def test_assert_trace_matches_reference_wasm(built, variant, cid):
    import ctypes as C
    from oracle import ref_wasm
    if not ref_wasm.available(variant):
        pytest.skip("oracle/_ref not built")
    ref = ref_wasm.RefWasm(variant)
    L = pkg.lib()
    buf = C.create_string_buffer(1024)
    seen = set()
    cases = [(3, 3), (3, 300), (600, 1), (0, 0), (3, 1), (257, 0), (258, 1), (256, 0), (1, 0), (64, 63), (64, 64),
             (0xFFFFFFFF, 0), (0, 0xFFFFFFFF), (300, 43), (300, 44), (5, 260), (5, 261)]
    for leaf, depth in cases:
        row = np.zeros(32, np.uint32)
        row[0], row[2:10], row[12], row[13], row[14], row[31] = 2, 1, leaf, 3, depth, 64
        rc = L.b3w_assert_trace(cid, row.ctypes.data, buf, len(buf))
        d = dict(n_blocks=2, block_count=0, h=[1] * 8, chunk_idx_low=0, chunk_idx_high=0, leaf_depth=leaf, total_depth=3,
                 depth=depth, m=[0] * 16, b=64)
        rc_ref, _ = ref.calculate(d)
        assert rc == rc_ref, (leaf, depth)
        assert buf.value.decode() == ref.err_msg(), (leaf, depth)
        seen.add(buf.value.decode())
    assert len(seen) == 4          # no assert + the three assert sites reachable with u32 inputs


def test_assert_trace_empty_for_compression(built):
    import ctypes as C
    buf = C.create_string_buffer(64)
    row = np.full(28, 0xFFFFFFFF, np.uint32)
    assert pkg.lib().b3w_assert_trace(0, row.ctypes.data, buf, len(buf)) == 0 and buf.value == b""


# ---- property tests (hypothesis) of the input normalisation: witness_calculator.js:319-323 `BigInt(n) % prime`, made
# non-negative, in every value form a JS caller can pass; and of the C-side Fr256 conversion -------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402

_u32 = st.integers(min_value=0, max_value=2 ** 32 - 1)
_forms = st.sampled_from(["int", "str", "hex", "plus_p", "minus_p", "neg_wrap", "np"])


def _encode(v, form):
    if form == "int":
        return v
    if form == "str":
        return str(v)
    if form == "hex":
        return hex(v)
    if form == "plus_p":
        return v + 3 * P
    if form == "minus_p":
        return str(v - 2 * P)            # negative: JS `%` keeps the sign, normalize() adds the prime back
    if form == "neg_wrap":
        return -(P - v) if v else 0
    return np.uint32(v)


@settings(max_examples=60, deadline=None)
@given(vals=st.lists(_u32, min_size=28, max_size=28), forms=st.lists(_forms, min_size=28, max_size=28))
def test_row_normalisation_property(wc, vals, forms):
    enc = [_encode(v, f) for v, f in zip(vals, forms)]
    inp = {"t": enc[24:26], "h": enc[0:8], "d": enc[27], "m": [enc[8:16], enc[16:24]], "b": enc[26]}   # any key order, nesting
    assert list(wc._row(inp)) == vals


@settings(max_examples=40, deadline=None)
@given(vals=st.lists(_u32, min_size=32, max_size=32), ks=st.lists(st.integers(min_value=0, max_value=3), min_size=32, max_size=32))
def test_inputs_from_fr_property(built, vals, ks):
    # nova rows under the Pallas scalar field: value + k*p for the k that still fit 256 bits
    PALLAS = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001
    fr = np.frombuffer(b"".join((v + k * PALLAS).to_bytes(32, "little") for v, k in zip(vals, ks)), np.uint8).copy()
    rows = np.zeros(32, np.uint32)
    assert pkg.lib().b3w_inputs_from_fr(2, fr.ctypes.data, 1, rows.ctypes.data) == 0
    assert list(rows) == vals


@settings(max_examples=30, deadline=None)
@given(v=st.integers(min_value=2 ** 32, max_value=P - 1), pos=st.integers(min_value=0, max_value=27))
def test_out_of_domain_values_are_always_refused(wc, v, pos):
    row = [1] * 28
    row[pos] = v
    inp = {"h": row[0:8], "m": row[8:24], "t": row[24:26], "b": row[26], "d": row[27]}
    with pytest.raises(pkg.B3WError) as e:
        wc._row(inp)
    assert e.value.code == _lib.B3W_ERR_DOMAIN
