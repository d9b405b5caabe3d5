"""CPU tests: pin the oracles against the reference's golden vectors and against each other."""
import hashlib

import numpy as np
import pytest

from oracle import blake3_ref, port, ref_wasm
from hot_proofs_blake3_circom_b200 import inputs as gen
from conftest import checksum_np

WS = 24093
needs_ref = pytest.mark.skipif(not ref_wasm.available("compression"), reason="oracle/_ref not built (needs /root/reference)")


def test_golden_fixture_is_the_reference_file(golden):
    # md5 quoted in SURVEY.md section 4 for build/blake3_compression/testInp/witness.wtns
    assert hashlib.md5(golden["wtns"].tobytes()).hexdigest() == "68ae2c223a0d4a55c3013776009da705"
    assert golden["wtns"].size == 76 + WS * 32


def test_golden_input_is_lcg_6429(golden):
    # test/witness_gen.test.ts:26,36 -> genRandomChunk(new LCG(6429))
    c = blake3_ref.gen_random_chunk(blake3_ref.LCG(6429))
    assert list(golden["row"]) == c["h"] + c["m"] + c["t"] + [c["b"], c["d"]]
    assert (gen.lcg_compression_inputs(1)[0] == golden["row"]).all()


def test_port_reproduces_golden_witness(golden, built):
    w = port.witness_batch("compression", golden["row"][None, :], nthreads=1)
    assert w[0].tobytes() == golden["wtns"].tobytes()[76:]
    # public.json = main.out[0..15] = witness slots 1..16
    pub = w[0].view(np.uint32).reshape(WS, 8)[1:17, 0]
    assert (pub == golden["public"]).all()


@needs_ref
def test_reference_wasm_reproduces_golden_witness(golden):
    ref = ref_wasm.RefWasm("compression")
    assert (ref.version, ref.n32, ref.witness_size, ref.input_size) == (2, 8, WS, 28)
    row = [int(x) for x in golden["row"]]
    rc, w = ref.calculate({"h": row[0:8], "m": row[8:24], "t": row[24:26], "b": row[26], "d": row[27]})
    assert rc == 0
    assert w.tobytes() == golden["wtns"].tobytes()[76:]


def test_port_matches_reference_cases(cases, built):
    w = port.witness_batch("compression", cases["rows"], nthreads=4)
    assert np.array_equal(w, cases["witness"])


def test_kat_out_words_match_plain_blake3(built):
    # test/blake3_hash.test.ts:30-59: one default chunk, then 5 randomised (b, d=3, t0, t1) chunks, one shared LCG
    lcg = blake3_ref.LCG(6429)
    chunks = [blake3_ref.gen_random_chunk(lcg)]
    for _ in range(5):
        b = (lcg.next() % 16) * 4
        t0, t1 = lcg.next(), lcg.next()
        chunks.append(blake3_ref.gen_random_chunk(lcg, b, 3, t0, t1))
    rows = np.array([c["h"] + c["m"] + c["t"] + [c["b"], c["d"]] for c in chunks], np.uint32)
    w = port.witness_batch("compression", rows, nthreads=2).view(np.uint32).reshape(len(chunks), WS, 8)
    for c, wi in zip(chunks, w):
        want = blake3_ref.compress(c["h"], c["m"], c["t"][0], c["t"][1], c["b"], c["d"])
        assert list(wi[1:17, 0]) == want
        assert not wi[1:17, 1:].any()


def test_plain_blake3_matches_blake3_package():
    # one-block message hashed with the blake3 package == compress(IV, m, 0, 0, len, CHUNK_START|CHUNK_END|ROOT)
    blake3 = pytest.importorskip("blake3")
    msg = bytes(range(64))
    m = list(np.frombuffer(msg, "<u4"))
    out = blake3_ref.compress(blake3_ref.IV, [int(x) for x in m], 0, 0, 64, 1 | 2 | 8)
    assert b"".join(int(x).to_bytes(4, "little") for x in out[:8]) == blake3.blake3(msg).digest()


@needs_ref
def test_port_matches_reference_wasm_random(built):
    rows = np.concatenate([gen.lcg_compression_inputs(3, first=1000), gen.splitmix_compression_inputs(5, first=12345)])
    ref = ref_wasm.RefWasm("compression")
    want, st, _ = ref.batch_u32(rows, nthreads=4)
    assert (st == 0).all()
    assert np.array_equal(port.witness_batch("compression", rows, nthreads=4), want)


@needs_ref
def test_port_field_semantics_outside_u32(built):
    # inputs that are legal field elements but not u32: the port computes in the field like the wasm
    ref = ref_wasm.RefWasm("compression")
    p = ref.prime
    iv = blake3_ref.IV
    for vals, want_rc in ((iv + [2 ** 32] + [5] * 15 + [0, 0, 64, 0], 0), (iv + [p - 1] + [5] * 15 + [0, 0, 64, 0], 0),
                          (iv + [7] * 16 + [0, 0, 2 ** 33, 0], 4), ([p - 5] + iv[1:] + [7] * 16 + [0, 0, 64, 0], 4)):
        rc_b, w_b = port.witness_fr("compression", vals)
        rc_a, w_a = ref.calculate({"h": vals[0:8], "m": vals[8:24], "t": vals[24:26], "b": vals[26], "d": vals[27]})
        assert rc_a == rc_b == want_rc
        if want_rc == 0:
            assert np.array_equal(w_a, w_b)
        else:
            assert "Error in template" in ref.err_msg()


def test_checksum_definitions_agree(cases, built):
    sums = port.witness_batch("compression", cases["rows"][:4], nthreads=2, want="sums")
    assert (sums == checksum_np(cases["witness"][:4], WS)).all()


def test_witness_structure_counts(golden):
    # 23 040 gadget bits + 336 carry bits + 716 words + w[0]; by value: 23 380 slots hold 0/1
    w = golden["wtns"][76:].view(np.uint64).reshape(WS, 4)
    assert not w[:, 1:].any()
    assert int((w[:, 0] <= 1).sum()) == 23380
    assert w[0, 0] == 1


@pytest.mark.parametrize("fixture,variant,blocks", [("compression_sums_2p20.npz", "compression", (0, 101, 255)),
                                                    ("compression_sums_2p24.npz", "compression", (0, 300, 4095)),
                                                    ("nova_pasta_o2_sums_2p20.npz", "nova_pasta_o2", (0, 77, 255))])
def test_sums_fixtures_are_oracle_b(built, fixture, variant, blocks):
    """tests/golden/*_sums_2p*.npz (per-block digests of Oracle B's checksums of 2^20 / 2^24 instances, what the GPU's streamed
    runs and bench.py's config 4 / 5 are held to) re-derived here for three blocks of each, first and last included."""
    import os
    import sys
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gdir)
    import make_golden_sums as mk
    g = np.load(os.path.join(gdir, fixture))
    nblk = (1 << int(g["log2_n"])) // mk.BLOCK
    assert g["block_digest"].shape == (nblk,) and int(g["block"]) == mk.BLOCK and len(bytes(g["sha256"])) == 64
    assert blocks[-1] == nblk - 1
    rows_fn = gen.splitmix_compression_inputs if variant == "compression" else gen.splitmix_nova_inputs
    ws = port.witness_size(variant)
    for b in blocks:
        rows = rows_fn(mk.BLOCK, first=b * mk.BLOCK)
        sums, status = port.witness_batch(variant, rows, want="sums+status")
        assert not status.any() and mk.block_digests(sums)[0] == g["block_digest"][b]
        if b == 0:
            assert np.array_equal(sums[:16], g["first16"])
            # the sums are the checksum of the witness bytes (conftest.checksum_np = b3w_checksum_device's definition)
            wit = port.witness_batch(variant, rows[:8])
            assert np.array_equal(checksum_np(wit, ws), sums[:8])
