"""fp_compile (r1cs_load.h), the host-side compiler of the stand-alone R1CS check, tested WITHOUT a GPU: the tables it makes
of an exported `.r1cs` file (booleanity masks, XOR runs, row tiles, virtual bits -- b3w_debug_r1cs_program) are evaluated
here with Python integers and must give the verdict of the file's own rows: no violation on Oracle B's witnesses, and on
corrupted ones the SMALLEST violated constraint index -- through the program with virtual bits where they are all valid,
through the plainly compiled program otherwise, which is the kernel's rule (kernels_r1cs_fast.cuh)."""
import ctypes as C
import os
import random
import sys

import numpy as np
import pytest

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import port, ref_wasm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
needs_ref = pytest.mark.skipif(not (ref_wasm.available("compression") and os.path.isdir("/root/reference/build")),
                               reason="the exporter needs the reference tree")

ITEM = np.dtype([("wire", "<u4"), ("meta", "<u4"), ("coef", "<i8")])
TILE = np.dtype([("item_off", "<u4"), ("row_off", "<u4"), ("nA", "<u2"), ("nB", "<u2"), ("nC", "<u2"), ("rows", "<u2")])
XOR = np.dtype([("x", "<u4"), ("y", "<u4"), ("o", "<u4"), ("len_id", "<u4")])
SECTIONS = {"bool_mask": (0, "<u4"), "bool_row": (1, "<u4"), "xors": (2, XOR), "xor_ids": (3, "<u4"), "tiles": (4, TILE),
            "vtiles": (5, TILE), "items": (6, ITEM), "row_ids": (7, "<u4"), "taken": (8, "u1")}


def program(blob, prime, ws, plain):
    L = pkg.lib()
    pb = prime.to_bytes(32, "little")
    out = {}
    for name, (sec, dt) in SECTIONS.items():
        n = C.c_size_t()
        assert L.b3w_debug_r1cs_program(blob, len(blob), pb, ws, int(plain), sec, None, 0, C.byref(n)) == 0, L.b3w_last_error()
        buf = np.zeros(max(n.value, 1), np.uint8)
        assert L.b3w_debug_r1cs_program(blob, len(blob), pb, ws, int(plain), sec, buf.ctypes.data, n.value, C.byref(n)) == 0
        out[name] = buf[:n.value].view(dt)
    return out


def item_value(it, w):
    ln, shift = int(it["meta"]) & 63, (int(it["meta"]) >> 8) & 255
    wire = int(it["wire"])
    v = sum(w[wire + j] << j for j in range(ln)) if ln else w[wire]
    return (int(it["coef"]) * v) << shift


def run_program(P, w, p, ws):
    """-> smallest violated row id, None when every row holds, or "fallback" when a virtual bit is not a bit"""
    vbase = ((ws + 31) // 32) * 32
    w = list(w) + [0] * (vbase - ws + 32 * len(P["vtiles"]))
    items = P["items"]
    if len(P["vtiles"]) and w[0] != 1:
        return "fallback"
    for g, t in enumerate(P["vtiles"]):
        for l in range(int(t["rows"])):
            at = lambda k: items[int(t["item_off"]) + k * 32 + l]
            L = sum(item_value(at(k), w) for k in range(int(t["nA"]))) % p
            u = at(int(t["nA"]))
            unit = (int(u["coef"]) << ((int(u["meta"]) >> 8) & 255)) % p
            if L == 0:
                bit = 0
            elif L == unit:
                bit = 1
            else:
                return "fallback"
            w[vbase + 32 * g + l] = bit
    bad = []
    mask = P["bool_mask"]
    for wd in np.nonzero(mask)[0]:
        for b in range(32):
            if (int(mask[wd]) >> b) & 1:
                s = int(wd) * 32 + b
                if (w[s] * (w[s] - w[0])) % p:
                    bad.append(int(P["bool_row"][s]))
    for e in P["xors"]:
        for j in range(int(e["len_id"]) & 63):
            x, y, o = w[int(e["x"]) + j], w[int(e["y"]) + j], w[int(e["o"]) + j]
            if (2 * x * y - x - y + o) % p:
                bad.append(int(P["xor_ids"][(int(e["len_id"]) >> 6) + j]))
    for t in P["tiles"]:
        nA, nB, nC = int(t["nA"]), int(t["nB"]), int(t["nC"])
        for l in range(int(t["rows"]) & 63):
            val = lambda k0, n: sum(item_value(items[int(t["item_off"]) + (k0 + k) * 32 + l], w) for k in range(n))
            A, B, Cc = val(0, nA), val(nA, nB), val(nA + nB, nC)
            if ((A * B - Cc) if (nA and nB) else Cc) % p:
                bad.append(int(P["row_ids"][int(t["row_off"]) + l]))
    return min(bad) if bad else None


def first_violated(rows, w, p):
    for i, (A, B, Cc) in enumerate(rows):
        la = sum(co * w[k] for k, co in A.items())
        lb = sum(co * w[k] for k, co in B.items())
        lc = sum(co * w[k] for k, co in Cc.items())
        if (la * lb - lc) % p:
            return i
    return None


@needs_ref
@pytest.mark.parametrize("variant,rows_fn,n_virtual", [("compression", gen.splitmix_compression_inputs, 0),
                                                      ("nova_pasta_o2", gen.splitmix_nova_inputs, 701),
                                                      ("nova_bn_o1", gen.splitmix_nova_inputs, 0)])
def test_compiled_program_equals_the_files_rows(tmp_path, variant, rows_fn, n_virtual):
    import export_r1cs as ex
    path, _ = ex.export(variant, str(tmp_path), trials=1, verbose=False)
    blob = open(path, "rb").read()
    r = ex.read_r1cs(path)
    p, ws, rows = r["prime"], r["n_wires"], r["rows"]
    P, P0 = program(blob, p, ws, plain=False), program(blob, p, ws, plain=True)
    assert P["taken"].all() and np.array_equal(P["taken"], P0["taken"])          # the exported systems compile completely
    assert sum(int(t["rows"]) for t in P["vtiles"]) == n_virtual and len(P0["vtiles"]) == 0
    covered = int(np.unpackbits(P["bool_mask"].view(np.uint8)).sum()) + sum(int(e["len_id"]) & 63 for e in P["xors"]) + \
        sum(int(t["rows"]) & 63 for t in P["tiles"]) + n_virtual
    assert covered == len(rows)                                                  # every row exactly once

    def verdict(w):
        got = run_program(P, w, p, ws)
        return run_program(P0, w, p, ws) if got == "fallback" else got

    wit = port.witness_batch(variant, rows_fn(3, first=5))
    rng = random.Random(variant)
    n_fallback = 0
    for k in range(3):
        body = wit[k].tobytes()
        w = [int.from_bytes(body[32 * i:32 * i + 32], "little") for i in range(ws)]
        assert run_program(P, w, p, ws) is None and run_program(P0, w, p, ws) is None and first_violated(rows, w, p) is None
        slots = [0, 1, ws - 1] + [rng.randrange(ws) for _ in range(9)] if k == 0 else [rng.randrange(ws) for _ in range(8)]
        if n_virtual and k == 1:                                                 # wires inside virtual-bit definitions: the word and bits of a run
            t = P["vtiles"][rng.randrange(len(P["vtiles"]))]
            for kk in range(int(t["nA"])):
                it = P["items"][int(t["item_off"]) + kk * 32 + rng.randrange(int(t["rows"]))]
                if int(it["coef"]):
                    slots.append(int(it["wire"]) + rng.randrange(max(int(it["meta"]) & 63, 1)))
        for s in slots:
            for new in ((w[s] + 1) % p, rng.randrange(p)):
                w2 = list(w)
                w2[s] = new
                want = first_violated(rows, w2, p)
                assert want is not None
                n_fallback += run_program(P, w2, p, ws) == "fallback"
                assert verdict(w2) == want, (variant, k, s, new)
    if n_virtual:
        assert n_fallback > 0                                                    # some corruption broke a virtual bit: both programs were exercised


@pytest.mark.parametrize("circuit,variant,rows_fn,non_bits", [(0, "compression", gen.splitmix_compression_inputs, 717),
                                                               (1, "nova_bn_o2", gen.splitmix_nova_inputs, None),
                                                               (2, "nova_pasta_o2", gen.splitmix_nova_inputs, None),
                                                               (3, "nova_bn_o1", gen.splitmix_nova_inputs, None)])
def test_side_table_layout_holds_every_witness(circuit, variant, rows_fn, non_bits):
    """The checker's streaming pass places the non-bit slots of a witness by the CIRCUIT's slot kinds (b3w_debug_side_layout =
    what install_slot_rows uploads): rank[w] is a prefix count, and no 32-slot word of any oracle witness -- random inputs,
    zeros, all-ones words -- holds more non-bit values than the layout reserves for it (such a witness would take the slow,
    still exact, 'irregular' path)."""
    L = pkg.lib()
    nw, total = C.c_uint32(), C.c_uint32()
    assert L.b3w_debug_side_layout(circuit, 3, None, 0, C.byref(nw), C.byref(total)) == 0, L.b3w_last_error()
    rank = np.zeros(nw.value, np.uint32)
    assert L.b3w_debug_side_layout(circuit, 3, rank.ctypes.data, nw.value, None, None) == 0
    assert L.b3w_debug_side_layout(circuit, 3, rank.ctypes.data, nw.value - 1, None, None) != 0          # too little room
    rows = rows_fn(48, first=5)
    rows[1] = 0
    rows[2, 8:24] = 0xFFFFFFFF                                  # message words all ones
    if variant != "compression":
        rows[1, 12], rows[1, 13] = 1, 1                         # leaf_depth = total_depth = 1, depth 0: a valid step
    wit, _, st = port.witness_batch(variant, rows, want="both")
    ws = wit.shape[1] // 32
    words = (ws + 31) // 32
    assert nw.value == words + 3 + 1 and (np.diff(rank.astype(np.int64)) >= 0).all() and (rank[words:] == rank[words]).all()
    assert total.value == (int(rank[words]) + 1) // 2 * 2 and (non_bits is None or rank[words] == non_bits)
    w = wit[st == 0].reshape(-1, ws, 32)
    assert len(w) >= 40
    isbit = (w[:, :, 1:] == 0).all(axis=2) & (w[:, :, 0] < 2)
    pad = np.ones((len(w), words * 32 - ws), bool)
    nonbit_per_word = (~np.concatenate([isbit, pad], axis=1)).reshape(len(w), words, 32).sum(axis=2)
    cap = np.diff(rank.astype(np.int64))[:words]
    assert (nonbit_per_word <= cap).all()
    assert (nonbit_per_word.max(axis=0) == cap).mean() > 0.9    # and the layout is tight: nearly every reserved entry is used by some witness


@needs_ref
@pytest.mark.parametrize("variant", ["compression", "nova_pasta_o2", "nova_bn_o1"])
def test_no_row_of_a_valid_witness_needs_the_general_fr_evaluator(variant):
    """tools/r1cs_replay.py replays the compiled program on oracle witnesses (one with a chunk index beyond 2^62, one within)
    and names the path of every tile row: plain integers, the 64 x 256-bit product for IsZero's inverse ("lone"), the compare
    against one large value on the right-hand side ("big_rhs").  None may be left to the general Fr evaluator -- two nova O1
    rows were, in every instance, until round 2's last session."""
    import r1cs_replay
    kinds = r1cs_replay.main(variant, verbose=False)
    assert sum(kinds.values()) > 500 and not [k for k in kinds if k.startswith("FR")], dict(kinds)
    if variant == "nova_bn_o1":
        assert kinds["big_rhs"] >= 2 and kinds["lone"] >= 60      # the index rows of the first witness; IsZero rows
