"""r1cs_replay.py -- replay the stand-alone checker's COMPILED program on the host, on real (oracle) witnesses, and report
which rows leave the cheap paths of kernels_r1cs_fast.cuh: a tile marked FP_TILE_FAST whose 64-bit pass would fail its
run-time bound check (field-valued or signed operands), rows that need the 64 x 256-bit product ("lone"), the compare
against one large value on the right-hand side ("big_rhs"), or the general Fr evaluator ("FR").  Needs /root/reference
(the exporter) and the oracle library; no GPU.

    python tools/r1cs_replay.py [compression | nova_bn_o1 | nova_pasta_o2 | nova_bn_o2]

This is how the two nova O1 rows that took the Fr evaluator in every instance (the 64-bit chunk index) and the FAST tiles
that were evaluated twice were found."""
import os, sys
from collections import Counter
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import export_r1cs as ex
from test_r1cs_program import program
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import port

FAST = 0x8000


def main(variant, verbose=True):
    """-> Counter of the path every tile row takes (int / lone / big_rhs / FR...), over two oracle witnesses"""
    total = Counter()
    out_dir = "/tmp/r1cs_replay"
    os.makedirs(out_dir, exist_ok=True)
    path, _ = ex.export(variant, out_dir, trials=1, verbose=False)
    blob = open(path, "rb").read()
    r = ex.read_r1cs(path)
    p, ws = r["prime"], r["n_wires"]
    P = program(blob, p, ws, plain=False)
    rows_in = gen.splitmix_compression_inputs(4, first=3) if variant == "compression" else gen.splitmix_nova_inputs(4, first=3)
    if variant != "compression":
        rows_in[0, 11] |= 0x80000000       # chunk_idx >= 2^63: beyond the tagged 8-byte values
        rows_in[1, 11] &= 0x0FFFFFFF       # ... and one that fits
    wit = port.witness_batch(variant, rows_in)

    def small(v):                      # the kernel's tagged 8-byte value: |v| < 2^62, else "BIG"
        if v < (1 << 62):
            return v
        if p - v < (1 << 62):
            return -(p - v)
        return None

    items = P["items"]

    def item(t, k, l):
        return items[int(t["item_off"]) + k * 32 + l]

    for inst in range(2):
        w = [int.from_bytes(bytes(wit[inst, 32 * s:32 * s + 32]), "little") for s in range(ws)]
        vbase = ((ws + 31) // 32) * 32
        w += [0] * (vbase - ws + 32 * len(P["vtiles"]))
        for g, t in enumerate(P["vtiles"]):                                   # virtual bits (fp_eval_virtuals)
            for l in range(int(t["rows"])):
                acc = 0
                for k in range(int(t["nA"])):
                    it = item(t, k, l)
                    ln, shift, wire = int(it["meta"]) & 63, (int(it["meta"]) >> 8) & 255, int(it["wire"])
                    acc += (int(it["coef"]) * (sum(w[wire + j] << j for j in range(ln)) if ln else w[wire])) << shift
                w[vbase + 32 * g + l] = 0 if acc % p == 0 else 1
        for ti, t in enumerate(P["tiles"]):
            fast = bool(int(t["rows"]) & FAST)
            nA, nB, nC = int(t["nA"]), int(t["nB"]), int(t["nC"])
            kinds = []
            fast_fails = False
            why = set()
            for l in range(int(t["rows"]) & 63):
                nbig, undec, big_side, bigcoef, L = 0, False, None, None, [0, 0, 0]
                for k in range(nA + nB + nC):
                    it = item(t, k, l)
                    ln, shift, cbits = int(it["meta"]) & 63, (int(it["meta"]) >> 8) & 255, (int(it["meta"]) >> 16) & 255
                    coef, side, wire = int(it["coef"]), (0 if k < nA else 1 if k < nA + nB else 2), int(it["wire"])
                    if coef == 0:
                        continue
                    if ln:
                        undec = undec or any(w[wire + j] > 1 for j in range(ln))
                        v, vb = sum(w[wire + j] << j for j in range(ln)), ln
                    else:
                        sv = small(w[wire])
                        bound = (int(it["meta"]) >> 24) & 63
                        if sv is None or sv < 0 or sv >> bound:
                            fast_fails = True
                            why.add("field" if sv is None else "negative" if sv < 0 else "over 2^%d" % bound)
                        if sv is None:
                            nbig, big_side, bigcoef = nbig + 1, side, (coef, shift)
                            continue
                        v, vb = sv, abs(sv).bit_length()
                    undec = undec or cbits + vb > 118
                    L[side] += (coef * v) << shift
                unit = bigcoef is not None and bigcoef[1] == 0 and abs(bigcoef[0]) == 1
                if nbig:
                    if not undec and nbig == 1 and unit and big_side <= 1 and nA and nB and L[big_side] == 0 \
                            and abs(L[1 - big_side]).bit_length() <= 62 and abs(L[2]).bit_length() <= 62:
                        kinds.append("lone")
                    elif not undec and nbig == 1 and unit and big_side == 2 and \
                            (not (nA and nB) or abs(L[0]).bit_length() + abs(L[1]).bit_length() <= 125):
                        kinds.append("big_rhs")
                    else:
                        kinds.append("FR(nbig=%d side=%s coef=%s)" % (nbig, big_side, bigcoef))
                elif undec or (nA and nB and abs(L[0]).bit_length() + abs(L[1]).bit_length() > 125):
                    kinds.append("FR")
                else:
                    kinds.append("int")
            c = Counter(kinds)
            total.update(c)
            if verbose and ((fast and fast_fails) or not fast or any(k != "int" for k in c)):
                print("instance %d tile %2d %-5s A,B,C items %d,%d,%d  %s%s" % (inst, ti, "FAST" if fast else "exact", nA, nB, nC, dict(c),
                                                                               "   <- FAST tile evaluated twice (%s)" % ", ".join(sorted(why)) if fast and fast_fails else ""))
    if verbose:
        print("(the debug dump compiles WITHOUT the circuit's slot-kind hint: install_slot_rows compiles the tiles flagged 'evaluated twice' for the exact path)")
    return total


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "nova_bn_o1")
