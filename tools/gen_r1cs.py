#!/usr/bin/env python3
"""gen_r1cs.py -- OFFLINE generator of the R1CS row tables used by the on-device satisfiability check.

The reference's .r1cs files are missing from the tree (/root/reference/.MISSING_LARGE_BLOBS), so the constraint
system is re-derived from the circom templates: tools/circuit_model.py emits one R1CS row (A, B, C) for every
`<==` / `===` of the source (signal level, before circom's simplification).  Here every signal is replaced by the
VALUE it carries (its slot descriptor, see gen_tables.py), coefficients of equal values are merged, rows that
become 0 = 0 (pure aliases) are dropped and duplicates removed.  For blake3_compression exactly 24 544 non-trivial
rows remain (23 376 quadratic + 1 168 linear) -- the O1 constraint count derived in SURVEY.md 8(a) A6.

Rows are grouped into shape classes (same nA, nB, nC and coefficients) so that the 32 lanes of a warp evaluate 32 rows
of identical shape; per class the term columns are stored term-major and run-length encoded.

Output: hot_proofs_blake3_circom_b200/csrc/r1cs_tables.h  (two row sets: COMPRESSION, NOVA).
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import circuit_model as cm  # noqa: E402
import gen_tables as gt  # noqa: E402

ONE = gt.desc_of(("C", 1))
KIND_INV = 4


def reduce_rows(model):
    b = model.b

    def term(i):
        if i == 0:
            return ONE, 1
        s = b.vals[i].s
        if s[0] == "C":
            return ONE, s[1]
        return gt.desc_of(s), 1

    def red(D):
        out = {}
        for i, c in D.items():
            k, mul = term(i)
            if c * mul:
                out[k] = out.get(k, 0) + c * mul
        return tuple(sorted((k, v) for k, v in out.items() if v))

    rows, n_nontrivial = set(), 0
    for A, B, C in b.cons:
        a, bb, c = red(A), red(B), red(C)
        if not a or not bb:
            a, bb = (), ()
            if not c:
                continue
        n_nontrivial += 1
        rows.add((a, bb, c))
    return rows, n_nontrivial


def slot_rows(model, w2s):
    """Rows in WITNESS-SLOT space for an O1 build: circom's O1 pass only merges aliases (a <== b) and drops signals
    pinned to constants, so signal -> slot follows from a union-find over the alias rows; every other row is kept with
    its signals replaced by slots (term key = slot index; slot 0 is the constant 1)."""
    b = model.b
    n = len(b.names)
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    def is_alias(A, B, C):
        if A or B or len(C) != 2 or 0 in C:
            return False
        v = list(C.values())
        return v[0] == -v[1] and abs(v[0]) == 1

    for A, B, C in b.cons:
        if is_alias(A, B, C):
            x, y = (find(k) for k in C)
            if x != y:
                parent[max(x, y)] = min(x, y)
    const_of = {}
    for A, B, C in b.cons:                      # sig - c = 0  /  sig = 0
        if not A and not B:
            sigs = [k for k in C if k != 0]
            if len(sigs) == 1 and abs(C[sigs[0]]) == 1:
                const_of[find(sigs[0])] = -C.get(0, 0) * C[sigs[0]]
    slot_of = {}
    for slot, sig in enumerate(w2s):
        slot_of.setdefault(find(int(sig)), slot)
    rows, missing = set(), 0
    for A, B, C in b.cons:
        if is_alias(A, B, C):
            continue

        def red(D):
            nonlocal missing
            out = {}
            for i, c in D.items():
                if i == 0:
                    key, mul = 0, 1
                else:
                    r = find(i)
                    if r in slot_of:
                        key, mul = slot_of[r], 1
                    elif r in const_of:
                        key, mul = 0, const_of[r]
                    else:
                        missing += 1
                        key, mul = 0, 0
                if c * mul:
                    out[key] = out.get(key, 0) + c * mul
            return tuple(sorted((k, v) for k, v in out.items() if v))
        a, bb, c = red(A), red(B), red(C)
        if not a or not bb:
            a, bb = (), ()
            if not c:
                continue
        rows.add((a, bb, c))
    return rows, missing


def check_slot_rows(rows, witness_ints, p):
    for a, bb, c in rows:
        la = sum(co * witness_ints[s] for s, co in a) % p
        lb = sum(co * witness_ints[s] for s, co in bb) % p
        lc = sum(co * witness_ints[s] for s, co in c) % p
        assert (la * lb - lc) % p == 0, (a, bb, c)


def classify_slots(rows, kind_of_slot):
    """Same grouping as classify(), for slot-space rows; kind_of_slot gives each slot's value kind (for the bounds)."""
    fake = []
    for a, bb, c in rows:
        enc = lambda part: tuple(((kind_of_slot[s] << 24) | s, co) for s, co in part)
        fake.append((enc(a), enc(bb), enc(c)))
    classes = classify(fake, one=(kind_of_slot[0] << 24) | 0)
    for cl in classes:
        cl["cols"] = [[d & 0xFFFFFF for d in col] for col in cl["cols"]]
    return classes


def check_rows(model, rows):
    """every row must hold on the model's own values (evaluated through the descriptors)."""
    b = model.b
    tmax = max(b.trace) + 1
    trace = [b.trace.get(i, 0) for i in range(tmax + 8)]
    p = b.p

    def val(d):
        return gt.expand(trace, [d], p)[0]
    for a, bb, c in rows:
        la = sum(co * val(d) for d, co in a) % p
        lb = sum(co * val(d) for d, co in bb) % p
        lc = sum(co * val(d) for d, co in c) % p
        assert (la * lb - lc) % p == 0, (a, bb, c)


def _kind(d):
    return d >> 24


def _word(d):
    return d & 0xFFFF


def _bit(d):
    return (d >> 16) & 31


def is_trace_identity(row):
    """True if the row holds for ANY content of the trace, i.e. it is an identity between values that the expansion
    derives from one and the same trace word: booleanity of an extracted bit, and  w == sum 2^i * bit_i(w)."""
    a, b, c = row
    if a and b:
        # x * (1 - x) = 0 with x an extracted bit
        if len(a) == 1 and not c and _kind(a[0][0]) == 0 and a[0][0] != ONE:
            bd = dict(b)
            x = a[0][0]
            if set(bd) == {ONE, x} and bd[ONE] * a[0][1] == -bd[x] * a[0][1]:
                return True
        return False
    # linear: substitute W32(t) -> sum_i 2^i BIT(t, i) and see whether everything cancels
    acc = {}
    for d, co in c:
        if d == ONE:
            acc[("one",)] = acc.get(("one",), 0) + co
        elif _kind(d) == 0:
            acc[(_word(d), _bit(d))] = acc.get((_word(d), _bit(d)), 0) + co
        elif _kind(d) == 1:
            for i in range(32):
                acc[(_word(d), i)] = acc.get((_word(d), i), 0) + (co << i)
        else:
            return False
    return all(v == 0 for v in acc.values())


def zero_mask_form(row):
    """A linear row over BIT / W32 / W64 values, rewritten on the bits of the trace words it mentions (W32(t) = sum 2^i
    BIT(t, i), W64(t) = W32(t) + 2^32 W32(t+1): identities of the expansion).  If what is left is a sum of bits with
    coefficients of ONE sign and no constant, the row holds exactly when every one of those bits is 0 (the sum is far
    below p, nothing wraps): returns {trace word: mask of bits that must be 0}.  None if the row has another form.
    Example: Bits34's  add1.inp = sum 2^i out_bits[i] + 2^32 u + 2^33 v  with inp = W64(s, hi), out_bits = bits of s,
    u / v = bits 0 / 1 of hi  <=>  hi & ~3 == 0."""
    a, b, c = row
    if a or b:
        return None
    acc = {}
    for d, co in c:
        if d == ONE:
            acc[("one",)] = acc.get(("one",), 0) + co
            continue
        k = _kind(d)
        if k == 0:
            acc[(_word(d), _bit(d))] = acc.get((_word(d), _bit(d)), 0) + co
        elif k in (1, 2):
            for i in range(32 * k):
                key = (_word(d) + (i >> 5), i & 31)
                acc[key] = acc.get(key, 0) + (co << i)
        else:
            return None
    acc = {k: v for k, v in acc.items() if v}
    if not acc or ("one",) in acc:
        return None
    if not (all(v > 0 for v in acc.values()) or all(v < 0 for v in acc.values())):
        return None
    if sum(abs(v) for v in acc.values()) >= 1 << 200:
        return None
    masks = {}
    for (w, i) in acc:
        masks[w] = masks.get(w, 0) | (1 << i)
    return masks


def fuse_rows(rows):
    """Trace-space row set for the fused check: identities dropped, and every complete group of 32 XOR rows
    2*x_i*y_i = x_i + y_i - o_i over the bits of three words folded into ONE word-level row
    X ^ rotr(Y, dy) == rotr(O, do)  (flag XORW; the rotation rides in the descriptor's bit field)."""
    kept, xor_groups = [], {}
    n_ident = 0
    zmask = {}
    for row in rows:
        if is_trace_identity(row):
            n_ident += 1
            continue
        zm = zero_mask_form(row)
        if zm is not None:
            zmask[row] = zm
            continue
        a, b, c = row
        # IsZero:  in * inv = 1 - out  with in = S64(t) and inv = INV(t) of the SAME trace word: the product is 1 for
        # in != 0 and 0 for in = 0 by the definition of the INV slot, so in trace space the row says  C.z == (in != 0)
        # and needs no field arithmetic (flag ISZ; kept with A = in, no B).
        if (len(a) == 1 and len(b) == 1 and a[0][1] == 1 and b[0][1] == 1 and _kind(a[0][0]) == 3 and _kind(b[0][0]) == KIND_INV
                and _word(a[0][0]) == _word(b[0][0])):
            kept.append((a, (), c))
            continue
        if (len(a) == 1 and len(b) == 1 and len(c) == 3 and a[0][1] == 2 and b[0][1] == 1
                and all(_kind(d) == 0 and d != ONE for d, _ in a + b + c)):
            cd = dict(c)
            x, y = a[0][0], b[0][0]
            o = [d for d in cd if d not in (x, y)]
            if len(o) == 1 and cd.get(x) == 1 and cd.get(y) == 1 and cd[o[0]] == -1 and x != y:
                xor_groups.setdefault((_word(x), _word(y), _word(o[0])), []).append((row, _bit(x), _bit(y), _bit(o[0])))
                continue
        kept.append(row)
    xorw = []
    for (tx, ty, to), members in xor_groups.items():
        dy = {(ky - kx) % 32 for _, kx, ky, ko in members}
        do = {(ko - kx) % 32 for _, kx, ky, ko in members}
        if len(members) == 32 and len({kx for _, kx, _, _ in members}) == 32 and len(dy) == 1 and len(do) == 1:
            xorw.append(((1 << 24) | tx, (1 << 24) | (dy.pop() << 16) | ty, (1 << 24) | (do.pop() << 16) | to))
        else:
            kept.extend(m[0] for m in members)
    return kept, xorw, n_ident, zmask


def check_xorw(model, xorw):
    b = model.b
    rot = lambda w, r: ((w >> r) | (w << (32 - r))) & 0xFFFFFFFF
    for x, y, o in xorw:
        X, Y, O = b.trace[_word(x)], b.trace[_word(y)], b.trace[_word(o)]
        assert X ^ rot(Y, _bit(y)) == rot(O, _bit(o)), (x, y, o)


def check_zmask(model, zmask):
    b = model.b
    for row, masks in zmask.items():
        for w, m in masks.items():
            assert b.trace.get(w, 0) & m == 0, (row, w, m)


MERGE_BELOW = 16        # shape classes with fewer rows than this are merged into per-row-coefficient classes
FLAG_FIELD, FLAG_WIDE, FLAG_XORW, FLAG_ROWCOEF, FLAG_ZMASK, FLAG_ISZ = 1, 2, 4, 8, 16, 32
_BOUND = {0: 1, 1: (1 << 32) - 1, 2: (1 << 64) - 1, 3: 1 << 63, 4: 1 << 255}


def _is_wide(nA, nB, nC, coefs, ones, kinds):
    """value bounds per kind decide whether 64-bit signed arithmetic is exact for a row of this shape"""
    mx = lambda lo, hi: sum(abs(coefs[t]) * (1 if ones[t] else _BOUND[kinds[t]]) for t in range(lo, hi))
    la, lb, lc = mx(0, nA), mx(nA, nA + nB), mx(nA + nB, nA + nB + nC)
    return la >= 1 << 62 or lb >= 1 << 62 or lc >= 1 << 62 or la * lb >= 1 << 62


def classify(rows, xorw=(), one=ONE, zmask=None):
    """-> list of classes: dict(nA, nB, nC, flags, coefs, cols) with cols[t] = list of descs (one per row).
    Rows of identical shape AND coefficients share one coefficient vector; the many small groups that remain (rows whose
    coefficients are all different: IV[i] * is_parent, 2^k recompositions split by circom, ...) are merged by (nA, nB,
    nC) into classes whose coefficients are stored per row (flag ROWCOEF, coefs term-major), so that a warp still
    evaluates 32 of them per step instead of one."""
    groups = {}
    for a, bb, c in rows:
        # order terms inside each part by coefficient, then kind, so that equal shapes line up; ONE first
        def key(t):
            d, co = t
            return (d != one, co, d >> 24, d)
        a2, b2, c2 = sorted(a, key=key), sorted(bb, key=key), sorted(c, key=key)
        field = any((d >> 24) == KIND_INV for d, _ in a2 + b2 + c2)
        if len(a2) == 1 and not b2:
            field = "isz"                # folded IsZero row (see fuse_rows)
        shape = (len(a2), len(b2), len(c2), tuple(co for _, co in a2 + b2 + c2), tuple(d == one for d, _ in a2 + b2 + c2), field,
                 tuple(d >> 24 for d, _ in a2 + b2 + c2))
        groups.setdefault(shape, []).append([d for d, _ in a2 + b2 + c2])
    classes = []
    merged = {}
    for shape, members in sorted(groups.items(), key=lambda kv: (-len(kv[1]), kv[0])):
        nA, nB, nC, coefs, ones, field, kinds = shape
        if field == "isz":
            assert nA == 1 and nB == 0 and coefs[0] == 1, shape
        elif field:
            assert nA == 1 and nB == 1 and coefs[0] == 1 and coefs[1] == 1, shape
        assert all(abs(x) < (1 << 100) for x in coefs)
        wide = not field and _is_wide(nA, nB, nC, coefs, ones, kinds)
        fflag = FLAG_ISZ if field == "isz" else (FLAG_FIELD if field else 0)
        if len(members) < MERGE_BELOW:
            g = merged.setdefault((nA, nB, nC, field), dict(rows=[], wide=False))
            g["rows"].extend((m, coefs) for m in members)
            g["wide"] = g["wide"] or wide
            continue
        members.sort(key=lambda m: [((d >> 24), d & 0xFFFF, (d >> 16) & 31) for d in m])   # (kind, word, bit): long runs
        cols = [[m[t] for m in members] for t in range(nA + nB + nC)]
        classes.append(dict(nA=nA, nB=nB, nC=nC, flags=fflag | (FLAG_WIDE if wide else 0), coefs=coefs,
                            cols=cols, count=len(members)))
    for (nA, nB, nC, field), g in sorted(merged.items()):
        rws = sorted(g["rows"], key=lambda mc: [((d >> 24), d & 0xFFFF, (d >> 16) & 31) for d in mc[0]])
        nt = nA + nB + nC
        cols = [[m[t] for m, _ in rws] for t in range(nt)]
        coefs = tuple(co[t] for t in range(nt) for _, co in rws)          # term-major, one per row
        classes.append(dict(nA=nA, nB=nB, nC=nC, flags=FLAG_ROWCOEF | (FLAG_ISZ if field == "isz" else (FLAG_FIELD if field else 0)) | (FLAG_WIDE if g["wide"] else 0),
                            coefs=coefs, cols=cols, count=len(rws)))
    if zmask:
        # one class row per (trace word, mask): column 0 = the word (as a W32 descriptor), column 1 = the mask itself
        members = sorted({(w, m) for masks in zmask.values() for w, m in masks.items()})
        classes.insert(0, dict(nA=0, nB=0, nC=2, flags=FLAG_ZMASK, coefs=(1, 1), cols=[[(1 << 24) | w for w, _ in members], [m for _, m in members]],
                               count=len(members)))
    if xorw:
        members = sorted(xorw, key=lambda m: [(d & 0xFFFF) for d in m])
        classes.insert(0, dict(nA=1, nB=1, nC=1, flags=FLAG_XORW, coefs=(1, 1, 1), cols=[[m[t] for m in members] for t in range(3)],
                               count=len(members)))
    return classes


def emit_set(name, classes, out):
    segs, coefs, cls = [], [], []
    for c in classes:
        seg_off = len(segs)
        for col in c["cols"]:
            r = gt.rle(col)
            segs.append(("COL", len(r)))
            segs.extend(r)
        cls.append((c["nA"], c["nB"], c["nC"], c["flags"], c["count"], len(coefs), seg_off))
        coefs.extend(c["coefs"])
    out.append("static const b3w_r1cs_class R1CS_CLASSES_%s[%d] = {" % (name, len(cls)))
    line = "  "
    for x in cls:
        item = "{%d,%d,%d,%d,%d,%d,%d}," % x
        if len(line) + len(item) > 118:
            out.append(line)
            line = "  "
        line += item
    out.append(line)
    out.append("};")
    # signed 128-bit coefficients as {low 64 bits, high 64 bits}
    out.append("static const uint64_t R1CS_COEFS_%s[%d][2] = {" % (name, len(coefs)))
    line = "  "
    for x in coefs:
        u = x & ((1 << 128) - 1)
        item = "{0x%xull,0x%xull}," % (u & ((1 << 64) - 1), u >> 64)
        if len(line) + len(item) > 118:
            out.append(line)
            line = "  "
        line += item
    out.append(line)
    out.append("};")
    # column streams: a {0xFFFFFFFF, n, 0} header announces n run records
    out.append("static const b3w_seg R1CS_COLS_%s[%d] = {" % (name, len(segs)))
    line = "  "
    for x in segs:
        item = "{0xffffffff,%d,0}," % x[1] if x[0] == "COL" else "{0x%x,%d,%d}," % x
        if len(line) + len(item) > 118:
            out.append(line)
            line = "  "
        line += item
    out.append(line)
    out.append("};")
    out.append("")


def build(check_only=False, verbose=True):
    rng = random.Random(77)
    sets = []
    for name in ("COMPRESSION", "NOVA"):
        models = []
        for trial in range(3):
            if name == "COMPRESSION":
                models.append(cm.CompressionModel(gt.random_inputs("compression", rng, edge=trial)))
            else:
                models.append(cm.NovaModel(cm.random_nova_inputs(rng, edge=(0, 4, 5)[trial])))
        rows, n_nontrivial = reduce_rows(models[0])
        for m in models[1:]:
            r2, _ = reduce_rows(m)
            assert r2 == rows, "constraint structure depends on the input"
        for m in models:
            check_rows(m, rows)
        classes = classify(rows)
        nnz = sum(len(a) + len(b) + len(c) for a, b, c in rows)
        if verbose:
            print("%-18s %6d non-trivial rows (%d unique; %d quadratic), %d classes, %d terms"
                  % (name, n_nontrivial, len(rows), sum(1 for r in rows if r[0]), len(classes), nnz))
        sets.append((name, classes, len(rows), nnz))
        kept, xorw, n_ident, zmask = fuse_rows(rows)
        for m in models:
            check_xorw(m, xorw)
            check_zmask(m, zmask)
        fclasses = classify(kept, xorw, zmask=zmask)
        n_zm = fclasses[1]["count"] if zmask else 0
        fnnz = sum(len(a) + len(b) + len(c) for a, b, c in kept) + 3 * len(xorw) + 2 * n_zm
        assert n_ident + 32 * len(xorw) + len(kept) + len(zmask) == len(rows)
        if verbose:
            print("%-18s %6d rows evaluated in trace space: %d identities dropped, %d XOR rows folded into %d word rows, "
                  "%d bit-recomposition rows folded into %d zero-mask rows, %d others; %d classes, %d terms"
                  % (name + "_FUSED", len(kept) + len(xorw) + n_zm, n_ident, 32 * len(xorw), len(xorw), len(zmask), n_zm,
                     len(kept), len(fclasses), fnnz))
        sets.append((name + "_FUSED", fclasses, len(kept) + len(xorw) + n_zm, fnnz))
    # slot-space sets for the stand-alone check of witnesses in HBM (O1 builds only)
    from oracle.ref_wasm import RefWasm
    for variant, setname in (("compression", "COMPRESSION_SLOTS"), ("nova_bn_o1", "NOVA_BN_O1_SLOTS")):
        ref = RefWasm(variant)
        import numpy as np
        w2s = np.frombuffer(ref.memory(gt.W2S_OFFSET[variant], 4 * ref.witness_size), np.uint32)
        srows = None
        for trial in range(3):
            if variant == "compression":
                inputs = gt.random_inputs("compression", rng, edge=trial)
                mdl = cm.CompressionModel(inputs)
            else:
                inputs = cm.random_nova_inputs(rng, edge=(0, 4, 5)[trial])
                mdl = cm.NovaModel(inputs, o1=True)
            r, missing = slot_rows(mdl, w2s)
            assert missing == 0, "%s: %d terms refer to signals that are not in the witness" % (variant, missing)
            assert srows is None or r == srows
            srows = r
            d, pos = {}, 0
            for nm, sz in ref.plan:
                d[nm] = inputs[pos:pos + sz]
                pos += sz
            rc, wit = ref.calculate(d)
            assert rc == 0
            wi = [int.from_bytes(wit[32 * i:32 * i + 32].tobytes(), "little") for i in range(ref.witness_size)]
            check_slot_rows(srows, wi, ref.prime)          # the reference's own witness satisfies every row
        kinds = [gt.desc_of(mdl.b.vals[int(sg)].s) >> 24 for sg in w2s]
        sclasses = classify_slots(srows, kinds)
        snnz = sum(len(a) + len(b) + len(c) for a, b, c in srows)
        if verbose:
            print("%-18s %6d rows in witness-slot space (%d quadratic), %d classes, %d terms; satisfied by the reference's witnesses"
                  % (setname, len(srows), sum(1 for r in srows if r[0]), len(sclasses), snnz))
        sets.append((setname, sclasses, len(srows), snnz))
    # O2 builds (blake3_nova_js, blake3_nova_pasta_js): circom's O2 pass removed one signal per linear constraint, so the
    # template-level rows above no longer fit their witnesses.  tools/export_r1cs.py re-derives the O2-form system by
    # solving every linear row for the signal the O2 witness dropped and substituting it; the result has integer
    # coefficients (<= 65 bits) and is THE SAME for both primes -- one built-in set serves both contexts.
    import export_r1cs as ex
    import numpy as np
    o2rows = None
    for variant in ("nova_bn_o2", "nova_pasta_o2"):
        ref = RefWasm(variant)
        w2s = np.frombuffer(ref.memory(gt.W2S_OFFSET[variant], 4 * ref.witness_size), np.uint32)
        rng2 = random.Random(2024)
        for trial in range(2):
            inputs = gt.random_inputs(variant, rng2, edge=0)
            mdl = gt.model_for(variant, inputs, ref.prime)
            rows_d, _ = ex.derive(mdl, w2s)
            r = {(tuple(sorted(A.items())), tuple(sorted(B.items())), tuple(sorted(C.items()))) for A, B, C in rows_d}
            assert o2rows is None or r == o2rows, "the O2-form system depends on the input or the prime"
            o2rows = r
            d, pos = {}, 0
            for nm, sz in ref.plan:
                d[nm] = inputs[pos:pos + sz]
                pos += sz
            rc, wit = ref.calculate(d)
            assert rc == 0
            wi = [int.from_bytes(wit[32 * i:32 * i + 32].tobytes(), "little") for i in range(ref.witness_size)]
            check_slot_rows(o2rows, wi, ref.prime)         # the reference's own witness satisfies every row
    o2classes = classify_slots(o2rows, [0] * ref.witness_size)   # value kinds are not needed: the stand-alone checker classifies at run time
    o2nnz = sum(len(a) + len(b) + len(c) for a, b, c in o2rows)
    if verbose:
        print("%-18s %6d rows in witness-slot space (%d quadratic), %d classes, %d terms; satisfied by the reference's witnesses (both primes)"
              % ("NOVA_O2_SLOTS", len(o2rows), sum(1 for r in o2rows if r[0]), len(o2classes), o2nnz))
    sets.append(("NOVA_O2_SLOTS", o2classes, len(o2rows), o2nnz))
    out = ["/* GENERATED by tools/gen_r1cs.py -- do not edit.",
           " * R1CS rows re-derived from the circom templates (the reference's .r1cs files are absent).",
           " *   *_FUSED sets: VALUE space, a term is a slot descriptor (trace_layout.h); identities dropped, XOR rows folded.",
           " *   *_SLOTS sets: WITNESS-SLOT space, a term is a slot index (all 24 544 rows for compression; NOVA_O2_SLOTS = the O2-form",
           " *                 system of blake3_nova_js / blake3_nova_pasta_js, identical for both primes).",
           " * Grouped by shape; term columns run-length encoded. */",
           "#pragma once", "#include <stdint.h>", '#include "slot_tables.h"',
           "typedef struct { uint16_t nA, nB, nC, flags; uint32_t count, coef_off, col_off; } b3w_r1cs_class;", ""]
    for name, classes, nrows, nnz in sets:
        if name in ("COMPRESSION", "NOVA"):
            continue                    # the value-space full sets are only an intermediate (counts printed above)
        out.append("#define R1CS_ROWS_%s %d" % (name, nrows))
        out.append("#define R1CS_TERMS_%s %d" % (name, nnz))
        emit_set(name, classes, out)
    text = "\n".join(out) + "\n"
    path = os.path.join(ROOT, "hot_proofs_blake3_circom_b200", "csrc", "r1cs_tables.h")
    if check_only:
        same = os.path.exists(path) and open(path).read() == text
        print("%s: %s" % (os.path.relpath(path, ROOT), "up to date" if same else "DIFFERS"))
        return same
    with open(path, "w") as f:
        f.write(text)
    print("wrote %s (%d bytes)" % (os.path.relpath(path, ROOT), len(text)))
    return True


if __name__ == "__main__":
    ok = build(check_only="--check-only" in sys.argv)
    sys.exit(0 if ok else 1)
