"""circuit_model.py -- offline model of the reference circuits used to GENERATE the kernel's slot tables.

This is a build-time tool, not the product path and not the oracle.  It re-derives
  (1) circom 2.1.6's signal numbering (per component: outputs, inputs, intermediates in declaration
      order, then sub-components sorted by name / array index, depth first; signal 0 is the constant 1), and
  (2) for every signal, WHERE its value lives in the compact per-instance "trace" that the CUDA kernel
      computes (hot_proofs_blake3_circom_b200/csrc/blake3wit.cu), as a symbolic source:
          ("C", c)        small constant c (0 or 1 only ever reach a witness slot)
          ("W", t)        the 32-bit trace word t
          ("Q", t)        the 64-bit value trace[t] | trace[t+1] << 32
          ("B", t, k)     bit k of trace word t
          ("S", t)        the signed 64-bit integer trace[t] | trace[t+1] << 32, as a field element (x mod p)
          ("I", t)        the inverse mod p of that signed integer (0 for 0)   [circomlib IsZero.inv, nova only]
Every template below mirrors one template of the reference; the file:line it follows is cited
(paths under /root/reference).  The model also carries concrete integer values so that the numbering
and semantics can be validated signal-by-signal against the reference wasm's memory (tests do this
through tools/gen_tables.py --check, which needs /root/reference; the product only needs the
generated tables).

Trace layout (u32 words), shared with the kernel -- keep in sync with csrc/trace_layout.h:
    TR_ZERO = 0, TR_ONE = 1
    TR_IN   = 2   .. 30   compression inputs h[8] m[16] t[2] b d (circuit declaration order)
    TR_OUT  = 30  .. 46   out[16]
    (46, 47 unused: pad to a 16-byte boundary)
    TR_HG   = 48  .. 944  112 half-G records of 8 words, index ((round*8 + g)*2 + half):
                          +0 a' = low 32 bits of a+b+xy     +1 carries of that sum (bit0 = u, bit1 = v)
                          +2 d  (before)                    +3 d' = rotr(d ^ a', R1)
                          +4 c' = low 32 bits of c+d'       +5 carry of that sum (bit0 = u)
                          +6 b  (before)                    +7 b' = rotr(b ^ c', R2)
    TR_NOVA = 944 ..      nova-only words (see NovaModel)
"""

TR_ZERO, TR_ONE, TR_IN, TR_OUT, TR_HG, TR_NOVA = 0, 1, 2, 30, 48, 944
M32 = 0xFFFFFFFF

BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
PALLAS_SCALAR = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001

IV = [0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19]
SIGMA = [2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8]   # circuits/blake3_common.circom:20-21


def rotr(x, r):
    return ((x >> r) | (x << (32 - r))) & M32


class V:
    """A value flowing through the model: concrete integer + symbolic source."""
    __slots__ = ("v", "s")

    def __init__(self, v, s):
        self.v, self.s = v, s

    def bit(self, k):
        """bit k of a word-valued V (ToBits & friends: (inp >> k) & 1)."""
        kind = self.s[0]
        if kind == "C":
            return V((self.v >> k) & 1, ("C", (self.v >> k) & 1))
        if kind == "W":
            assert k < 32
            return V((self.v >> k) & 1, ("B", self.s[1], k))
        if kind == "Q":
            t = self.s[1] + (k >> 5)
            return V((self.v >> k) & 1, ("B", t, k & 31))
        if kind == "B":
            assert k == 0 or True
            return V((self.v >> k) & 1, self.s if k == 0 else ("C", 0))
        raise ValueError("bit of %r" % (self.s,))


def const(c):
    return V(c, ("C", c))


class Builder:
    def __init__(self, prime=BN254_R):
        self.p = prime
        self.names = ["one"]
        self.vals = [const(1)]
        self.trace = {}          # trace index -> u32 value (sparse while building)
        self.cons = []           # R1CS rows (A, B, C): dicts signal index -> integer coefficient; (A.z)*(B.z) = C.z

    # --- signal allocation ---
    def alloc(self, name, *dims):
        """Allocate a (possibly multi-dimensional) signal array; returns nested lists of indices."""
        def rec(prefix, ds):
            if not ds:
                self.names.append(prefix)
                self.vals.append(None)
                return len(self.names) - 1
            return [rec("%s[%d]" % (prefix, i), ds[1:]) for i in range(ds[0])]
        return rec(name, dims)

    def set(self, idx, val):
        assert isinstance(val, V)
        assert self.vals[idx] is None, "signal %s assigned twice" % self.names[idx]
        self.vals[idx] = val

    def get(self, idx):
        v = self.vals[idx]
        assert v is not None, "signal %s read before assignment" % self.names[idx]
        return v

    # --- trace ---
    def tw(self, t, value):
        """Define trace word t := value (u32) and return it as a word V."""
        value &= M32
        assert self.trace.get(t, value) == value, "trace word %d redefined" % t
        self.trace[t] = value
        return V(value, ("W", t))

    def finish(self):
        for i, v in enumerate(self.vals):
            if v is None:                      # declared but never assigned -> holds 0 in the wasm
                self.vals[i] = const(0)


class Comp:
    """Base class: a component instance = a block of signals + sub-components."""

    def __init__(self, b, name):
        self.b, self.name = b, name

    def sig(self, nm, *dims):
        return self.b.alloc(self.name + "." + nm, *dims)

    def subs(self, **ctors):
        """Instantiate sub-components in circom's order: sorted by name; arrays by index.
        ctors: name -> callable(b, fullname) or list of callables (component array)."""
        for nm in sorted(ctors):
            c = ctors[nm]
            if isinstance(c, list):
                setattr(self, nm, [f(self.b, "%s.%s[%d]" % (self.name, nm, i)) for i, f in enumerate(c)])
            else:
                setattr(self, nm, c(self.b, "%s.%s" % (self.name, nm)))

    def S(self, idx, val):
        self.b.set(idx, val)

    # --- constraints (signal level, before any circom simplification); signal 0 is the constant 1 ---
    def c_mul(self, A, B, C):
        self.b.cons.append((dict(A), dict(B), dict(C)))

    def c_lin(self, terms):
        """sum coeff*signal = 0"""
        self.b.cons.append(({}, {}, dict(terms)))

    def c_alias(self, a, b):
        """a <== b between plain signals (lists are zipped)"""
        if isinstance(a, list):
            assert len(a) == len(b)
            for x, y in zip(a, b):
                self.c_alias(x, y)
        else:
            self.c_lin({a: 1, b: -1})

    def c_bool(self, x):
        self.c_mul({x: 1}, {0: 1, x: -1}, {})

    def G(self, idx):
        return self.b.get(idx)


# ------------------------------------------------------------------------------------------------
# circuits/blake3_common.circom
# ------------------------------------------------------------------------------------------------
class ToBits(Comp):
    """circuits/blake3_common.circom:142-154"""

    def __init__(self, b, name, n=32):
        super().__init__(b, name)
        self.n = n
        self.out = self.sig("out", n)
        self.inp = self.sig("inp")
        for i in range(n):
            self.c_bool(self.out[i])                                        # :149
        self.c_lin({self.inp: 1, **{self.out[i]: -(1 << i) for i in range(n)}})   # :153

    def run(self, x):
        self.S(self.inp, x)
        bits = [x.bit(i) for i in range(self.n)]
        for i in range(self.n):
            self.S(self.out[i], bits[i])
        self.ok = x.v == sum(bt.v << i for i, bt in enumerate(bits))     # inp === sum (:153)
        return bits


class XOR2(Comp):
    """circuits/blake3_common.circom:42-50"""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.out, self.x, self.y = self.sig("out"), self.sig("x"), self.sig("y")
        self.c_mul({self.x: 2}, {self.y: 1}, {self.x: 1, self.y: 1, self.out: -1})   # out <== x + y - 2*x*y (:49)

    def run(self, x, y, o):
        self.S(self.x, x), self.S(self.y, y), self.S(self.out, o)


class XorWord2(Comp):
    """circuits/blake3_common.circom:55-80 (n = 32).  out word lives at trace word t_out."""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.out_word = self.sig("out_word")
        self.x, self.y = self.sig("x"), self.sig("y")
        self.out_bits = self.sig("out_bits", 32)
        self.subs(tb_x=lambda b, n: ToBits(b, n), tb_y=lambda b, n: ToBits(b, n),
                  xor=[(lambda b, n: XOR2(b, n))] * 32)
        self.c_alias(self.tb_x.inp, self.x), self.c_alias(self.tb_y.inp, self.y)      # :65-66
        for i in range(32):
            self.c_alias(self.xor[i].x, self.tb_x.out[i]), self.c_alias(self.xor[i].y, self.tb_y.out[i])   # :72-73
            self.c_alias(self.out_bits[i], self.xor[i].out)                          # :74
        self.c_lin({self.out_word: 1, **{self.out_bits[i]: -(1 << i) for i in range(32)}})   # :79

    def run(self, x, y, t_out):
        self.S(self.x, x), self.S(self.y, y)
        xb, yb = self.tb_x.run(x), self.tb_y.run(y)
        w = self.b.tw(t_out, (x.v ^ y.v) & M32)
        for i in range(32):
            o = w.bit(i)
            assert o.v == xb[i].v ^ yb[i].v
            self.xor[i].run(xb[i], yb[i], o)
            self.S(self.out_bits[i], o)
        self.S(self.out_word, w)
        return w


class Bits3x(Comp):
    """Bits33 / Bits34: circuits/blake3_common.circom:160-178, 183-203."""

    def __init__(self, b, name, extra):
        super().__init__(b, name)
        self.extra = extra
        self.out_bits = self.sig("out_bits", 32)
        self.out_word = self.sig("out_word")
        self.inp = self.sig("inp")
        self.u = self.sig("u")
        if extra == 2:
            self.v = self.sig("v")
        for i in range(32):
            self.c_bool(self.out_bits[i])
        self.c_bool(self.u)
        rec = {self.inp: 1, self.u: -(1 << 32), **{self.out_bits[i]: -(1 << i) for i in range(32)}}
        if extra == 2:
            self.c_bool(self.v)
            rec[self.v] = -(1 << 33)
        self.c_lin(rec)                                                             # inp === sum + ... (:176 / :201)
        self.c_lin({self.out_word: 1, **{self.out_bits[i]: -(1 << i) for i in range(32)}})   # out_word <== sum

    def run(self, total, t_lo):
        """total = the integer sum; trace[t_lo] = low word, trace[t_lo+1] = carries."""
        lo = self.b.tw(t_lo, total & M32)
        hi = self.b.tw(t_lo + 1, total >> 32)
        self.S(self.inp, V(total, ("Q", t_lo)))
        bits = [lo.bit(i) for i in range(32)]
        for i in range(32):
            self.S(self.out_bits[i], bits[i])
        self.S(self.u, hi.bit(0))
        if self.extra == 2:
            self.S(self.v, hi.bit(1))
        self.S(self.out_word, lo)
        self.ok = (total >> 32) < (1 << self.extra)
        return lo, bits


class RotXorBits(Comp):
    """circuits/blake3_compression.circom:29-47"""

    def __init__(self, b, name, R):
        super().__init__(b, name)
        self.R = R
        self.out_bits = self.sig("out_bits", 32)
        self.out_word = self.sig("out_word")
        self.inp1_bits = self.sig("inp1_bits", 32)
        self.inp2_bits = self.sig("inp2_bits", 32)
        self.aux = self.sig("aux", 32)
        for i in range(32):
            self.c_mul({self.inp1_bits[i]: 2}, {self.inp2_bits[i]: 1},
                       {self.inp1_bits[i]: 1, self.inp2_bits[i]: 1, self.aux[i]: -1})   # :37
            self.c_alias(self.out_bits[i], self.aux[(i + R) % 32])                   # :42
        self.c_lin({self.out_word: 1, **{self.out_bits[i]: -(1 << i) for i in range(32)}})   # :46

    def run(self, b1, b2, w_out):
        R = self.R
        for i in range(32):
            self.S(self.inp1_bits[i], b1[i]), self.S(self.inp2_bits[i], b2[i])
        outb = [w_out.bit(i) for i in range(32)]
        for i in range(32):
            # aux[i] = inp1[i] ^ inp2[i];  out_bits[i] = aux[(i+R)%32]  =>  aux[j] = out_bits[(j-R)%32]
            a = outb[(i - R) % 32]
            assert a.v == b1[i].v ^ b2[i].v
            self.S(self.aux[i], a)
        for i in range(32):
            self.S(self.out_bits[i], outb[i])
        self.S(self.out_word, w_out)
        return outb


class RotXorWordBits(Comp):
    """circuits/blake3_compression.circom:53-67"""

    def __init__(self, b, name, R):
        super().__init__(b, name)
        self.out_bits = self.sig("out_bits", 32)
        self.out_word = self.sig("out_word")
        self.inp1_word = self.sig("inp1_word")
        self.inp2_bits = self.sig("inp2_bits", 32)
        self.subs(rx=lambda b, n: RotXorBits(b, n, R), tb=lambda b, n: ToBits(b, n))
        self.c_alias(self.tb.inp, self.inp1_word)                                    # :61
        self.c_alias(self.rx.inp1_bits, self.tb.out)                                # :62
        self.c_alias(self.rx.inp2_bits, self.inp2_bits)                             # :63
        self.c_alias(self.out_bits, self.rx.out_bits)                               # :64
        self.c_alias(self.out_word, self.rx.out_word)                               # :65

    def run(self, word, bits2, w_out):
        self.S(self.inp1_word, word)
        for i in range(32):
            self.S(self.inp2_bits[i], bits2[i])
        tb = self.tb.run(word)
        outb = self.rx.run(tb, bits2, w_out)
        for i in range(32):
            self.S(self.out_bits[i], outb[i])
        self.S(self.out_word, w_out)
        return outb


# ------------------------------------------------------------------------------------------------
# circuits/blake3_compression.circom
# ------------------------------------------------------------------------------------------------
class HalfFunG(Comp):
    """circuits/blake3_compression.circom:72-100"""

    def __init__(self, b, name, idx4, R1, R2):
        super().__init__(b, name)
        self.idx4, self.R1, self.R2 = idx4, R1, R2
        self.out = self.sig("out", 16)
        self.v = self.sig("v", 16)
        self.xy = self.sig("xy")
        self.subs(add1=lambda b, n: Bits3x(b, n, 2), add3=lambda b, n: Bits3x(b, n, 1),
                  rxor2=lambda b, n: RotXorWordBits(b, n, R1), rxor4=lambda b, n: RotXorWordBits(b, n, R2))
        a, bb, c, d = idx4
        for i in range(16):
            if i not in idx4:
                self.c_alias(self.out[i], self.v[i])                                 # :77-81
        self.c_lin({self.add1.inp: 1, self.v[a]: -1, self.v[bb]: -1, self.xy: -1})  # :88
        self.c_alias(self.rxor2.inp1_word, self.v[d])                               # :89
        self.c_alias(self.rxor2.inp2_bits, self.add1.out_bits)                      # :90
        self.c_lin({self.add3.inp: 1, self.v[c]: -1, self.rxor2.out_word: -1})      # :91
        self.c_alias(self.rxor4.inp1_word, self.v[bb])                              # :92
        self.c_alias(self.rxor4.inp2_bits, self.add3.out_bits)                      # :93
        self.c_alias(self.out[a], self.add1.out_word), self.c_alias(self.out[d], self.rxor2.out_word)   # :95-96
        self.c_alias(self.out[c], self.add3.out_word), self.c_alias(self.out[bb], self.rxor4.out_word)  # :97-98

    def run(self, v, xy, t_rec):
        """v: list of 16 V; xy: V; t_rec: trace index of this half's 8-word record."""
        a, bb, c, d = self.idx4
        b = self.b
        for i in range(16):
            self.S(self.v[i], v[i])
        self.S(self.xy, xy)
        # the kernel stores the "before" words redundantly in the record; the signals keep the
        # sources they already have (v[d], v[b]) -- same values.
        b.tw(t_rec + 2, v[d].v)
        b.tw(t_rec + 6, v[bb].v)
        a_new, a_bits = self.add1.run(v[a].v + v[bb].v + xy.v, t_rec + 0)
        d_new = b.tw(t_rec + 3, rotr(v[d].v ^ a_new.v, self.R1))
        self.rxor2.run(v[d], a_bits, d_new)
        c_new, c_bits = self.add3.run(v[c].v + d_new.v, t_rec + 4)
        b_new = b.tw(t_rec + 7, rotr(v[bb].v ^ c_new.v, self.R2))
        self.rxor4.run(v[bb], c_bits, b_new)
        out = list(v)
        out[a], out[d], out[c], out[bb] = a_new, d_new, c_new, b_new
        for i in range(16):
            self.S(self.out[i], out[i])
        self.ok = self.add1.ok and self.add3.ok and self.rxor2.tb.ok and self.rxor4.tb.ok
        return out


class MixFunG(Comp):
    """circuits/blake3_compression.circom:106-123"""

    def __init__(self, b, name, idx4):
        super().__init__(b, name)
        self.out = self.sig("out", 16)
        self.inp = self.sig("inp", 16)
        self.x, self.y = self.sig("x"), self.sig("y")
        self.subs(half1=lambda b, n: HalfFunG(b, n, idx4, 16, 12), half2=lambda b, n: HalfFunG(b, n, idx4, 8, 7))
        self.c_alias(self.half1.v, self.inp), self.c_alias(self.half1.xy, self.x)    # :115-116
        self.c_alias(self.half2.v, self.half1.out), self.c_alias(self.half2.xy, self.y)   # :119-120
        self.c_alias(self.out, self.half2.out)                                      # :121

    def run(self, inp, x, y, t_rec):
        for i in range(16):
            self.S(self.inp[i], inp[i])
        self.S(self.x, x), self.S(self.y, y)
        mid = self.half1.run(inp, x, t_rec)
        out = self.half2.run(mid, y, t_rec + 8)
        for i in range(16):
            self.S(self.out[i], out[i])
        self.ok = self.half1.ok and self.half2.ok
        return out


G_IDX = [(0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15),
         (0, 5, 10, 15), (1, 6, 11, 12), (2, 7, 8, 13), (3, 4, 9, 14)]   # blake3_compression.circom:145-153


class SingleRound(Comp):
    """circuits/blake3_compression.circom:128-161"""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.out = self.sig("out", 16)
        self.inp = self.sig("inp", 16)
        self.msg = self.sig("msg", 16)
        self.vs = self.sig("vs", 9, 16)
        self.subs(GS=[(lambda b, n, q=q: MixFunG(b, n, q)) for q in G_IDX])
        self.c_alias(self.vs[0], self.inp)                                           # :141
        for g in range(8):
            self.c_alias(self.GS[g].x, self.msg[2 * g]), self.c_alias(self.GS[g].y, self.msg[2 * g + 1])   # :145-153
            self.c_alias(self.GS[g].inp, self.vs[g]), self.c_alias(self.vs[g + 1], self.GS[g].out)         # :156-157
        self.c_alias(self.out, self.vs[8])                                          # :160

    def run(self, inp, msg, t_rec):
        for i in range(16):
            self.S(self.inp[i], inp[i]), self.S(self.msg[i], msg[i]), self.S(self.vs[0][i], inp[i])
        cur = inp
        self.ok = True
        for g in range(8):
            cur = self.GS[g].run(cur, msg[2 * g], msg[2 * g + 1], t_rec + 16 * g)
            self.ok = self.ok and self.GS[g].ok
            for i in range(16):
                self.S(self.vs[g + 1][i], cur[i])
        for i in range(16):
            self.S(self.out[i], cur[i])
        return cur


class Blake3Permute(Comp):
    """circuits/blake3_common.circom:15-26"""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.out = self.sig("out", 16)
        self.inp = self.sig("inp", 16)
        for j in range(16):
            self.c_alias(self.out[j], self.inp[SIGMA[j]])                            # :24

    def run(self, inp):
        out = [inp[SIGMA[j]] for j in range(16)]
        for i in range(16):
            self.S(self.inp[i], inp[i]), self.S(self.out[i], out[i])
        return out


class IVc(Comp):
    """circuits/blake3_compression.circom:17-24"""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.out = self.sig("out", 8)
        for j in range(8):
            self.c_lin({self.out[j]: 1, 0: -IV[j]})                                  # :23

    def run(self):
        o = [const(x) for x in IV]
        for i in range(8):
            self.S(self.out[i], o[i])
        return o


class Blake3Compression(Comp):
    """circuits/blake3_compression.circom:171-228"""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.out = self.sig("out", 16)
        self.h, self.m, self.t = self.sig("h", 8), self.sig("m", 16), self.sig("t", 2)
        self.bb, self.d = self.sig("b"), self.sig("d")
        self.init = self.sig("init", 16)
        self.subs(iv=lambda b, n: IVc(b, n), outXor=[lambda b, n: XorWord2(b, n)] * 16,
                  permuters=[lambda b, n: Blake3Permute(b, n)] * 6, rounds=[lambda b, n: SingleRound(b, n)] * 7)
        self.c_alias(self.init[0:8], self.h)                                         # :184
        self.c_alias(self.init[8:12], self.iv.out[0:4])                             # :185
        self.c_alias(self.init[12:14], self.t)                                      # :186
        self.c_alias(self.init[14], self.bb), self.c_alias(self.init[15], self.d)   # :187
        self.c_alias(self.rounds[0].msg, self.m), self.c_alias(self.rounds[0].inp, self.init)   # :194-195
        for i in range(6):
            self.c_alias(self.permuters[i].inp, self.m if i == 0 else self.permuters[i - 1].out)   # :203-207
            self.c_alias(self.rounds[i + 1].msg, self.permuters[i].out)              # :208
            self.c_alias(self.rounds[i + 1].inp, self.rounds[i].out)                 # :209
        for i in range(16):
            self.c_alias(self.outXor[i].x, self.rounds[6].out[i])                    # :217, :224
            self.c_alias(self.outXor[i].y, self.rounds[6].out[i + 8] if i < 8 else self.h[i - 8])   # :218, :226
            self.c_alias(self.out[i], self.outXor[i].out_word)                       # :219, :227

    def run(self, h, m, t, bb, d):
        """h[8], m[16], t[2], bb, d: V (their sources decided by the caller)."""
        for i in range(8):
            self.S(self.h[i], h[i])
        for i in range(16):
            self.S(self.m[i], m[i])
        self.S(self.t[0], t[0]), self.S(self.t[1], t[1]), self.S(self.bb, bb), self.S(self.d, d)
        iv = self.iv.run()
        init = list(h) + iv[:4] + [t[0], t[1], bb, d]
        for i in range(16):
            self.S(self.init[i], init[i])
        state, msg = init, list(m)
        self.ok = True
        for r in range(7):
            state = self.rounds[r].run(state, msg, TR_HG + 128 * r)
            self.ok = self.ok and self.rounds[r].ok
            if r < 6:
                msg = self.permuters[r].run(msg)
        out = []
        for i in range(8):
            out.append(self.outXor[i].run(state[i], state[i + 8], TR_OUT + i))
        for i in range(8, 16):
            out.append(self.outXor[i].run(state[i], h[i - 8], TR_OUT + i))
        for i in range(16):
            self.S(self.out[i], out[i])
            self.ok = self.ok and self.outXor[i].tb_x.ok and self.outXor[i].tb_y.ok
        return out


class CompressionModel:
    """main = Blake3Compression()  (circuits/main/blake3_compression.circom:6)."""
    n_inputs = 28

    def __init__(self, inputs, prime=BN254_R):
        """inputs: 28 u32 in declaration order h[8] m[16] t[2] b d."""
        assert len(inputs) == 28
        b = self.b = Builder(prime)
        b.tw(TR_ZERO, 0), b.tw(TR_ONE, 1)
        iv = [b.tw(TR_IN + i, x) for i, x in enumerate(inputs)]
        main = self.main = Blake3Compression(b, "main")
        main.run(iv[0:8], iv[8:24], iv[24:26], iv[26], iv[27])
        b.finish()
        self.ok = main.ok


# ================================================================================================
# Nova step circuit  (circuits/blake3_nova.circom, AS BUILT into the committed wasm files: without the two
# Num2Bits(8) range checks of lines 25-30) + circomlib 2.0.5 templates
# (not vendored in the reference; semantics per SURVEY.md appendix B, validated against the wasm).
# ================================================================================================
# Nova trace words, relative to TR_NOVA.  Emitted to csrc/nova_trace.h by tools/gen_tables.py.
NV = {}
_nv_next = [0]


def _nv(name, count=1):
    NV[name] = TR_NOVA + _nv_next[0]
    _nv_next[0] += count
    return NV[name]


_nv("IN", 32)            # the 32 inputs in declaration order: n_blocks block_count h[8] chunk_idx_low chunk_idx_high
#                          leaf_depth total_depth depth m[16] b      (circuits/blake3_nova.circom:173-191)
_nv("V1")                # check_parent.n2b.in  = depth + 256 - (leaf_depth - 1)
_nv("V2")                # exceed_depth.lt.n2b.in = leaf_depth + 256 - (depth + 1)
_nv("LDM1")              # leaf_depth - 1
_nv("DP1")               # depth + 1
_nv("IS_PARENT")
_nv("EXCEED")            # exceed_depth.out (0 whenever the witness exists)
_nv("IS_ROOT")
_nv("NOT_ROOT")
_nv("NOT_PARENT")
_nv("BC_FIRST")          # check_block_counts[0].out
_nv("BC_LAST")           # check_block_counts[1].out
_nv("IS_LAST")           # is_last_block = BC_LAST * not_parent (= last_block_flag_set.out)
_nv("FIRST_SET")         # first_block_flag_set.out
_nv("URF_TMP")           # use_root_flag_tmp.out
_nv("URF")               # use_root_flag
_nv("DLP")               # down_left_path.out
_nv("CDD")               # check_decr_depth.out
_nv("DECR")              # decr_depth
_nv("DEPTH_OUT")
_nv("PAD0")
_nv("NEG_DEPTH", 2)      # signed 64: 0 - depth                      (check_root.isz.in)
_nv("NEG_BC", 2)         # signed 64: 0 - block_count                (check_block_counts[0].isz.in)
_nv("NBM1", 2)           # signed 64: n_blocks - 1                   (check_block_counts[1].in[1])
_nv("BC_DIFF", 2)        # signed 64: n_blocks - 1 - block_count     (check_block_counts[1].isz.in)
_nv("BC_OUT", 2)         # block_count + (1 - is_parent)  (up to 2^32)
_nv("EQ_OUT", 2)         # bit i = eqs[i].out
_nv("BAD", 2)            # bit i = bit_at_depth[i]
_nv("TMPIV", 8)
_nv("TMP_DOWN", 16)
_nv("M_IS_PAR", 16)
_nv("TMP_IS_PAR", 16)
_nv("EQ_IN1", 128)       # signed 64 x 64: total_depth - i - 2        (eqs[i].in[1])
_nv("EQ_D", 128)         # signed 64 x 64: total_depth - i - 2 - depth (eqs[i].isz.in)
NOVA_TRACE_WORDS = TR_NOVA + _nv_next[0]


def s64_words(x):
    """two's complement 64-bit -> (lo, hi) u32"""
    x &= (1 << 64) - 1
    return x & M32, x >> 32


class NovaBuilder(Builder):
    def ts64(self, t, x):
        """Define trace[t], trace[t+1] := signed 64-bit x; returns the field-valued V (kind S)."""
        lo, hi = s64_words(x)
        self.tw(t, lo), self.tw(t + 1, hi)
        return V(x % self.p, ("S", t))

    def inv_of(self, s):
        """IsZero's inv of an S-kind value: kind I refers to the same signed argument."""
        assert s.s[0] == "S"
        return V(pow(s.v, -1, self.p) if s.v else 0, ("I", s.s[1]))


class IsZero(Comp):
    """circomlib comparators.circom IsZero: signals out; in; inv."""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.out, self.inp, self.inv = self.sig("out"), self.sig("in"), self.sig("inv")
        self.c_mul({self.inp: 1}, {self.inv: 1}, {0: 1, self.out: -1})               # out <== -in*inv + 1
        self.c_mul({self.inp: 1}, {self.out: 1}, {})                                # in*out === 0

    def run(self, x, out):
        """x: S-kind value; out: V for the result word/bit (caller decides where it lives)."""
        self.S(self.inp, x)
        self.S(self.inv, self.b.inv_of(x))
        assert out.v == (1 if x.v == 0 else 0)
        self.S(self.out, out)
        return out


class IsEqual(Comp):
    """circomlib IsEqual: out; in[2]; sub isz.   isz.in = in[1] - in[0]."""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.out, self.inp = self.sig("out"), self.sig("in", 2)
        self.subs(isz=lambda b, n: IsZero(b, n))
        self.c_lin({self.isz.inp: 1, self.inp[1]: -1, self.inp[0]: 1})               # isz.in <== in[1] - in[0]
        self.c_alias(self.out, self.isz.out)

    def run(self, in0, in1, diff, out):
        self.S(self.inp[0], in0), self.S(self.inp[1], in1)
        assert diff.v == (in1.v - in0.v) % self.b.p
        self.isz.run(diff, out)
        self.S(self.out, out)
        return out


class Num2Bits(Comp):
    """circomlib bitify.circom Num2Bits(n): out[n]; in."""

    def __init__(self, b, name, n):
        super().__init__(b, name)
        self.n = n
        self.out, self.inp = self.sig("out", n), self.sig("in")
        for i in range(n):
            self.c_bool(self.out[i])
        self.c_lin({self.inp: 1, **{self.out[i]: -(1 << i) for i in range(n)}})

    def run(self, x):
        self.S(self.inp, x)
        bits = []
        for i in range(self.n):
            bt = x.bit(i) if i < 64 else const(0)
            bits.append(bt)
            self.S(self.out[i], bt)
        self.ok = x.v == sum(bt.v << i for i, bt in enumerate(bits))
        return bits


class LessThan(Comp):
    """circomlib LessThan(n): out; in[2]; sub n2b = Num2Bits(n+1).  n2b.in = in[0] + 2^n - in[1]; out = 1 - n2b.out[n]."""

    def __init__(self, b, name, n):
        super().__init__(b, name)
        self.n = n
        self.out, self.inp = self.sig("out"), self.sig("in", 2)
        self.subs(n2b=lambda b, nm: Num2Bits(b, nm, n + 1))
        self.c_lin({self.n2b.inp: 1, self.inp[0]: -1, 0: -(1 << n), self.inp[1]: 1})   # n2b.in <== in[0] + (1<<n) - in[1]
        self.c_lin({self.out: 1, 0: -1, self.n2b.out[n]: 1})                        # out <== 1 - n2b.out[n]

    def run(self, in0, in1, vword, out):
        self.S(self.inp[0], in0), self.S(self.inp[1], in1)
        bits = self.n2b.run(vword)
        assert out.v == 1 - bits[self.n].v
        self.S(self.out, out)
        self.ok = self.n2b.ok
        return out


class GreaterEqThan(Comp):
    """circomlib GreaterEqThan(n): out; in[2]; sub lt = LessThan(n) with lt.in = (in[1], in[0]+1)."""

    def __init__(self, b, name, n):
        super().__init__(b, name)
        self.out, self.inp = self.sig("out"), self.sig("in", 2)
        self.subs(lt=lambda b, nm: LessThan(b, nm, n))
        self.c_alias(self.lt.inp[0], self.inp[1])
        self.c_lin({self.lt.inp[1]: 1, self.inp[0]: -1, 0: -1})                      # lt.in[1] <== in[0] + 1
        self.c_alias(self.out, self.lt.out)

    def run(self, in0, in1, in0p1, vword, out):
        self.S(self.inp[0], in0), self.S(self.inp[1], in1)
        self.lt.run(in1, in0p1, vword, out)
        self.S(self.out, out)
        self.ok = self.lt.ok
        return out


class Gate2(Comp):
    """circomlib gates.circom AND / OR: out; a; b.   AND: out <== a*b;  OR: out <== a + b - a*b"""

    def __init__(self, b, name, kind="AND"):
        super().__init__(b, name)
        self.out, self.a, self.bb = self.sig("out"), self.sig("a"), self.sig("b")
        if kind == "AND":
            self.c_mul({self.a: 1}, {self.bb: 1}, {self.out: 1})
        else:
            self.c_mul({self.a: 1}, {self.bb: 1}, {self.a: 1, self.bb: 1, self.out: -1})

    def run(self, a, b, out):
        self.S(self.a, a), self.S(self.bb, b), self.S(self.out, out)
        return out


class NOT(Comp):
    """circomlib NOT: out; in.   out = 1 + in - 2*in"""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.out, self.inp = self.sig("out"), self.sig("in")
        self.c_lin({self.out: 1, 0: -1, self.inp: 1})                               # out <== 1 + in - 2*in

    def run(self, x, out):
        assert out.v == 1 - x.v
        self.S(self.inp, x), self.S(self.out, out)
        return out


class CheckDepth(Comp):
    """Blake3NovaTreePath_CheckDepth -- circuits/blake3_nova.circom:13-45 minus lines 25-30 (as built)."""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.is_root, self.is_parent = self.sig("is_root"), self.sig("is_parent")
        self.depth, self.leaf_depth = self.sig("depth"), self.sig("leaf_depth")
        self.subs(check_root=lambda b, n: IsEqual(b, n), check_parent=lambda b, n: LessThan(b, n, 8),
                  exceed_depth=lambda b, n: GreaterEqThan(b, n, 8))
        self.c_alias(self.check_root.inp[0], self.depth), self.c_lin({self.check_root.inp[1]: 1})   # :20-21
        self.c_alias(self.is_root, self.check_root.out)                             # :23
        self.c_alias(self.check_parent.inp[0], self.depth)                          # :32
        self.c_lin({self.check_parent.inp[1]: 1, self.leaf_depth: -1, 0: 1})        # :33
        self.c_alias(self.is_parent, self.check_parent.out)                         # :38
        self.c_alias(self.exceed_depth.inp[0], self.depth), self.c_alias(self.exceed_depth.inp[1], self.leaf_depth)   # :42-43
        self.c_lin({self.exceed_depth.out: 1})                                      # exceed_depth.out === 0 (:44)

    def run(self, depth, leaf_depth):
        b = self.b
        self.S(self.depth, depth), self.S(self.leaf_depth, leaf_depth)
        is_root = b.tw(NV["IS_ROOT"], 1 if depth.v == 0 else 0)
        self.check_root.run(depth, const(0), b.ts64(NV["NEG_DEPTH"], -depth.v), is_root)          # :19-23
        self.S(self.is_root, is_root)
        v1 = depth.v + 256 - (leaf_depth.v - 1)
        v2 = leaf_depth.v + 256 - (depth.v + 1)
        self.ok = 0 <= v1 < 512 and 0 <= v2 < 512
        if not self.ok:
            return None, None
        V1, V2 = b.tw(NV["V1"], v1), b.tw(NV["V2"], v2)
        is_parent = b.tw(NV["IS_PARENT"], 1 - ((v1 >> 8) & 1))
        self.check_parent.run(depth, b.tw(NV["LDM1"], leaf_depth.v - 1), V1, is_parent)           # :31-33
        self.S(self.is_parent, is_parent)                                                        # :38
        b.tw(NV["EXCEED"], 1 - ((v2 >> 8) & 1))
        self.ok = self.ok and ((v2 >> 8) & 1) == 1                                               # exceed_depth.out === 0 (:44)
        if not self.ok:
            return None, None
        # the constraint pins the signal to the constant 0 (circom drops it from the witness), so it is modelled as one
        self.exceed_depth.run(depth, leaf_depth, b.tw(NV["DP1"], depth.v + 1), V2, const(0))      # :41-43
        return is_root, is_parent


class DownLeftPath(Comp):
    """Blake3GetDownLeftPath -- circuits/blake3_nova.circom:47-84"""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.out = self.sig("out")
        self.depth, self.leaf_idx = self.sig("depth"), self.sig("leaf_idx")
        self.is_parent, self.total_depth = self.sig("is_parent"), self.sig("total_depth")
        self.bit_at_depth = self.sig("bit_at_depth", 65)
        self.subs(eqs=[lambda b, n: IsEqual(b, n)] * 64, n2b=lambda b, n: Num2Bits(b, n, 65))
        self.c_alias(self.n2b.inp, self.leaf_idx)                                    # :59
        for i in range(64):
            self.c_alias(self.eqs[i].inp[0], self.depth)                             # :64, :69
            self.c_lin({self.eqs[i].inp[1]: 1, self.total_depth: -1, 0: i + 2})
            prev = {self.bit_at_depth[i - 1]: -1} if i else {}
            self.c_mul({0: 1, self.n2b.out[i]: -1}, {self.eqs[i].out: 1}, {self.bit_at_depth[i]: 1, **prev})   # :65, :70
        self.c_mul({self.is_parent: 1}, {self.bit_at_depth[63]: 1}, {self.out: 1, 0: -1, self.is_parent: 1})   # :79
        self.c_bool(self.out)                                                       # :81

    def run(self, depth, leaf_idx, is_parent, total_depth):
        b = self.b
        self.S(self.depth, depth), self.S(self.leaf_idx, leaf_idx)
        self.S(self.is_parent, is_parent), self.S(self.total_depth, total_depth)
        bits = self.n2b.run(leaf_idx)                                                            # :57-59
        eq_mask = 0
        for i in range(64):
            if total_depth.v - i - 2 == depth.v:
                eq_mask |= 1 << i
        eqw = [b.tw(NV["EQ_OUT"], eq_mask & M32), b.tw(NV["EQ_OUT"] + 1, eq_mask >> 32)]
        acc, bad_mask = 0, 0
        for i in range(64):
            acc += (1 - bits[i].v) * ((eq_mask >> i) & 1)                                        # :65, :70
            assert acc in (0, 1)
            bad_mask |= acc << i
        badw = [b.tw(NV["BAD"], bad_mask & M32), b.tw(NV["BAD"] + 1, bad_mask >> 32)]
        for i in range(64):
            in1 = b.ts64(NV["EQ_IN1"] + 2 * i, total_depth.v - i - 2)                            # :64, :69
            d = b.ts64(NV["EQ_D"] + 2 * i, total_depth.v - i - 2 - depth.v)
            self.eqs[i].run(depth, in1, d, eqw[i >> 5].bit(i & 31))
            self.S(self.bit_at_depth[i], badw[i >> 5].bit(i & 31))
        out = b.tw(NV["DLP"], (1 - is_parent.v) + is_parent.v * ((bad_mask >> 63) & 1))           # :79
        self.S(self.out, out)
        self.ok = out.v in (0, 1) and self.n2b.ok                                                # :81
        return out


class FinalM(Comp):
    """Blake3GetFinal_m -- circuits/blake3_nova.circom:86-120.  out_m lives in the compression input words."""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.out_m = self.sig("out_m", 16)
        self.h, self.m = self.sig("h", 8), self.sig("m", 16)
        self.is_parent, self.depth = self.sig("is_parent"), self.sig("depth")
        self.total_depth, self.chunk_idx = self.sig("total_depth"), self.sig("chunk_idx")
        self.m_is_parent, self.tmp_down, self.tmp_is_par = self.sig("m_is_parent", 16), self.sig("tmp_down", 16), self.sig("tmp_is_par", 16)
        self.subs(down_left_path=lambda b, n: DownLeftPath(b, n))
        dl = self.down_left_path
        self.c_alias(dl.depth, self.depth), self.c_alias(dl.leaf_idx, self.chunk_idx)   # :98-99
        self.c_alias(dl.is_parent, self.is_parent), self.c_alias(dl.total_depth, self.total_depth)   # :100-101
        for i in range(16):
            if i < 8:
                self.c_mul({self.h[i]: 1}, {dl.out: 1}, {self.tmp_down[i]: 1})                            # :109
                self.c_mul({self.m[i]: 1}, {0: 1, dl.out: -1}, {self.m_is_parent[i]: 1, self.tmp_down[i]: -1})   # :111
            else:
                self.c_mul({self.h[i - 8]: 1}, {0: 1, dl.out: -1}, {self.tmp_down[i]: 1})                  # :113
                self.c_mul({self.m[i - 8]: 1}, {dl.out: 1}, {self.m_is_parent[i]: 1, self.tmp_down[i]: -1})   # :114
            self.c_mul({self.m_is_parent[i]: 1}, {self.is_parent: 1}, {self.tmp_is_par[i]: 1})            # :116
            self.c_mul({self.m[i]: 1}, {0: 1, self.is_parent: -1}, {self.out_m[i]: 1, self.tmp_is_par[i]: -1})   # :117

    def run(self, h, m, is_parent, depth, total_depth, chunk_idx):
        b = self.b
        for i in range(8):
            self.S(self.h[i], h[i])
        for i in range(16):
            self.S(self.m[i], m[i])
        self.S(self.is_parent, is_parent), self.S(self.depth, depth)
        self.S(self.total_depth, total_depth), self.S(self.chunk_idx, chunk_idx)
        dlp = self.down_left_path.run(depth, chunk_idx, is_parent, total_depth)                  # :97-101
        out = []
        for i in range(16):
            if i < 8:
                td = h[i].v * dlp.v                                                              # :109
                mp = m[i].v * (1 - dlp.v) + td                                                   # :111
            else:
                td = h[i - 8].v * (1 - dlp.v)                                                    # :113
                mp = m[i - 8].v * dlp.v + td                                                     # :114
            tp = mp * is_parent.v                                                                # :116
            om = m[i].v * (1 - is_parent.v) + tp                                                 # :117
            self.S(self.tmp_down[i], b.tw(NV["TMP_DOWN"] + i, td))
            self.S(self.m_is_parent[i], b.tw(NV["M_IS_PAR"] + i, mp))
            self.S(self.tmp_is_par[i], b.tw(NV["TMP_IS_PAR"] + i, tp))
            o = b.tw(TR_IN + 8 + i, om)
            self.S(self.out_m[i], o)
            out.append(o)
        self.ok = self.down_left_path.ok
        return out


class GetFlag(Comp):
    """Blake3GetFlag(D_FLAGS = 0) -- circuits/blake3_nova.circom:122-167"""

    def __init__(self, b, name):
        super().__init__(b, name)
        self.out, self.is_last_block = self.sig("out"), self.sig("is_last_block")
        self.is_parent, self.is_root = self.sig("is_parent"), self.sig("is_root")
        self.block_count, self.n_blocks = self.sig("block_count"), self.sig("n_blocks")
        self.use_root_flag = self.sig("use_root_flag")
        self.subs(not_root=lambda b, n: NOT(b, n), not_parent=lambda b, n: NOT(b, n),
                  check_block_counts=[lambda b, n: IsEqual(b, n)] * 2,
                  first_block_flag_set=lambda b, n: Gate2(b, n, "AND"), last_block_flag_set=lambda b, n: Gate2(b, n, "AND"),
                  use_root_flag_tmp=lambda b, n: Gate2(b, n, "OR"))
        cb = self.check_block_counts
        self.c_alias(self.not_root.inp, self.is_root), self.c_alias(self.not_parent.inp, self.is_parent)   # :136-137
        self.c_alias(cb[0].inp[0], self.block_count), self.c_lin({cb[0].inp[1]: 1})                       # :141-142
        self.c_alias(cb[1].inp[0], self.block_count)                                                     # :144
        self.c_lin({cb[1].inp[1]: 1, self.n_blocks: -1, 0: 1})                                           # :145
        self.c_mul({cb[1].out: 1}, {self.not_parent.out: 1}, {self.is_last_block: 1})                    # :148
        self.c_alias(self.first_block_flag_set.a, cb[0].out), self.c_alias(self.first_block_flag_set.bb, self.not_parent.out)   # :151
        self.c_alias(self.last_block_flag_set.a, cb[1].out), self.c_alias(self.last_block_flag_set.bb, self.not_parent.out)     # :152
        self.c_alias(self.use_root_flag_tmp.a, self.is_parent), self.c_alias(self.use_root_flag_tmp.bb, cb[1].out)             # :157
        self.c_mul({self.use_root_flag_tmp.out: 1}, {self.is_root: 1}, {self.use_root_flag: 1})          # :158
        self.c_lin({self.out: 1, self.first_block_flag_set.out: -1, self.last_block_flag_set.out: -2,
                    self.use_root_flag: -8, self.is_parent: -4})                                         # :161-165 (D_FLAGS = 0)

    def run(self, is_parent, is_root, block_count, n_blocks):
        b = self.b
        self.S(self.is_parent, is_parent), self.S(self.is_root, is_root)
        self.S(self.block_count, block_count), self.S(self.n_blocks, n_blocks)
        self.not_root.run(is_root, b.tw(NV["NOT_ROOT"], 1 - is_root.v))                           # :136
        not_parent = self.not_parent.run(is_parent, b.tw(NV["NOT_PARENT"], 1 - is_parent.v))      # :137
        first = b.tw(NV["BC_FIRST"], 1 if block_count.v == 0 else 0)
        self.check_block_counts[0].run(block_count, const(0), b.ts64(NV["NEG_BC"], -block_count.v), first)     # :141-142
        last = b.tw(NV["BC_LAST"], 1 if block_count.v == n_blocks.v - 1 else 0)
        self.check_block_counts[1].run(block_count, b.ts64(NV["NBM1"], n_blocks.v - 1),
                                       b.ts64(NV["BC_DIFF"], n_blocks.v - 1 - block_count.v), last)           # :144-145
        is_last = b.tw(NV["IS_LAST"], last.v * not_parent.v)                                      # :148
        self.S(self.is_last_block, is_last)
        fs = self.first_block_flag_set.run(first, not_parent, b.tw(NV["FIRST_SET"], first.v * not_parent.v))   # :151
        ls = self.last_block_flag_set.run(last, not_parent, is_last)                              # :152
        tmp = self.use_root_flag_tmp.run(is_parent, last, b.tw(NV["URF_TMP"], is_parent.v + last.v - is_parent.v * last.v))   # :157
        urf = b.tw(NV["URF"], tmp.v * is_root.v)                                                  # :158
        self.S(self.use_root_flag, urf)
        out = b.tw(TR_IN + 27, fs.v + 2 * ls.v + 8 * urf.v + 4 * is_parent.v)                     # :161-165
        self.S(self.out, out)
        return out, is_last


class Blake3Nova(Comp):
    """Blake3Nova(0) -- circuits/blake3_nova.circom:169-267"""

    def __init__(self, b, name):
        super().__init__(b, name)
        s = self.sig
        self.n_blocks_out, self.block_count_out, self.h_out = s("n_blocks_out"), s("block_count_out"), s("h_out", 8)
        self.total_depth_out, self.depth_out = s("total_depth_out"), s("depth_out")
        self.chunk_idx_low_out, self.chunk_idx_high_out, self.leaf_depth_out = s("chunk_idx_low_out"), s("chunk_idx_high_out"), s("leaf_depth_out")
        self.n_blocks, self.block_count, self.h = s("n_blocks"), s("block_count"), s("h", 8)
        self.chunk_idx_low, self.chunk_idx_high = s("chunk_idx_low"), s("chunk_idx_high")
        self.leaf_depth, self.total_depth, self.depth = s("leaf_depth"), s("total_depth"), s("depth")
        self.m, self.bb = s("m", 16), s("b")
        self.tmpIV, self.h_compression, self.decr_depth = s("tmpIV", 8), s("h_compression", 8), s("decr_depth")
        self.subs(blake3Compression=lambda b, n: Blake3Compression(b, n), check_decr_depth=lambda b, n: Gate2(b, n, "OR"),
                  check_depth=lambda b, n: CheckDepth(b, n), comp_d=lambda b, n: GetFlag(b, n),
                  final_m=lambda b, n: FinalM(b, n), iv=lambda b, n: IVc(b, n))
        cd, fl, fm, bc = self.check_depth, self.comp_d, self.final_m, self.blake3Compression
        self.c_alias(cd.depth, self.depth), self.c_alias(cd.leaf_depth, self.leaf_depth)                 # :206-207
        self.c_alias(fl.is_parent, cd.is_parent), self.c_alias(fl.is_root, cd.is_root)                   # :211-212
        self.c_alias(fl.block_count, self.block_count), self.c_alias(fl.n_blocks, self.n_blocks)         # :213-214
        self.c_alias(fm.h, self.h), self.c_alias(fm.m, self.m), self.c_alias(fm.is_parent, cd.is_parent)  # :222-224
        self.c_alias(fm.depth, self.depth), self.c_alias(fm.total_depth, self.total_depth)               # :225-226
        self.c_lin({fm.chunk_idx: 1, self.chunk_idx_low: -1, self.chunk_idx_high: -(1 << 32)})           # :227
        for i in range(8):
            self.c_mul({self.iv.out[i]: 1}, {cd.is_parent: 1}, {self.tmpIV[i]: 1})                        # :231
            self.c_mul({self.h[i]: 1}, {0: 1, cd.is_parent: -1}, {self.h_compression[i]: 1, self.tmpIV[i]: -1})   # :232
            self.c_alias(self.h_out[i], bc.out[i])                                                       # :248
        self.c_alias(bc.m, fm.out_m), self.c_alias(bc.h, self.h_compression)                             # :236-237
        self.c_alias(bc.d, fl.out), self.c_alias(bc.bb, self.bb)                                         # :238-239
        self.c_mul({self.chunk_idx_high: 1}, {0: 1, cd.is_parent: -1}, {bc.t[1]: 1})                      # :244
        self.c_mul({self.chunk_idx_low: 1}, {0: 1, cd.is_parent: -1}, {bc.t[0]: 1})                       # :245
        self.c_lin({self.block_count_out: 1, self.block_count: -1, 0: -1, cd.is_parent: 1})              # :251
        self.c_alias(self.n_blocks_out, self.n_blocks)                                                   # :252
        self.c_alias(self.check_decr_depth.a, fl.is_last_block), self.c_alias(self.check_decr_depth.bb, cd.is_parent)   # :255-256
        self.c_mul({self.check_decr_depth.out: 1}, {0: 1, cd.is_root: -1}, {self.decr_depth: 1})          # :258
        self.c_bool(self.decr_depth)                                                                     # :259
        self.c_lin({self.depth_out: 1, self.depth: -1, self.decr_depth: 1})                              # :262
        self.c_alias(self.total_depth_out, self.total_depth), self.c_alias(self.chunk_idx_low_out, self.chunk_idx_low)   # :263-264
        self.c_alias(self.chunk_idx_high_out, self.chunk_idx_high), self.c_alias(self.leaf_depth_out, self.leaf_depth)   # :265-266

    def run(self, w):
        """w: the 32 input words as V (trace words NV_IN..)."""
        b = self.b
        n_blocks, block_count, h = w[0], w[1], w[2:10]
        low, high, leaf_depth, total_depth, depth, m, bb = w[10], w[11], w[12], w[13], w[14], w[15:31], w[31]
        for sidx, val in ((self.n_blocks, n_blocks), (self.block_count, block_count), (self.chunk_idx_low, low),
                          (self.chunk_idx_high, high), (self.leaf_depth, leaf_depth), (self.total_depth, total_depth),
                          (self.depth, depth), (self.bb, bb)):
            self.S(sidx, val)
        for i in range(8):
            self.S(self.h[i], h[i])
        for i in range(16):
            self.S(self.m[i], m[i])
        is_root, is_parent = self.check_depth.run(depth, leaf_depth)                              # :205-207
        self.ok = self.check_depth.ok
        if not self.ok:
            return
        d, is_last = self.comp_d.run(is_parent, is_root, block_count, n_blocks)                   # :210-214
        iv = self.iv.run()
        chunk_idx = V(low.v + (high.v << 32), ("Q", NV["IN"] + 10))                               # :227
        out_m = self.final_m.run(h, m, is_parent, depth, total_depth, chunk_idx)                  # :222-227
        hc = []
        for i in range(8):
            tiv = b.tw(NV["TMPIV"] + i, IV[i] * is_parent.v)                                      # :231
            self.S(self.tmpIV[i], tiv)
            x = b.tw(TR_IN + i, h[i].v * (1 - is_parent.v) + tiv.v)                               # :232
            self.S(self.h_compression[i], x)
            hc.append(x)
        t1 = b.tw(TR_IN + 25, high.v * (1 - is_parent.v))                                         # :244
        t0 = b.tw(TR_IN + 24, low.v * (1 - is_parent.v))                                          # :245
        b.tw(TR_IN + 26, bb.v)          # the kernel copies b next to the other compression inputs; the signal IS main.b
        out = self.blake3Compression.run(hc, out_m, [t0, t1], bb, d)                              # :235-245
        for i in range(8):
            self.S(self.h_out[i], out[i])                                                        # :248
        bco = block_count.v + (1 - is_parent.v)                                                  # :251
        lo, hi = s64_words(bco)
        b.tw(NV["BC_OUT"], lo), b.tw(NV["BC_OUT"] + 1, hi)
        self.S(self.block_count_out, V(bco, ("Q", NV["BC_OUT"])))
        self.S(self.n_blocks_out, n_blocks)                                                      # :252
        cdd = self.check_decr_depth.run(is_last, is_parent, b.tw(NV["CDD"], is_last.v + is_parent.v - is_last.v * is_parent.v))   # :254-256
        decr = b.tw(NV["DECR"], cdd.v * (1 - is_root.v))                                          # :258
        self.S(self.decr_depth, decr)
        self.S(self.depth_out, b.tw(NV["DEPTH_OUT"], depth.v - decr.v))                           # :262
        self.S(self.total_depth_out, total_depth), self.S(self.chunk_idx_low_out, low)           # :263-265
        self.S(self.chunk_idx_high_out, high), self.S(self.leaf_depth_out, leaf_depth)
        self.ok = (self.ok and self.final_m.ok and self.blake3Compression.ok and decr.v in (0, 1))


class NovaModel:
    """main = Blake3Nova(0)  (circuits/main/blake3_nova.circom:6)."""
    n_inputs = 32

    def __init__(self, inputs, prime=BN254_R, o1=False):
        assert len(inputs) == 32
        b = self.b = NovaBuilder(prime)
        b.tw(TR_ZERO, 0), b.tw(TR_ONE, 1)
        w = [b.tw(NV["IN"] + i, x) for i, x in enumerate(inputs)]
        main = self.main = Blake3Nova(b, "main")
        main.run(w)
        b.finish()
        self.ok = main.ok


def random_nova_inputs(rng, edge=0):
    """Step inputs covering leaf / first / last / parent / root cases."""
    r32 = lambda: rng.getrandbits(32)
    leaf_depth = rng.randrange(1, 65)
    total_depth = leaf_depth if rng.random() < 0.8 else rng.randrange(1, 65)
    depth = rng.randrange(0, leaf_depth)
    if edge == 1:
        leaf_depth = total_depth = 1
        depth = 0
    elif edge == 2:
        leaf_depth = total_depth = 64
        depth = 63
    n_blocks = rng.randrange(1, 17)
    block_count = rng.randrange(0, n_blocks) if rng.random() < 0.9 else rng.randrange(0, 17)
    if rng.random() < 0.3:
        block_count = n_blocks - 1
    b = rng.choice([0, 1, 4, 63, 64, rng.randrange(0, 65)])
    if edge == 3:                      # depth == leaf_depth: exceed_depth.out === 0 fails  (blake3_nova.circom:44)
        depth = leaf_depth
    elif edge == 4:                    # legal but far outside the honest domain: huge depth, generic inverses
        depth = 4000000000 + rng.randrange(1000)
        leaf_depth = depth + rng.randrange(1, 258)
        total_depth = rng.getrandbits(32)
    elif edge == 5:                    # n_blocks = 0, block_count = 2^32 - 1 (block_count_out = 2^32)
        n_blocks, block_count, depth = 0, 0xFFFFFFFF, leaf_depth - 1
    elif edge == 6:                    # Num2Bits(9) range assert: leaf_depth - depth > 257
        depth, leaf_depth = 3, 300
    return ([n_blocks, block_count] + [r32() for _ in range(8)] + [r32(), r32() if rng.random() < 0.5 else 0] +
            [leaf_depth, total_depth, depth] + [r32() for _ in range(16)] + [b])
