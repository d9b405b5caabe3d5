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
          ("F", t)        the 256-bit field element trace[t..t+8)
          ("N", t)        the field element  p - trace[t]  (0 if trace[t] == 0)      [nova O1 only]
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
