#!/usr/bin/env python3
"""wasm2c.py -- ahead-of-time translator from a circom-generated .wasm witness program to C.

TEST INFRASTRUCTURE ONLY (part of oracle/, never on the product path).

Why this exists: the reference's hot path *is* a WebAssembly program (build/**/**.wasm, produced by
circom 2.1.6) hosted by witness_calculator.js.  This image has no Node and no wasm runtime, so in
order to run the UNMODIFIED reference program we translate its byte code to C and compile it with
gcc.  The translation is purely mechanical (one C statement per wasm instruction); nothing about
BLAKE3 or circom is known to this file.  Output goes to oracle/_ref/ (git-ignored; it is derived
from the reference's binary artefacts and must not enter this repo's history).

Supported subset = everything the five reference programs use (MVP integer opcodes, void block
types, one funcref table, one memory, active data segments with constant offsets) plus the rest of
the i32/i64 integer instruction set.  Anything else raises.

Generated C ABI (one translation unit per wasm, instance-based so that it is thread safe):
    typedef struct W W;                       /* defined in wasm_rt.h */
    void     wasm_instantiate(W *w);          /* allocate memory, copy data segments */
    <ret>    wx_<export>(W *w, <args>);       /* one per exported function */
    imports: <ret> wi_<module>_<name>(W *w, <args>);   /* provided by the harness */
"""
import sys, struct


class Reader:
    def __init__(self, b, p=0, end=None):
        self.b, self.p, self.end = b, p, len(b) if end is None else end

    def eof(self):
        return self.p >= self.end

    def u8(self):
        v = self.b[self.p]
        self.p += 1
        return v

    def leb_u(self):
        r = s = 0
        while True:
            c = self.u8()
            r |= (c & 0x7F) << s
            s += 7
            if not c & 0x80:
                return r

    def leb_s(self, bits):
        r = s = 0
        while True:
            c = self.u8()
            r |= (c & 0x7F) << s
            s += 7
            if not c & 0x80:
                if c & 0x40:
                    r -= 1 << s
                return r

    def name(self):
        n = self.leb_u()
        s = self.b[self.p:self.p + n].decode()
        self.p += n
        return s

    def bytes(self, n):
        s = self.b[self.p:self.p + n]
        self.p += n
        return s


VT = {0x7F: "i32", 0x7E: "i64"}
CT = {"i32": "uint32_t", "i64": "uint64_t"}


class Module:
    def __init__(self, data):
        assert data[:8] == b"\0asm\x01\0\0\0", "not a wasm v1 module"
        self.types, self.imports, self.funcs, self.exports = [], [], [], []
        self.table, self.codes, self.datas, self.mem_pages = [], [], [], (0, None)
        self.func_names = {}
        r = Reader(data, 8)
        while not r.eof():
            sid = r.u8()
            size = r.leb_u()
            s = Reader(data, r.p, r.p + size)
            r.p += size
            getattr(self, "sec_%d" % sid, lambda s: None)(s)

    def sec_0(self, s):
        if s.name() != "name":
            return
        while not s.eof():
            sub = s.u8()
            n = s.leb_u()
            if sub == 1:
                for _ in range(s.leb_u()):
                    i = s.leb_u()
                    self.func_names[i] = s.name()
            else:
                s.p += n

    def sec_1(self, s):
        for _ in range(s.leb_u()):
            assert s.u8() == 0x60
            ps = [VT[s.u8()] for _ in range(s.leb_u())]
            rs = [VT[s.u8()] for _ in range(s.leb_u())]
            assert len(rs) <= 1
            self.types.append((ps, rs))

    def sec_2(self, s):
        for _ in range(s.leb_u()):
            mod, nm, kind = s.name(), s.name(), s.u8()
            if kind == 0:
                self.imports.append((mod, nm, s.leb_u()))
            elif kind == 2:  # imported memory (not used by the reference, but harmless)
                flag = s.u8()
                lo = s.leb_u()
                hi = s.leb_u() if flag & 1 else None
                self.mem_pages = (lo, hi)
            else:
                raise NotImplementedError("import kind %d" % kind)

    def sec_3(self, s):
        self.funcs = [s.leb_u() for _ in range(s.leb_u())]

    def sec_4(self, s):
        n = s.leb_u()
        assert n == 1
        assert s.u8() == 0x70
        flag = s.u8()
        s.leb_u()
        if flag & 1:
            s.leb_u()

    def sec_5(self, s):
        assert s.leb_u() == 1
        flag = s.u8()
        lo = s.leb_u()
        hi = s.leb_u() if flag & 1 else None
        self.mem_pages = (lo, hi)

    def sec_6(self, s):
        assert s.leb_u() == 0, "globals unsupported"

    def sec_7(self, s):
        for _ in range(s.leb_u()):
            nm, kind, idx = s.name(), s.u8(), s.leb_u()
            self.exports.append((nm, kind, idx))

    def const_expr(self, s):
        assert s.u8() == 0x41
        v = s.leb_s(32)
        assert s.u8() == 0x0B
        return v & 0xFFFFFFFF

    def sec_9(self, s):
        for _ in range(s.leb_u()):
            assert s.leb_u() == 0
            off = self.const_expr(s)
            fs = [s.leb_u() for _ in range(s.leb_u())]
            while len(self.table) < off + len(fs):
                self.table.append(None)
            self.table[off:off + len(fs)] = fs

    def sec_10(self, s):
        for _ in range(s.leb_u()):
            size = s.leb_u()
            end = s.p + size
            locs = []
            for _ in range(s.leb_u()):
                n = s.leb_u()
                locs += [VT[s.u8()]] * n
            self.codes.append((locs, s.p, end))
            s.p = end

    def sec_11(self, s):
        for _ in range(s.leb_u()):
            assert s.leb_u() == 0
            off = self.const_expr(s)
            n = s.leb_u()
            self.datas.append((off, s.bytes(n)))


BIN = {  # opcode -> (type, C operator or template using a,b)
    0x6A: ("i32", "a + b"), 0x6B: ("i32", "a - b"), 0x6C: ("i32", "a * b"),
    0x6D: ("i32", "(uint32_t)((int32_t)a / (int32_t)b)"), 0x6E: ("i32", "a / b"),
    0x6F: ("i32", "(uint32_t)((int32_t)a % (int32_t)b)"), 0x70: ("i32", "a % b"),
    0x71: ("i32", "a & b"), 0x72: ("i32", "a | b"), 0x73: ("i32", "a ^ b"),
    0x74: ("i32", "a << (b & 31)"), 0x75: ("i32", "(uint32_t)((int32_t)a >> (b & 31))"),
    0x76: ("i32", "a >> (b & 31)"),
    0x77: ("i32", "(a << (b & 31)) | (a >> ((32 - (b & 31)) & 31))"),
    0x78: ("i32", "(a >> (b & 31)) | (a << ((32 - (b & 31)) & 31))"),
    0x7C: ("i64", "a + b"), 0x7D: ("i64", "a - b"), 0x7E: ("i64", "a * b"),
    0x7F: ("i64", "(uint64_t)((int64_t)a / (int64_t)b)"), 0x80: ("i64", "a / b"),
    0x81: ("i64", "(uint64_t)((int64_t)a % (int64_t)b)"), 0x82: ("i64", "a % b"),
    0x83: ("i64", "a & b"), 0x84: ("i64", "a | b"), 0x85: ("i64", "a ^ b"),
    0x86: ("i64", "a << (b & 63)"), 0x87: ("i64", "(uint64_t)((int64_t)a >> (b & 63))"),
    0x88: ("i64", "a >> (b & 63)"),
    0x89: ("i64", "(a << (b & 63)) | (a >> ((64 - (b & 63)) & 63))"),
    0x8A: ("i64", "(a >> (b & 63)) | (a << ((64 - (b & 63)) & 63))"),
}
CMP = {  # opcode -> (operand type, expr)
    0x46: ("i32", "a == b"), 0x47: ("i32", "a != b"),
    0x48: ("i32", "(int32_t)a < (int32_t)b"), 0x49: ("i32", "a < b"),
    0x4A: ("i32", "(int32_t)a > (int32_t)b"), 0x4B: ("i32", "a > b"),
    0x4C: ("i32", "(int32_t)a <= (int32_t)b"), 0x4D: ("i32", "a <= b"),
    0x4E: ("i32", "(int32_t)a >= (int32_t)b"), 0x4F: ("i32", "a >= b"),
    0x51: ("i64", "a == b"), 0x52: ("i64", "a != b"),
    0x53: ("i64", "(int64_t)a < (int64_t)b"), 0x54: ("i64", "a < b"),
    0x55: ("i64", "(int64_t)a > (int64_t)b"), 0x56: ("i64", "a > b"),
    0x57: ("i64", "(int64_t)a <= (int64_t)b"), 0x58: ("i64", "a <= b"),
    0x59: ("i64", "(int64_t)a >= (int64_t)b"), 0x5A: ("i64", "a >= b"),
}
LOAD = {  # opcode -> (result type, C memory type)
    0x28: ("i32", "uint32_t"), 0x29: ("i64", "uint64_t"),
    0x2C: ("i32", "int8_t"), 0x2D: ("i32", "uint8_t"), 0x2E: ("i32", "int16_t"), 0x2F: ("i32", "uint16_t"),
    0x30: ("i64", "int8_t"), 0x31: ("i64", "uint8_t"), 0x32: ("i64", "int16_t"), 0x33: ("i64", "uint16_t"),
    0x34: ("i64", "int32_t"), 0x35: ("i64", "uint32_t"),
}
STORE = {  # opcode -> (value type, C memory type)
    0x36: ("i32", "uint32_t"), 0x37: ("i64", "uint64_t"),
    0x3A: ("i32", "uint8_t"), 0x3B: ("i32", "uint16_t"),
    0x3C: ("i64", "uint8_t"), 0x3D: ("i64", "uint16_t"), 0x3E: ("i64", "uint32_t"),
}


class FuncGen:
    def __init__(self, mod, fidx, out):
        self.m, self.fidx, self.out = mod, fidx, out
        nimp = len(mod.imports)
        self.ps, self.rs = mod.types[mod.funcs[fidx - nimp]]
        self.locs, self.start, self.end = mod.codes[fidx - nimp]
        self.ltypes = self.ps + self.locs
        self.stack = []      # list of types; slot name derived from depth+type
        self.used = set()    # (depth, type) slots used
        self.body = []
        self.nlabel = 0
        self.ind = 1

    def emit(self, s):
        self.body.append("  " * self.ind + s)

    def slot(self, d, t):
        self.used.add((d, t))
        return "s%d%s" % (d, "i" if t == "i32" else "l")

    def push(self, t, expr):
        d = len(self.stack)
        self.stack.append(t)
        self.emit("%s = %s;" % (self.slot(d, t), expr))

    def pop(self, t=None):
        tt = self.stack.pop()
        if t is not None:
            assert tt == t, "type mismatch in func %d: %s vs %s" % (self.fidx, tt, t)
        return self.slot(len(self.stack), tt)

    def sig(self, ps, rs, name):
        args = ["W *w"] + ["%s p%d" % (CT[t], i) for i, t in enumerate(ps)]
        return "%s %s(%s)" % (CT[rs[0]] if rs else "void", name, ", ".join(args))

    def call_expr(self, name, ps, rs):
        args = [self.pop(t) for t in reversed(ps)][::-1]
        e = "%s(%s)" % (name, ", ".join(["w"] + args))
        if rs:
            self.push(rs[0], e)
        else:
            self.emit(e + ";")

    def gen(self):
        m = self.m
        r = Reader(m.raw, self.start, self.end)
        # control stack entries: [kind, label, stack height, label_used]
        ctl = [["func", "Lret", 0, False]]
        dead = 0  # >0: skipping unreachable code, counts nested blocks opened while dead
        while not r.eof():
            op = r.u8()
            if dead:
                # skip instruction, only track structure
                if op in (0x02, 0x03, 0x04):
                    r.u8()
                    dead += 1
                elif op == 0x0B:
                    dead -= 1
                    if dead == 0:
                        self.close_block(ctl)
                elif op == 0x05:
                    if dead == 1:
                        dead = 0
                        self.do_else(ctl)
                else:
                    self.skip_imm(r, op)
                continue
            if op == 0x00:
                self.emit("wasm_trap(w, 1);")
                dead = 1
                self.stack = self.stack[:ctl[-1][2]]
            elif op == 0x01:
                pass
            elif op in (0x02, 0x03):
                assert r.u8() == 0x40, "only void block types supported"
                self.nlabel += 1
                lab = "L%d" % self.nlabel
                if op == 0x03:
                    self.emit("%s:;" % lab)
                    ctl.append(["loop", lab, len(self.stack), True])
                else:
                    ctl.append(["block", lab, len(self.stack), False])
                self.emit("{")
                self.ind += 1
            elif op == 0x04:
                assert r.u8() == 0x40
                c = self.pop("i32")
                self.nlabel += 1
                ctl.append(["if", "L%d" % self.nlabel, len(self.stack), False])
                self.emit("if (%s) {" % c)
                self.ind += 1
            elif op == 0x05:
                self.do_else(ctl)
            elif op == 0x0B:
                self.close_block(ctl)
            elif op in (0x0C, 0x0D):
                depth = r.leb_u()
                tgt = ctl[-1 - depth]
                tgt[3] = True
                if tgt[0] == "func":
                    stmt = self.ret_stmt()
                else:
                    stmt = "goto %s;" % tgt[1]
                if op == 0x0D:
                    c = self.pop("i32")
                    self.emit("if (%s) %s" % (c, stmt))
                else:
                    self.emit(stmt)
                    dead = 1
                    self.stack = self.stack[:ctl[-1][2]]
            elif op == 0x0F:
                self.emit(self.ret_stmt())
                dead = 1
                self.stack = self.stack[:ctl[-1][2]]
            elif op == 0x10:
                f = r.leb_u()
                ps, rs = m.func_type(f)
                self.call_expr(m.cname(f), ps, rs)
            elif op == 0x11:
                ti = r.leb_u()
                assert r.u8() == 0
                ps, rs = m.types[ti]
                idx = self.pop("i32")
                args = [self.pop(t) for t in reversed(ps)][::-1]
                res = None
                if rs:
                    d = len(self.stack)
                    self.stack.append(rs[0])
                    res = self.slot(d, rs[0])
                self.emit("switch (%s) {" % idx)
                for ei, f in enumerate(m.table):
                    if f is None or m.func_type(f) != (ps, rs):
                        continue
                    call = "%s(%s)" % (m.cname(f), ", ".join(["w"] + args))
                    self.emit("  case %d: %s%s; break;" % (ei, (res + " = ") if res else "", call))
                self.emit("  default: wasm_trap(w, 2);")
                self.emit("}")
            elif op == 0x1A:
                self.pop()
            elif op == 0x1B:
                c = self.pop("i32")
                b = self.pop()
                t = self.stack[-1]
                a = self.pop(t)
                self.push(t, "%s ? %s : %s" % (c, a, b))
            elif op == 0x20:
                i = r.leb_u()
                self.push(self.ltypes[i], self.lname(i))
            elif op == 0x21:
                i = r.leb_u()
                self.emit("%s = %s;" % (self.lname(i), self.pop(self.ltypes[i])))
            elif op == 0x22:
                i = r.leb_u()
                t = self.ltypes[i]
                assert self.stack[-1] == t
                self.emit("%s = %s;" % (self.lname(i), self.slot(len(self.stack) - 1, t)))
            elif op in LOAD:
                r.leb_u()
                off = r.leb_u()
                t, ct = LOAD[op]
                a = self.pop("i32")
                self.push(t, "(%s)wasm_ld_%s(w, (uint64_t)%s + %du)" % (CT[t], ct, a, off))
            elif op in STORE:
                r.leb_u()
                off = r.leb_u()
                t, ct = STORE[op]
                v = self.pop(t)
                a = self.pop("i32")
                self.emit("wasm_st_%s(w, (uint64_t)%s + %du, (%s)%s);" % (ct, a, off, ct, v))
            elif op == 0x3F:
                r.u8()
                self.push("i32", "w->pages")
            elif op == 0x40:
                r.u8()
                n = self.pop("i32")
                self.push("i32", "wasm_grow(w, %s)" % n)
            elif op == 0x41:
                self.push("i32", "%du" % (r.leb_s(32) & 0xFFFFFFFF))
            elif op == 0x42:
                self.push("i64", "%dull" % (r.leb_s(64) & 0xFFFFFFFFFFFFFFFF))
            elif op == 0x45:
                self.push("i32", "(%s == 0)" % self.pop("i32"))
            elif op == 0x50:
                self.push("i32", "(%s == 0)" % self.pop("i64"))
            elif op in CMP:
                t, e = CMP[op]
                b = self.pop(t)
                a = self.pop(t)
                self.push("i32", "(" + self.subst(e, a, b) + ")")
            elif op in BIN:
                t, e = BIN[op]
                b = self.pop(t)
                a = self.pop(t)
                if op in (0x6D, 0x6E, 0x6F, 0x70, 0x7F, 0x80, 0x81, 0x82):
                    self.emit("if (%s == 0) wasm_trap(w, 3);" % b)
                self.push(t, self.subst(e, a, b))
            elif op in (0x67, 0x68, 0x69, 0x79, 0x7A, 0x7B):
                t = "i32" if op < 0x70 else "i64"
                fn = {0x67: "wasm_clz32", 0x68: "wasm_ctz32", 0x69: "__builtin_popcount",
                      0x79: "wasm_clz64", 0x7A: "wasm_ctz64", 0x7B: "__builtin_popcountll"}[op]
                self.push(t, "%s(%s)" % (fn, self.pop(t)))
            elif op == 0xA7:
                self.push("i32", "(uint32_t)%s" % self.pop("i64"))
            elif op == 0xAC:
                self.push("i64", "(uint64_t)(int64_t)(int32_t)%s" % self.pop("i32"))
            elif op == 0xAD:
                self.push("i64", "(uint64_t)%s" % self.pop("i32"))
            else:
                raise NotImplementedError("opcode 0x%02x in func %d" % (op, self.fidx))
        assert not ctl, "unbalanced control stack in func %d" % self.fidx
        # assemble
        o = self.out
        o.append(self.sig(self.ps, self.rs, m.cname(self.fidx)) + " {")
        for i, t in enumerate(self.locs):
            o.append("  %s l%d = 0;" % (CT[t], i + len(self.ps)))
        for d, t in sorted(self.used):
            o.append("  %s %s;" % (CT[t], self.slot(d, t)))
        o.extend(self.body)
        o.append("}")
        o.append("")

    @staticmethod
    def subst(e, a, b):
        return e.replace("a", "\0").replace("b", b).replace("\0", a)

    def lname(self, i):
        return ("p%d" if i < len(self.ps) else "l%d") % i

    def ret_stmt(self):
        if self.rs:
            return "return %s;" % self.slot(len(self.stack) - 1, self.rs[0])
        return "return;"

    def do_else(self, ctl):
        top = ctl[-1]
        assert top[0] == "if"
        self.stack = self.stack[:top[2]]
        self.ind -= 1
        self.emit("} else {")
        self.ind += 1
        top[0] = "else"

    def close_block(self, ctl):
        top = ctl.pop()
        if top[0] == "func":
            if self.rs and self.stack:
                self.emit(self.ret_stmt())
            return
        self.stack = self.stack[:top[2]]
        self.ind -= 1
        self.emit("}")
        if top[0] != "loop" and top[3]:
            self.emit("%s:;" % top[1])

    def skip_imm(self, r, op):
        if op in (0x0C, 0x0D, 0x10, 0x20, 0x21, 0x22, 0x23, 0x24):
            r.leb_u()
        elif op == 0x11:
            r.leb_u()
            r.u8()
        elif op == 0x0E:
            for _ in range(r.leb_u() + 1):
                r.leb_u()
        elif 0x28 <= op <= 0x3E:
            r.leb_u()
            r.leb_u()
        elif op in (0x3F, 0x40):
            r.u8()
        elif op == 0x41:
            r.leb_s(32)
        elif op == 0x42:
            r.leb_s(64)


def translate(wasm_bytes):
    m = Module(wasm_bytes)
    m.raw = wasm_bytes
    nimp = len(m.imports)

    def func_type(f):
        return m.types[m.imports[f][2]] if f < nimp else m.types[m.funcs[f - nimp]]

    def cname(f):
        if f < nimp:
            return "wi_%s_%s" % (m.imports[f][0], m.imports[f][1])
        return "f%d" % f

    m.func_type, m.cname = func_type, cname
    out = ['#include "wasm_rt.h"', ""]
    for f in range(nimp):
        ps, rs = func_type(f)
        out.append("extern " + FuncGen.sig(None, ps, rs, cname(f)) + ";")
    for f in range(nimp, nimp + len(m.funcs)):
        ps, rs = func_type(f)
        nm = m.func_names.get(f)
        out.append("static " + FuncGen.sig(None, ps, rs, cname(f)) + ";" + ("  /* %s */" % nm if nm else ""))
    out.append("")
    for f in range(nimp, nimp + len(m.funcs)):
        body = []
        FuncGen(m, f, body).gen()
        body[0] = "static " + body[0]
        out.extend(body)
    # data segments
    out.append("static const struct { uint32_t off, len; const unsigned char *p; } wasm_data[] = {")
    for off, b in m.datas:
        lit = "".join("\\x%02x" % c for c in b)
        out.append('  {%du, %du, (const unsigned char *)"%s"},' % (off, len(b), lit))
    out.append("};")
    out.append("void wasm_instantiate(W *w) {")
    out.append("  wasm_alloc(w, %du);" % m.mem_pages[0])
    out.append("  for (unsigned i = 0; i < sizeof(wasm_data) / sizeof(wasm_data[0]); i++)")
    out.append("    memcpy(w->mem + wasm_data[i].off, wasm_data[i].p, wasm_data[i].len);")
    out.append("}")
    for nm, kind, idx in m.exports:
        if kind != 0:
            continue
        ps, rs = func_type(idx)
        args = ", ".join(["w"] + ["p%d" % i for i in range(len(ps))])
        out.append(FuncGen.sig(None, ps, rs, "wx_" + nm) + " { %s%s(%s); }" % ("return " if rs else "", cname(idx), args))
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    with open(src, "rb") as f:
        c = translate(f.read())
    with open(dst, "w") as f:
        f.write(c)
