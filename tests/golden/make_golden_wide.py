#!/usr/bin/env python3
"""make_golden_wide.py -- regenerates tests/golden/compression_wide_cases.npz (run HERE, where /root/reference exists).

Inputs of blake3_compression OUTSIDE the u32 domain, run through the reference's own witness program (Oracle A,
oracle/_ref): message words beyond 2^32 / negative (valid witnesses), and inputs on which the
reference throws "Assert Failed." -- with the per-template trace it prints (witness_calculator.js:21-43).
  fr       (n, 28, 32) u8   the inputs as little-endian field elements, declaration order h[8] m[16] t[2] b d
  status   (n,)        i32  0 or 4 (the wasm's exceptionHandler code)
  text     (n,)        S    the printErrorMessage lines ("" when status is 0)
  valid    (k,)        i64  indices of the instances with status 0
  witness  (k, 24093*32) u8 their witnesses
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_wasm import RefWasm  # noqa: E402
from oracle.blake3_ref import LCG, gen_random_chunk  # noqa: E402


def cases(p):
    c = gen_random_chunk(LCG(6429))
    base = c["h"] + c["m"] + c["t"] + [c["b"], c["d"]]
    out = []

    def with_(**kw):
        v = list(base)
        for k, x in kw.items():
            v[int(k[1:])] = x % p
        out.append(v)
    # the cases SURVEY.md 8(a) A8 names, and the edges of the 34-bit window
    with_(i8=2**32); with_(i8=p - 1); with_(i26=2**33)
    with_(i8=2**33); with_(i8=2**34 - 1); with_(i8=p - 2**32); with_(i13=p - 2**33); with_(i8=2**100)
    for j in range(16):
        with_(**{"i%d" % (8 + j): 2**32 + j})
    for j in range(0, 16, 3):
        with_(**{"i%d" % (8 + j): -1 - j})
    for k in (0, 3, 4, 7, 24, 25, 26, 27):
        with_(**{"i%d" % k: 2**32 + 5})
    with_(i0=2**100); with_(i0=p - 5, i8=2**32 + 77)
    rnd = random.Random(0xB3B30008)

    def wide_val():
        k = rnd.randrange(7)
        if k == 0:
            return 2**32 + rnd.randrange(2**32)
        if k == 1:
            return p - 1 - rnd.randrange(2**20)
        if k == 2:
            return p - rnd.randrange(1, 2**32)
        if k == 3:
            return rnd.randrange(2**32, 2**33)
        if k == 4:
            return rnd.randrange(p)
        if k == 5:
            return 2**34 - 1 - rnd.randrange(2**31)
        return 2**33 + rnd.randrange(-5, 5)
    for it in range(40):
        v = [rnd.randrange(2**32) for _ in range(28)]
        v[26], v[27] = rnd.randrange(65), rnd.randrange(16)
        mode = it % 5
        if mode <= 2:
            for j in rnd.sample(range(16), rnd.randrange(1, 4)):
                v[8 + j] = rnd.choice([2**32 + rnd.randrange(2**31), p - rnd.randrange(1, 2**31)]) if mode else wide_val()
        elif mode == 3:
            v[rnd.choice(list(range(8)) + [24, 25, 26, 27])] = wide_val()
        else:
            g = rnd.randrange(4)
            x = rnd.randrange(p)
            v[g], v[8 + 2 * g] = x, (p - x + rnd.randrange(2**32)) % p
        out.append(v)
    return out


def main():
    ref = RefWasm("compression")
    p = ref.prime
    vals = cases(p)
    n = len(vals)
    fr = np.zeros((n, 28, 32), np.uint8)
    status = np.zeros(n, np.int32)
    text, wit, valid = [], [], []
    for i, v in enumerate(vals):
        for k, x in enumerate(v):
            fr[i, k] = np.frombuffer(int(x % p).to_bytes(32, "little"), np.uint8)
        rc, w = ref.calculate({"h": v[0:8], "m": v[8:24], "t": v[24:26], "b": v[26], "d": v[27]})
        status[i] = rc
        text.append(ref.err_msg().encode() if rc else b"")
        if rc == 0:
            valid.append(i)
            wit.append(w)
    np.savez_compressed(os.path.join(HERE, "compression_wide_cases.npz"), fr=fr, status=status, text=np.array(text),
                        valid=np.array(valid, np.int64), witness=np.stack(wit))
    print("compression_wide_cases.npz:", n, "cases,", len(valid), "valid,", int((status == 4).sum()), "assert")


if __name__ == "__main__":
    main()
