#!/usr/bin/env python3
"""make_golden_nova_wide.py -- regenerates tests/golden/nova_wide_cases.npz (run HERE, where /root/reference exists).

Inputs of the three nova witness programs OUTSIDE the u32 domain, run through the reference's own wasm (Oracle A):
field-valued n_blocks / block_count / depths, a split chunk index, wide message words, and inputs on which the
reference throws "Assert Failed." (with the text it prints).  Per variant v in (nova_bn_o2, nova_pasta_o2, nova_bn_o1):
  <v>_fr (n, 32, 32) u8 | <v>_status (n,) i32 | <v>_text (n,) S | <v>_valid (k,) i64 | <v>_witness (k, ws*32) u8
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_wasm import RefWasm  # noqa: E402

VARIANTS = ("nova_bn_o2", "nova_pasta_o2", "nova_bn_o1")


def as_input(v):
    return {"n_blocks": v[0], "block_count": v[1], "h": v[2:10], "chunk_idx_low": v[10], "chunk_idx_high": v[11],
            "leaf_depth": v[12], "total_depth": v[13], "depth": v[14], "m": v[15:31], "b": v[31]}


def cases(p, seed):
    rnd = random.Random(seed)

    def step(leaf=None, depth=None):
        leaf = leaf or rnd.randrange(1, 65)
        depth = rnd.randrange(leaf) if depth is None else depth
        nb = rnd.randrange(1, 17)
        return [nb, rnd.randrange(nb)] + [rnd.randrange(2**32) for _ in range(8)] + \
               [rnd.randrange(2**32), rnd.randrange(2**32), leaf, leaf, depth] + [rnd.randrange(2**32) for _ in range(16)] + \
               [rnd.randrange(65)]
    out = []
    X = rnd.randrange(p)
    v = step(); v[0] = 2**40; out.append(v)                                   # n_blocks: any field element
    v = step(); v[0] = p - 3; v[1] = p - 4; out.append(v)                      # ... and block_count == n_blocks - 1 in the field
    v = step(); v[1] = X; out.append(v)
    v = step(); v[14] = X; v[12] = (X + 1) % p; v[13] = (X + 1) % p; out.append(v)         # leaf step at a field-valued depth
    v = step(); v[14] = X; v[12] = (X + 256) % p; v[13] = (X + 40) % p; out.append(v)      # parent, eqs[38] fires
    v = step(); v[14] = X; v[12] = (X + 257) % p; out.append(v)                            # Num2Bits(9) fails
    v = step(); v[14] = 5; v[12] = 5; out.append(v)                                        # exceed_depth
    v = step(); v[13] = X; out.append(v)                                                   # total_depth unrelated to depth
    v = step(leaf=9, depth=3); v[10] = 2**64 + 77; v[11] = 0; out.append(v)                # parent: chunk_idx_low alone >= 2^64
    v = step(leaf=9, depth=3); v[10] = (p - 5 * 2**32) % p; v[11] = 5; out.append(v)       # parent: low + 2^32 high == 0 mod p
    v = step(leaf=9, depth=3); v[11] = 2**33; out.append(v)                                # Num2Bits(65) fails
    v = step(leaf=9, depth=8); v[10] = 2**32; out.append(v)                                # leaf: t[0] >= 2^32 asserts in the compression
    v = step(leaf=9, depth=3); v[2] = X; v[25] = p - 1; out.append(v)                      # parent: h and m[8..15] do not reach the compression
    v = step(leaf=9, depth=8); v[2] = 2**32 + 5; out.append(v)                             # leaf: h[0] asserts at the output xor
    v = step(leaf=9, depth=8); v[6] = 2**32 + 5; out.append(v)                             # leaf: h[4] asserts in round 0
    v = step(leaf=9, depth=8); v[15] = 2**32; v[20] = p - 1; out.append(v)                 # leaf: wide message words, valid
    v = step(leaf=9, depth=3); v[15] = 2**32 + 9; out.append(v)                            # parent: wide sibling word
    v = step(leaf=9, depth=8); v[31] = 2**32; out.append(v)                                # b
    for it in range(14):
        v = step()
        for k in rnd.sample(range(32), rnd.randrange(1, 4)):
            v[k] = rnd.choice([rnd.randrange(p), 2**32 + rnd.randrange(2**20), p - 1 - rnd.randrange(2**20)])
        if it % 2:
            Y = rnd.randrange(p)
            d = rnd.randrange(1, 200)
            v[14], v[12], v[13] = Y, (Y + d) % p, (Y + rnd.randrange(1, 70)) % p
        out.append(v)
    return out


def main():
    arrays = {}
    for variant in VARIANTS:
        ref = RefWasm(variant)
        p = ref.prime
        vals = cases(p, 0xB3B30009)
        n = len(vals)
        fr = np.zeros((n, 32, 32), np.uint8)
        status = np.zeros(n, np.int32)
        text, wit, valid = [], [], []
        for i, v in enumerate(vals):
            for k, x in enumerate(v):
                fr[i, k] = np.frombuffer(int(x % p).to_bytes(32, "little"), np.uint8)
            rc, w = ref.calculate(as_input(v))
            status[i] = rc
            text.append(ref.err_msg().encode() if rc else b"")
            if rc == 0:
                valid.append(i)
                wit.append(w)
        arrays.update({variant + "_fr": fr, variant + "_status": status, variant + "_text": np.array(text),
                       variant + "_valid": np.array(valid, np.int64), variant + "_witness": np.stack(wit)})
        print(variant, n, "cases,", len(valid), "valid,", int((status == 4).sum()), "assert")
    np.savez_compressed(os.path.join(HERE, "nova_wide_cases.npz"), **arrays)


if __name__ == "__main__":
    main()
