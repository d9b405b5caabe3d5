"""blake3_ref.py -- plain u32 BLAKE3 compression, the KAT oracle for out[16].  TEST INFRASTRUCTURE ONLY.

Restates the reference's JS test oracle /root/reference/test/blake3_utils/compressions.js
(g :10-20, round :22-37, permute :39-45, compress :64-109), which works on bit strings; here on ints.
The reference tests compare only the 16 output words against it (test/blake3_hash.test.ts:30-59).
"""
IV = [0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19]
MSG_PERMUTATION = [2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8]     # compressions.js:58
M = 0xFFFFFFFF


def _rotr(x, n):
    return ((x >> n) | (x << (32 - n))) & M


def g(s, a, b, c, d, mx, my):                      # compressions.js:10-20
    s[a] = (s[a] + s[b] + mx) & M
    s[d] = _rotr(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & M
    s[b] = _rotr(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b] + my) & M
    s[d] = _rotr(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & M
    s[b] = _rotr(s[b] ^ s[c], 7)


def round_(s, m):                                  # compressions.js:22-37
    g(s, 0, 4, 8, 12, m[0], m[1]); g(s, 1, 5, 9, 13, m[2], m[3])
    g(s, 2, 6, 10, 14, m[4], m[5]); g(s, 3, 7, 11, 15, m[6], m[7])
    g(s, 0, 5, 10, 15, m[8], m[9]); g(s, 1, 6, 11, 12, m[10], m[11])
    g(s, 2, 7, 8, 13, m[12], m[13]); g(s, 3, 4, 9, 14, m[14], m[15])


def compress(h, m, t0, t1, b, d):                  # compressions.js:64-109
    s = list(h) + IV[:4] + [t0, t1, b, d]
    m = list(m)
    for r in range(7):
        round_(s, m)
        if r < 6:
            m = [m[MSG_PERMUTATION[i]] for i in range(16)]
    return [s[i] ^ s[i + 8] for i in range(8)] + [s[i + 8] ^ h[i] for i in range(8)]


class LCG:                                         # test/utils.ts:4-21
    def __init__(self, seed):
        self.seed = seed

    def next(self):
        self.seed = (1664525 * self.seed + 1013904223) % 2 ** 32
        return self.seed


def gen_random_chunk(lcg, b=64, d=0, t0=0, t1=0, h=IV):   # test/utils.ts:34-56
    assert b % 4 == 0 and b <= 64
    lcg.next()
    m = [lcg.next() for _ in range((b + 3) // 4)] + [0] * (16 - b // 4)
    return {"h": list(h), "m": m, "b": b, "d": d, "t": [t0, t1]}
