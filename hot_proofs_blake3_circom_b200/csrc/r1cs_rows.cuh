// r1cs_rows.cuh -- GENERAL row evaluators of the stand-alone R1CS check:  (A.z) * (B.z) == C.z  for rows whose shape the
// compiled fast program (kernels_r1cs_fast.cuh) does not cover -- in practice rows of a loaded iden3 `.r1cs` file
// (b3w_r1cs_load) with coefficients that are arbitrary field elements (e.g. circom's own O2 output: 2^-31 mod p).
// Rows are grouped in shape classes (r1cs.cuh) and cut into row BLOCKS by the host (stg_blockify, r1cs_load.h); a value
// source (CompactSrc, kernels_r1cs_fast.cuh) hands out the witness values as tagged 8-byte words.  Exact for ANY values
// and coefficients: signed 64/128-bit integers whenever every term is small, Montgomery arithmetic in Fr otherwise.
// Included by blake3wit.cu only, after r1cs.cuh.
#pragma once

#define STG_TAG_NEG (1ull << 62)
#define STG_TAG_BIG (2ull << 62)
#define STG_PAYLOAD ((1ull << 62) - 1)
#define B3W_NOT_CANONICAL 0xFFFFFFFEu /* first_bad: a slot holds a value >= p */
#define R1CS_FLAG_COEF64 128u /* slot-space sets: every coefficient of the class fits int64 (set by stg_blockify) */
#define R1CS_FLAG_MATRIX 256u /* slot-space sets: `terms` holds the [term][row] matrix of this class, not row blocks */
#define R1CS_FLAG_BIGCOEF 64u /* loaded sets only: coefficients are full field elements, stored per class in coef_fr */
#define R1CS_FLAG_FAST64 1024u /* slot-space sets: eligible for the 64-bit evaluator below (set by stg_blockify) */
#define R1CS_FLAG_XORROW 2048u /* slot-space sets: every row is  (a x)(b y) = k x + k y - k o  with  a b = 2 k: for bits, o = x xor y (set by stg_blockify) */
#define R1CS_FLAG_BOOLROW 512u /* slot-space sets: every row of the class is  (a x) * (b x - b w0) = 0, i.e. "x is 0 or 1" (set by stg_blockify) */
// The 64-bit fast path.  Almost every term of these systems is (small coefficient) x (word) or (power of two) x (bit).
// With a value below 2^40 in magnitude (STG_FAST_VMAX, tested per term), a term whose coefficient is at
// most 2^16 in magnitude is below 2^56, and a term with a larger coefficient (< 2^56: classes flagged FAST64 by the host)
// is below 2^56 too PROVIDED its value is 0 or 1 -- which is tested per term.  A linear combination has at most
// 64 terms, so it stays below 2^62 and plain int64 arithmetic is exact; a row that fails the per-term test (or meets a
// genuine field element) is re-evaluated exactly in Fr.  One 128-bit product per row decides A*B == C.
#define STG_FAST_VMAX 40
#define STG_FAST_COEF 16

// header words are streamed (each is used once per instance): keep them out of L1, where the coefficients live
__device__ __forceinline__ uint32_t stg_ld_stream(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ fr_t fr_from_i128(i128 x, const fr_t &p) {
  const bool neg = x < 0;
  const unsigned __int128 m = neg ? (unsigned __int128)(-x) : (unsigned __int128)x;
  fr_t r = fr_zero();
  r.l[0] = (uint32_t)m; r.l[1] = (uint32_t)(m >> 32); r.l[2] = (uint32_t)(m >> 64); r.l[3] = (uint32_t)(m >> 96);
  return neg ? fr_neg(r, p) : r;
}

__device__ __forceinline__ i128 staged_coef(const r1cs_class_dev &c, const r1cs_tables_dev &T, uint32_t t, uint32_t r) {
  const uint32_t ci = (c.flags & R1CS_FLAG_ROWCOEF) ? c.coef_off + t * c.count + r : c.coef_off + t;
  return ((i128)T.coef_hi[ci] << 64) | (i128)(uint64_t)T.coef_lo[ci];
}

// coefficient * value of one term, exact in 128 bits.  COEF64 classes (every coefficient fits int64: all but the
// 2^64 of Num2Bits(65)) need one 64x64->128 multiply; `ok` turns false when the slot holds a genuine field element.
template <class Src>
__device__ __forceinline__ i128 staged_term(const Src &src, const r1cs_class_dev &c, const r1cs_tables_dev &T, uint32_t t, uint32_t r,
                                           uint32_t wire, bool &ok) {
  const uint64_t x = src.get(wire);
  ok = ok && !(x & STG_TAG_BIG);
  const int64_t v = (x & STG_TAG_NEG) ? -(int64_t)(x & STG_PAYLOAD) : (int64_t)(x & STG_PAYLOAD);
  const uint32_t ci = (c.flags & R1CS_FLAG_ROWCOEF) ? c.coef_off + t * c.count + r : c.coef_off + t;
  if (c.flags & R1CS_FLAG_COEF64) return (i128)T.coef_lo[ci] * (i128)v;
  return (((i128)T.coef_hi[ci] << 64) | (i128)(uint64_t)T.coef_lo[ci]) * (i128)v;
}

// 64-bit term (see STG_FAST_*): `slow` turns true when the row has to be re-evaluated exactly
template <class Src>
__device__ __forceinline__ int64_t staged_term64(const Src &src, const r1cs_class_dev &c, const r1cs_tables_dev &T, uint32_t t, uint32_t r,
                                                uint32_t wire, bool &slow) {
  const uint64_t x = src.get(wire);
  const uint64_t mag = x & STG_PAYLOAD;
  const int64_t v = (x & STG_TAG_NEG) ? -(int64_t)mag : (int64_t)mag;
  const uint32_t ci = (c.flags & R1CS_FLAG_ROWCOEF) ? c.coef_off + t * c.count + r : c.coef_off + t;
  const int64_t co = T.coef_lo[ci];
  const uint64_t aco = co < 0 ? (uint64_t)0 - (uint64_t)co : (uint64_t)co;
  slow = slow || (x & STG_TAG_BIG) != 0 || (mag >> STG_FAST_VMAX) != 0 || ((aco >> STG_FAST_COEF) != 0 && mag > 1);
  return co * v;
}
__device__ __forceinline__ bool staged_verdict64(const r1cs_class_dev &c, int64_t L0, int64_t L1, int64_t L2) {
  if (c.nA == 0 || c.nB == 0) return L2 == 0;
  return (i128)L0 * (i128)L1 == (i128)L2;
}

// verdict of one row from its three exact linear combinations (integer path), with the Fr fallback left to the caller
__device__ __forceinline__ bool staged_int_verdict(const r1cs_class_dev &c, i128 L0, i128 L1, i128 L2, bool &need_fr) {
  const i128 lim = (i128)1 << 62;
  need_fr = false;
  if (c.nA == 0 || c.nB == 0) return L2 == 0;
  if (L0 > -lim && L0 < lim && L1 > -lim && L1 < lim) return L0 * L1 == L2;
  need_fr = true;
  return false;
}

// Row BLOCKS.  The term columns of a class are almost everywhere arithmetic progressions (32 booleanity rows over 32
// consecutive bit slots, the same gadget row in consecutive gadget instances, ...), so instead of one table entry per
// term per row -- 470 KB per compression witness, streamed from L2 for every instance -- the host cuts each class into
// blocks of <= 32 consecutive rows in which every column is affine, and stores per block only
//   { first row, rows, then per term { wire of the first row, wire step per row } }
// (~100 KB per set, read as warp-uniform loads: one block = one warp step, lane = row).  Irregular systems degrade
// gracefully to short blocks.  Built by stg_blockify() on the host for the built-in sets and for loaded .r1cs files alike.
__device__ __forceinline__ uint32_t stg_wire(const uint32_t *hdr, uint32_t t, uint32_t lane) {
  const uint2 bd = __ldg(reinterpret_cast<const uint2 *>(hdr + 2) + t);
  return bd.x + lane * bd.y;
}

// wire of term t of row r: from the block header (hdr != NULL, lane = row inside the block) or from the term matrix
__device__ __forceinline__ uint32_t stg_row_wire(const r1cs_class_dev &c, const r1cs_tables_dev &T, const uint32_t *hdr, uint32_t lane,
                                                uint32_t t, uint32_t r) {
  return hdr ? stg_wire(hdr, t, lane) : __ldg(T.terms + c.term_off + t * c.count + r);
}

// exact evaluation of one row in Fr (slow path: a term is a genuine field element, the coefficients are, or the integer
// product could overflow)
template <class Src>
__device__ __noinline__ bool staged_row_fr(const Src &src, const r1cs_class_dev &c, const r1cs_tables_dev &T, const uint32_t *hdr,
                                           uint32_t lane, uint32_t r) {
  const field_consts &F = *src.F;
  const uint32_t n[3] = {c.nA, c.nB, c.nC};
  fr_t L[3];
  uint32_t t = 0;
  for (int part = 0; part < 3; part++) {
    fr_t acc = fr_zero();
    for (uint32_t j = 0; j < n[part]; j++, t++) {
      const fr_t v = src.field(stg_row_wire(c, T, hdr, lane, t, r));
      fr_t co;
      if (c.flags & R1CS_FLAG_BIGCOEF) co = T.coef_fr[(c.flags & R1CS_FLAG_ROWCOEF) ? c.coef_off + t * c.count + r : c.coef_off + t];
      else co = fr_from_i128(staged_coef(c, T, t, r), F.p);
      const fr_t prod = fr_montmul(fr_montmul(co, F.r2, F.p, F.n0), v, F.p, F.n0);     // co * v
      acc = fr_add(acc, prod, F.p);
    }
    L[part] = acc;
  }
  fr_t lhs = fr_zero();
  if (c.nA && c.nB) lhs = fr_montmul(fr_montmul(L[0], L[1], F.p, F.n0), F.r2, F.p, F.n0);
  bool eq = true;
#pragma unroll
  for (int j = 0; j < 8; j++) eq = eq && (lhs.l[j] == L[2].l[j]);
  return eq;
}

// one block: lane = row.  The header is fetched 32 words at a time by the whole warp (one coalesced L2 access per 16
// terms, the next chunk already in flight) and handed round with shuffles.  Returns the violated row's id (class order,
// or the file's constraint index) or B3W_NO_ROW.
template <bool FAST, class Src>
__device__ __forceinline__ uint32_t staged_block(const Src &src, const r1cs_class_dev &c, const r1cs_tables_dev &T, const uint32_t *hdr,
                                                uint32_t lane) {
  typedef typename std::conditional<FAST, int64_t, i128>::type acc_t;
  const uint32_t nt = (uint32_t)c.nA + c.nB + c.nC, hw = 2u + 2u * nt, nchunks = (hw + 31u) >> 5;
  uint32_t h = lane < hw ? Src::ld_table(hdr + lane) : 0u;
  const uint32_t row0 = __shfl_sync(0xffffffffu, h, 0), len = __shfl_sync(0xffffffffu, h, 1);
  const bool active = lane < len;
  const uint32_t ln = active ? lane : 0u;                   // idle lanes shadow row 0 of the block: every load stays in range
  const uint32_t r = row0 + ln;
  if (FAST && (c.flags & R1CS_FLAG_BOOLROW)) {              // "x is 0 or 1": the wire of the A term is all that matters
    const uint32_t base = __shfl_sync(0xffffffffu, h, 2), step = __shfl_sync(0xffffffffu, h, 3);
    if (!active || src.get(base + ln * step) < 2ull) return B3W_NO_ROW;
    return T.row_ids ? T.row_ids[c.row_off + r] : c.row_off + r;
  }
  if (FAST && (c.flags & R1CS_FLAG_XORROW)) {               // terms: x | y | the three of {x, y, o} in some order (hw = 12)
    uint32_t wv[5];
#pragma unroll
    for (int t = 0; t < 5; t++) wv[t] = __shfl_sync(0xffffffffu, h, 2 + 2 * t) + ln * __shfl_sync(0xffffffffu, h, 3 + 2 * t);
    const uint64_t vx = src.get(wv[0]), vy = src.get(wv[1]), vo = src.get(wv[2] ^ wv[3] ^ wv[4] ^ wv[0] ^ wv[1]);
    bool holds = (vx | vy | vo) < 2ull ? vo == (vx ^ vy) : staged_row_fr(src, c, T, hdr, ln, r);      // not all bits: exact evaluation
    if (holds || !active) return B3W_NO_ROW;
    return T.row_ids ? T.row_ids[c.row_off + r] : c.row_off + r;
  }
  acc_t L0 = 0, L1 = 0, acc = 0;
  bool ok = !(c.flags & R1CS_FLAG_BIGCOEF), slow = false;
  const uint32_t nAB = (uint32_t)c.nA + c.nB;
  for (uint32_t k = 0; k < nchunks; k++) {
    const uint32_t nxt = 32u * (k + 1u) + lane;
    const uint32_t hn = (k + 1u < nchunks && nxt < hw) ? Src::ld_table(hdr + nxt) : 0u;
    const uint32_t t_lo = k == 0 ? 0u : 16u * k - 1u, t_hi = min(nt, 16u * k + 15u);
    for (uint32_t t = t_lo; t < t_hi; t++) {
      const uint32_t l = (2u + 2u * t) & 31u;
      const uint32_t base = __shfl_sync(0xffffffffu, h, l), step = __shfl_sync(0xffffffffu, h, l + 1u);
      if (t == c.nA) { L0 = acc; acc = 0; }               // part boundaries are warp-uniform
      if (t == nAB) { L1 = acc; acc = 0; }
      if (FAST) acc += (acc_t)staged_term64(src, c, T, t, r, base + ln * step, slow);
      else acc += (acc_t)staged_term(src, c, T, t, r, base + ln * step, ok);
    }
    h = hn;
  }
  if (nt == c.nA) { L0 = acc; acc = 0; }
  if (nt == nAB) { L1 = acc; acc = 0; }
  bool need_fr, holds;
  if (FAST) {
    need_fr = slow;
    holds = !slow && staged_verdict64(c, (int64_t)L0, (int64_t)L1, (int64_t)acc);
  } else {
    need_fr = !ok;
    holds = ok && staged_int_verdict(c, (i128)L0, (i128)L1, (i128)acc, need_fr);
  }
  if (need_fr && active) holds = staged_row_fr(src, c, T, hdr, ln, r);
  if (holds || !active) return B3W_NO_ROW;
  return T.row_ids ? T.row_ids[c.row_off + r] : c.row_off + r;
}

// MATRIX classes: rows whose columns are not affine over consecutive rows (blocks would hold ~2 rows: the 33..35-term
// bit recompositions, whose gadget instances are unevenly spaced in the witness) keep one table entry per term per row;
// lane = row, 32 rows per warp step whatever their wires are.
template <bool FAST, class Src>
__device__ __forceinline__ uint32_t staged_matrix_row(const Src &src, const r1cs_class_dev &c, const r1cs_tables_dev &T, uint32_t r0) {
  typedef typename std::conditional<FAST, int64_t, i128>::type acc_t;
  const bool active = r0 < c.count;
  const uint32_t r = active ? r0 : c.count - 1u;
  const uint32_t *m = T.terms + c.term_off + r;
  const uint32_t nt = (uint32_t)c.nA + c.nB + c.nC, nAB = (uint32_t)c.nA + c.nB;
  acc_t L0 = 0, L1 = 0, acc = 0;
  bool ok = !(c.flags & R1CS_FLAG_BIGCOEF), slow = false;
  uint32_t w = Src::ld_table(m);
  if (FAST && (c.flags & R1CS_FLAG_BOOLROW)) {
    if (!active || src.get(w) < 2ull) return B3W_NO_ROW;
    return T.row_ids ? T.row_ids[c.row_off + r] : c.row_off + r;
  }
  for (uint32_t t = 0; t < nt; t++) {
    const uint32_t wn = t + 1u < nt ? Src::ld_table(m + (size_t)(t + 1u) * c.count) : 0u;      // next term's wire in flight
    if (t == c.nA) { L0 = acc; acc = 0; }
    if (t == nAB) { L1 = acc; acc = 0; }
    if (FAST) acc += (acc_t)staged_term64(src, c, T, t, r, w, slow);
    else acc += (acc_t)staged_term(src, c, T, t, r, w, ok);
    w = wn;
  }
  if (nt == c.nA) { L0 = acc; acc = 0; }
  if (nt == nAB) { L1 = acc; acc = 0; }
  bool need_fr, holds;
  if (FAST) {
    need_fr = slow;
    holds = !slow && staged_verdict64(c, (int64_t)L0, (int64_t)L1, (int64_t)acc);
  } else {
    need_fr = !ok;
    holds = ok && staged_int_verdict(c, (i128)L0, (i128)L1, (i128)acc, need_fr);
  }
  if (need_fr && active) holds = staged_row_fr(src, c, T, nullptr, 0, r);
  if (holds || !active) return B3W_NO_ROW;
  return T.row_ids ? T.row_ids[c.row_off + r] : c.row_off + r;
}
