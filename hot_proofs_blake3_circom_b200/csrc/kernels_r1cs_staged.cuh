// kernels_r1cs_staged.cuh -- the GENERAL stand-alone R1CS check:  (A.z) * (B.z) == C.z  for every row of a constraint
// system loaded at run time from an iden3 `.r1cs` file (b3w_r1cs_load: the artefact format of the reference's build/,
// e.g. rust_fold/src/blake3_circuit.rs:71-81 `CircomConfig::new(wasm, r1cs)`), over witnesses resident in HBM.
// (The built-in template-derived rows, whose value kinds are known offline, keep the lighter one-warp-per-instance
// evaluator k_r1cs_check_witness in kernels_aux.cuh / r1cs.cuh.)
//
// One CTA per instance.  Phase 1 streams the instance's witness (771 KB) from HBM exactly once, coalesced, and keeps
// a compact copy in shared memory: 8 bytes per slot (62-bit magnitude + tag: small non-negative / small negative,
// i.e. the slot holds p - k / genuine field element = index into a side table of full 256-bit values); a slot >= p
// is reported as non-canonical.  Phase 2 evaluates the rows from shared memory: exact signed 128-bit integers whenever
// every term is small, Montgomery arithmetic in Fr otherwise (field-valued slots, coefficients that are not small
// integers, products that could overflow) -- so ANY satisfied row is accepted and any violated row rejected, whatever
// the values and coefficients are.  Measured on B200 (profiles/): phase 1 alone 6.9 M witnesses/s (5.4 TB/s of reads);
// with phase 2, 0.9 M/s with 128-bit row arithmetic throughout, 1.22 M/s (compression, 24 544 rows) since rows of small
// coefficients run in 64-bit arithmetic and booleanity rows are a single comparison (STG_FAST_*, BOOLROW).  What is left
// is latency: the witness copy takes 209 KB of shared memory, so one CTA per SM has to hide the L2 latency of the block
// headers / term matrices (~1 200 warp steps per instance over 32 warps) by itself; rows that touch genuine field
// elements (nova: ~200 slots) pay the generic Fr fallback (nova O1: 0.41 M/s).
// Included by blake3wit.cu only, after r1cs.cuh.
#pragma once

#define STG_THREADS 1024
#ifndef STG_EXP_SKIP_P2
#define STG_EXP_SKIP_P2 0          /* experiment builds only */
#endif
#define STG_MAX_BIG 512       /* field-valued slots per witness the side table holds (nova O1: < 200) */
#define STG_MAX_CLASSES 96    /* shape classes per set (built-in: <= 27); more -> b3w_r1cs_load refuses */
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

struct StagedSrc {
  static __device__ __forceinline__ uint32_t ld_table(const uint32_t *p) { return stg_ld_stream(p); }     // row-block headers / term matrices
  const uint64_t *val;       // shared: tag | payload per slot
  const uint32_t *big;       // shared: 8 limbs per big value
  const field_consts *F;
  __device__ __forceinline__ uint64_t get(uint32_t s) const { return val[s]; }      // tag | payload
  __device__ __forceinline__ bool small(uint32_t s, i128 &v) const {
    const uint64_t x = val[s];
    if (x & STG_TAG_BIG) return false;
    v = (x & STG_TAG_NEG) ? -(i128)(x & STG_PAYLOAD) : (i128)x;
    return true;
  }
  __device__ __forceinline__ fr_t field(uint32_t s) const {
    const uint64_t x = val[s];
    fr_t r;
    if (x & STG_TAG_BIG) {
      const uint32_t *b = big + 8 * (uint32_t)(x & 0xFFFFFFFFu);
#pragma unroll
      for (int j = 0; j < 8; j++) r.l[j] = b[j];
      return r;
    }
    r = fr_from_u64(x & STG_PAYLOAD);
    return (x & STG_TAG_NEG) ? fr_neg(r, F->p) : r;
  }
};

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

__global__ void __launch_bounds__(STG_THREADS, 1)
k_r1cs_check_staged(const uint8_t *__restrict__ wit, uint64_t n, uint32_t ws, const r1cs_tables_dev T, const field_consts *__restrict__ F,
                    uint8_t *__restrict__ status, uint32_t *__restrict__ first_bad) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  uint64_t *val = reinterpret_cast<uint64_t *>(s_raw);
  uint32_t *big = reinterpret_cast<uint32_t *>(s_raw + (size_t)((ws + 1) & ~1u) * 8);
  __shared__ uint32_t s_nbig, s_bad, s_noncanon;
  __shared__ r1cs_class_dev s_cls[STG_MAX_CLASSES];
  __shared__ uint32_t s_nblk[STG_MAX_CLASSES];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  for (uint32_t k = tid; k < T.n_classes; k += STG_THREADS) { s_cls[k] = T.cls[k]; s_nblk[k] = T.cls_blocks[k]; }
  fr_t p;
#pragma unroll
  for (int j = 0; j < 8; j++) p.l[j] = F->p.l[j];
  for (uint64_t i = blockIdx.x; i < n; i += gridDim.x) {
    __syncthreads();                                        // the previous instance's rows are done with val / big
    if (tid == 0) { s_nbig = 0; s_bad = B3W_NO_ROW; s_noncanon = 0; }
    __syncthreads();
    // ---- phase 1: stream the witness once (4 slots per thread in flight), keep 8 bytes per slot ----
    const uint4 *w = reinterpret_cast<const uint4 *>(wit + i * (uint64_t)ws * 32);
    for (uint32_t s0 = tid; s0 < ws; s0 += 4 * STG_THREADS) {
      uint4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t s = min(s0 + u * STG_THREADS, ws - 1);
        a[u] = __ldcs(w + 2 * s);
        b[u] = __ldcs(w + 2 * s + 1);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t s = s0 + u * STG_THREADS;
        if (s >= ws) break;
        uint64_t v;
        if ((a[u].z | a[u].w | b[u].x | b[u].y | b[u].z | b[u].w) == 0 && (a[u].y >> 30) == 0) {
          v = ((uint64_t)a[u].y << 32) | a[u].x;
        } else {
          fr_t x, d;
          x.l[0] = a[u].x; x.l[1] = a[u].y; x.l[2] = a[u].z; x.l[3] = a[u].w;
          x.l[4] = b[u].x; x.l[5] = b[u].y; x.l[6] = b[u].z; x.l[7] = b[u].w;
          const uint32_t borrow = fr_raw_sub(d, p, x);       // p - x: a small negative integer stored canonically?
          if (borrow || fr_is_zero(d)) atomicOr(&s_noncanon, 1u);      // x >= p: not a canonical field element
          if (!borrow && (d.l[2] | d.l[3] | d.l[4] | d.l[5] | d.l[6] | d.l[7]) == 0 && (d.l[1] >> 30) == 0) {
            v = STG_TAG_NEG | ((uint64_t)d.l[1] << 32) | d.l[0];
          } else {
            const uint32_t k = atomicAdd(&s_nbig, 1u);
            v = STG_TAG_BIG | k;
            if (k < STG_MAX_BIG) {
#pragma unroll
              for (int j = 0; j < 8; j++) big[8 * k + j] = x.l[j];
            }
          }
        }
        val[s] = v;
      }
    }
    __syncthreads();
    uint32_t bad = B3W_NO_ROW;
    if (s_nbig > STG_MAX_BIG) {
      bad = 0;                                               // more field-valued slots than any witness of these circuits holds
    } else if (!s_noncanon && !STG_EXP_SKIP_P2) {
      // ---- phase 2: every row from shared memory, one block of <= 32 rows per warp step ----
      const StagedSrc src{val, big, F};
      // the fast paths need wire 0 to be the constant 1 (BOOLROW classes rely on it); classes with wide coefficients keep
      // the 128-bit evaluator
      const bool fast_ok = val[0] == 1ull;
      for (uint32_t ci = 0; ci < T.n_classes; ci++) {
        const r1cs_class_dev c = s_cls[ci];
        const uint32_t nb = s_nblk[ci], hw = 2u + 2u * (c.nA + c.nB + c.nC);
        const bool fast = fast_ok && (c.flags & R1CS_FLAG_FAST64);
        if (c.flags & R1CS_FLAG_MATRIX) {
          for (uint32_t r = tid; r < ((c.count + 31u) & ~31u); r += STG_THREADS)
            bad = min(bad, fast ? staged_matrix_row<true>(src, c, T, r) : staged_matrix_row<false>(src, c, T, r));
        } else {
          for (uint32_t b = warp; b < nb; b += STG_THREADS / 32) {
            const uint32_t *hdr = T.terms + c.term_off + (size_t)b * hw;
            bad = min(bad, fast ? staged_block<true>(src, c, T, hdr, lane) : staged_block<false>(src, c, T, hdr, lane));
          }
        }
      }
    }
    if (bad != B3W_NO_ROW) atomicMin(&s_bad, bad);
    __syncthreads();
    if (tid == 0) {
      const uint32_t verdict = s_noncanon ? B3W_NOT_CANONICAL : s_bad;
      if (status) status[i] = verdict == B3W_NO_ROW ? 0 : B3W_R1CS_VIOLATION;
      if (first_bad) first_bad[i] = verdict;
    }
  }
}
