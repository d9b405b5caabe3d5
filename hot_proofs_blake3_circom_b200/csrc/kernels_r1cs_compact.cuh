// kernels_r1cs_compact.cuh -- the stand-alone R1CS check with a COMPACT shared-memory copy of the witness.
// Included by blake3wit.cu only, after kernels_r1cs_staged.cuh (it reuses that file's row evaluators, row blocks and
// exactness rules; only the value source differs).
//
// k_r1cs_check_staged keeps 8 bytes per slot (193 KB for 24 093 slots): one CTA per SM, which then has to hide the L2
// latency of its ~1 200 row-block headers per instance by itself -- 1.2 M witnesses/s, a sixth of what streaming the
// witnesses from HBM allows.  97 % of the slots of these witnesses are bits, so this kernel keeps per slot ONE bit "the
// slot holds 0 or 1" and ONE bit "its value", plus the 8-byte tagged value of every other slot in a side table that is
// found by rank (population count of the not-a-bit map before the slot): ~22 KB per instance instead of 193 KB, i.e.
// several CTAs of 256 threads per SM instead of 1 of 1024, and the latency of one CTA's table loads is covered by the others.
// Exactness is unchanged: the same 64-bit / 128-bit / Fr row evaluators run on values read through CompactSrc::get();
// a genuine field element is re-read from the witness in HBM when the Fr path needs it; a witness with more non-bit
// slots than the side table holds (not one of these circuits' witnesses, but the checker must still answer) is
// evaluated with every non-bit value converted from HBM on the fly.
// Measured on B200 (profiles/r01i_r1cs_check.jsonl, 2^15 witnesses): streaming + classification alone 6.3 TB/s; with the
// rows 2.33 M witnesses/s for blake3_compression (24 544 rows; the 8-byte staged copy 1.65, the warp-per-instance evaluator
// 1.24) and 1.18 M/s for nova O1.  The row arithmetic is what is left (issue-bound): booleanity rows and XOR rows are
// single comparisons (BOOLROW / XORROW), the bit-recomposition rows still go term by term.
#pragma once

#ifndef CPT_EXP
#define CPT_EXP 0                 /* experiment builds only: 1 = no row evaluation, 2 = pass A only */
#endif
#define CPT_THREADS 256
#define CPT_SIDE_MAX 1536u        /* non-bit slots per witness in the side table (compression 713, nova O1 ~1 250) */

// 32-byte slot -> tagged 8-byte value (the staged checker's encoding); BIG = "genuine field element", payload = slot index
__device__ __forceinline__ uint64_t cpt_classify(const uint4 a, const uint4 b, uint32_t slot, const fr_t &p, bool &noncanon) {
  if ((a.z | a.w | b.x | b.y | b.z | b.w) == 0 && (a.y >> 30) == 0) {
    return ((uint64_t)a.y << 32) | a.x;
  }
  fr_t x, d;
  x.l[0] = a.x; x.l[1] = a.y; x.l[2] = a.z; x.l[3] = a.w;
  x.l[4] = b.x; x.l[5] = b.y; x.l[6] = b.z; x.l[7] = b.w;
  const uint32_t borrow = fr_raw_sub(d, p, x);               // p - x: a small negative integer stored canonically?
  noncanon = noncanon || borrow || fr_is_zero(d);            // x >= p: not a canonical field element
  if (!borrow && (d.l[2] | d.l[3] | d.l[4] | d.l[5] | d.l[6] | d.l[7]) == 0 && (d.l[1] >> 30) == 0) {
    return STG_TAG_NEG | ((uint64_t)d.l[1] << 32) | d.l[0];
  }
  return STG_TAG_BIG | slot;
}

struct CompactSrc {
  // several CTAs per SM walk the same tables: let them live in L1 (about 170 KB are left next to 4 compact copies)
  static __device__ __forceinline__ uint32_t ld_table(const uint32_t *p) { return __ldg(p); }
  const uint32_t *isbit, *bitval, *rank;     // shared: one bit per slot (x2), non-bit slots before each 32-slot word
  const uint64_t *side;                      // shared: tagged values of the non-bit slots, in slot order
  const uint4 *wit;                          // this instance's witness in HBM (2 x uint4 per slot)
  const field_consts *F;
  bool side_ok;                              // false: more non-bit slots than the side table holds
  __device__ __forceinline__ uint64_t get(uint32_t s) const {
    const uint32_t w = s >> 5, b = s & 31u, m = isbit[w];
    if ((m >> b) & 1u) return (bitval[w] >> b) & 1u;
    if (side_ok) return side[rank[w] + __popc(~m & ((1u << b) - 1u))];
    bool nc = false;                         // (a non-canonical slot was already reported by the classification pass)
    return cpt_classify(__ldg(wit + 2 * s), __ldg(wit + 2 * s + 1), s, F->p, nc);
  }
  __device__ __forceinline__ bool small(uint32_t s, i128 &v) const {
    const uint64_t x = get(s);
    if (x & STG_TAG_BIG) return false;
    v = (x & STG_TAG_NEG) ? -(i128)(x & STG_PAYLOAD) : (i128)x;
    return true;
  }
  __device__ __forceinline__ fr_t field(uint32_t s) const {
    const uint64_t x = get(s);
    fr_t r;
    if (x & STG_TAG_BIG) {
      const uint4 a = __ldg(wit + 2 * s), b = __ldg(wit + 2 * s + 1);
      r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
      return r;
    }
    r = fr_from_u64(x & STG_PAYLOAD);
    return (x & STG_TAG_NEG) ? fr_neg(r, F->p) : r;
  }
};

#define CPT_CTAS_PER_SM 4          /* 32 warps per SM at 64 registers (the row evaluators use 128-bit and Fr arithmetic) */
__global__ void __launch_bounds__(CPT_THREADS, CPT_CTAS_PER_SM)
k_r1cs_check_compact(const uint8_t *__restrict__ wit, uint64_t n, uint32_t ws, const r1cs_tables_dev T, const field_consts *__restrict__ F,
                     uint8_t *__restrict__ status, uint32_t *__restrict__ first_bad) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  const uint32_t words = (ws + 31u) >> 5;
  uint32_t *isbit = reinterpret_cast<uint32_t *>(s_raw), *bitval = isbit + words, *rank = bitval + words;      // rank: words + 1
  uint64_t *side = reinterpret_cast<uint64_t *>(s_raw + (size_t)((3 * words + 1 + 1) & ~1u) * 4);
  __shared__ uint32_t s_bad, s_flags, s_warp_tot[CPT_THREADS / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  fr_t p;
#pragma unroll
  for (int j = 0; j < 8; j++) p.l[j] = F->p.l[j];
  for (uint64_t i = blockIdx.x; i < n; i += gridDim.x) {
    __syncthreads();                                          // the previous instance's rows are done with the copy
    if (tid == 0) { s_bad = B3W_NO_ROW; s_flags = 0; }
    const uint4 *w = reinterpret_cast<const uint4 *>(wit + i * (uint64_t)ws * 32);
    // ---- pass A: stream the witness once; lane = slot inside a 32-slot word; two words per warp step in flight ----
    for (uint32_t wd = warp; wd < words; wd += 2 * (CPT_THREADS / 32)) {
      uint4 a[2], b[2];
      uint32_t s[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        s[u] = (wd + u * (CPT_THREADS / 32)) * 32 + lane;
        const uint32_t sc = min(s[u], ws - 1);
        a[u] = __ldcs(w + 2 * sc);
        b[u] = __ldcs(w + 2 * sc + 1);
      }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const uint32_t wu = wd + u * (CPT_THREADS / 32);
        if (wu >= words) break;                               // warp-uniform
        const bool in = s[u] < ws;
        const bool bit = in && (a[u].y | a[u].z | a[u].w | b[u].x | b[u].y | b[u].z | b[u].w) == 0 && a[u].x < 2u;
        const uint32_t mb = __ballot_sync(0xffffffffu, bit || !in);      // slots past the end count as bits (value 0)
        const uint32_t mv = __ballot_sync(0xffffffffu, bit && a[u].x == 1u);
        if (lane == 0) { isbit[wu] = mb; bitval[wu] = mv; rank[wu] = __popc(~mb); }
      }
    }
    __syncthreads();
    // ---- exclusive prefix sum of the per-word non-bit counts (rank[words] = total) ----
    {
      const uint32_t per = (words + CPT_THREADS - 1) / CPT_THREADS, lo = tid * per, hi = min(lo + per, words);
      uint32_t sum = 0;
      for (uint32_t k = lo; k < hi; k++) sum += rank[k];
      uint32_t inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += y;
      }
      if (lane == 31) s_warp_tot[warp] = inc;
      __syncthreads();
      uint32_t base = 0;
      for (uint32_t k = 0; k < warp; k++) base += s_warp_tot[k];
      uint32_t run = base + inc - sum;
      for (uint32_t k = lo; k < hi; k++) { const uint32_t c = rank[k]; rank[k] = run; run += c; }
      if (tid == CPT_THREADS - 1) rank[words] = run;
    }
    __syncthreads();
    const uint32_t n_side = rank[words];
    const bool side_ok = n_side <= CPT_SIDE_MAX;
    // ---- pass B: the non-bit slots again (3 % of the witness, L2 hits): classify, fill the side table ----
    if (CPT_EXP < 2) {
      bool noncanon = false;
      for (uint32_t wd = warp; wd < words; wd += CPT_THREADS / 32) {
        const uint32_t m = ~isbit[wd];
        if (!((m >> lane) & 1u)) continue;
        const uint32_t s = wd * 32 + lane;
        const uint64_t v = cpt_classify(__ldg(w + 2 * s), __ldg(w + 2 * s + 1), s, p, noncanon);
        if (side_ok) side[rank[wd] + __popc(m & ((1u << lane) - 1u))] = v;
      }
      if (noncanon) atomicOr(&s_flags, 1u);
    }
    __syncthreads();
    uint32_t bad = B3W_NO_ROW;
    if (!(s_flags & 1u) && CPT_EXP == 0) {
      // ---- every row from the compact copy, one block of <= 32 rows per warp step ----
      const CompactSrc src{isbit, bitval, rank, side, w, F, side_ok};
      const bool fast_ok = src.get(0) == 1ull;
      for (uint32_t ci = 0; ci < T.n_classes; ci++) {
        const r1cs_class_dev c = T.cls[ci];
        const uint32_t nb = T.cls_blocks[ci], hw = 2u + 2u * (c.nA + c.nB + c.nC);
        const bool fast = fast_ok && (c.flags & R1CS_FLAG_FAST64);
        if (c.flags & R1CS_FLAG_MATRIX) {
          for (uint32_t r = tid; r < ((c.count + 31u) & ~31u); r += CPT_THREADS)
            bad = min(bad, fast ? staged_matrix_row<true>(src, c, T, r) : staged_matrix_row<false>(src, c, T, r));
        } else {
          for (uint32_t b = warp; b < nb; b += CPT_THREADS / 32) {
            const uint32_t *hdr = T.terms + c.term_off + (size_t)b * hw;
            // this warp's next header on its way into L1 while the current block is evaluated
            if (b + CPT_THREADS / 32 < nb && lane * 32u < hw)
              asm volatile("prefetch.global.L1 [%0];" ::"l"(hdr + (size_t)(CPT_THREADS / 32) * hw + lane * 32u));
            bad = min(bad, fast ? staged_block<true>(src, c, T, hdr, lane) : staged_block<false>(src, c, T, hdr, lane));
          }
        }
      }
    }
    if (bad != B3W_NO_ROW) atomicMin(&s_bad, bad);
    __syncthreads();
    if (tid == 0) {
      const uint32_t verdict = (s_flags & 1u) ? B3W_NOT_CANONICAL : s_bad;
      if (status) status[i] = verdict == B3W_NO_ROW ? 0 : B3W_R1CS_VIOLATION;
      if (first_bad) first_bad[i] = verdict;
    }
  }
}
