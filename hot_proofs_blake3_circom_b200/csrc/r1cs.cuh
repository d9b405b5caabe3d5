// r1cs.cuh -- on-device R1CS satisfiability check:  (A.z) * (B.z) - C.z == 0 for every row, over Fr.
//
// The reference never checks constraints in the witness path itself; its tests do, through circom_tester
// (`expectPass`, test/blake3_hash.test.ts:36,57) and the Rust side through bellpepper
// (rust_fold/src/utils.rs:78-85 enforces every row).  The rows here are re-derived from the circom templates
// by tools/gen_r1cs.py (r1cs_tables.h), grouped into shape classes: the 32 lanes of a warp evaluate 32 rows of
// identical shape, term columns are term-major so that lane loads coalesce.
//
// This file holds the FUSED check (value source TraceSrc); the stand-alone check of witnesses resident in HBM lives in
// kernels_r1cs_fast.cuh (a program compiled from the rows, evaluated on a compact shared-memory copy of the witness) with
// r1cs_rows.cuh for the rows that program does not take.
//   TraceSrc  -- the fused check: a term is a slot DESCRIPTOR and its value is taken from the shared-memory trace
//                that the expansion is about to read (nothing is re-read from HBM).  In trace space the rows that are
//                identities for ANY trace content (booleanity of a bit extracted by shift-and-mask; w == sum 2^i
//                bit_i(w)) need no evaluation, and the 32 rows 2*x_i*y_i = x_i + y_i - o_i over the bits of one word
//                triple are evaluated bit-sliced as ONE word comparison (class flag XORW), and a recomposition row
//                whose only content in trace space is "these bits of that word are 0" (e.g. Bits34: the carry word
//                of a 34-bit sum is < 4) is evaluated as ONE mask test (class flag ZMASK): the *_FUSED row sets;
// Arithmetic: every row of these circuits except IsZero's `in*inv = 1 - out` is an identity between integers far
// below p (bits, u32 words, <= 66-bit sums, small negatives), so it is evaluated exactly in signed 64-bit (NARROW
// classes) or signed 128-bit (WIDE) integers; the 67 IsZero rows per nova witness go through Montgomery
// multiplication in Fr (FIELD classes) in the stand-alone check and are folded to integer tests (ISZ) in the fused one.  A slot that holds anything but a small (|x| < 2^63) integer inside an
// integer row cannot satisfy it and is reported as a violation.
#pragma once
#include <stdint.h>
#include "fr.cuh"
#include "trace_layout.h"

#define R1CS_FLAG_FIELD 1u
#define R1CS_FLAG_WIDE 2u
#define R1CS_FLAG_XORW 4u     /* fused set only: 32 XOR rows of one word triple, X ^ rotr(Y, dy) == rotr(O, do) */
#define R1CS_FLAG_ROWCOEF 8u  /* coefficients are stored per row (term-major) instead of once per class */
#define R1CS_FLAG_ZMASK 16u   /* fused set only: a bit-recomposition row, folded to  trace[word] & mask == 0 */
#define R1CS_FLAG_ISZ 32u     /* fused set only: IsZero's in * inv = C.z with inv = INV(in), folded to  C.z == (in != 0) */
#define B3W_NO_ROW 0xFFFFFFFFu

struct r1cs_class_dev {
  uint16_t nA, nB, nC, flags;
  uint32_t count, coef_off, term_off, row_off;     // term_off: start of this class's term matrix; row_off: first row id
};

struct r1cs_tables_dev {
  const r1cs_class_dev *cls;
  const int64_t *coef_lo;      // low 64 bits of each coefficient (two's complement)
  const int64_t *coef_hi;      // high 64 bits
  const uint32_t *terms;       // descriptors (TraceSrc) or slot indices (StagedSrc), [class][term][row]
  uint32_t n_classes;
  const fr_t *coef_fr;         // loaded sets only: full field-element coefficients of the BIGCOEF classes (else NULL)
  const uint32_t *row_ids;     // loaded sets only: constraint index in the .r1cs file of each class row (else NULL)
  const uint32_t *cls_blocks;  // slot-space sets: `terms` holds row BLOCKS (kernels_r1cs_staged.cuh), this many per class
};

typedef __int128 i128;

struct TraceSrc {
  const uint32_t *trace;
  const field_consts *F;
  __device__ __forceinline__ bool xorw(const r1cs_class_dev &c, const r1cs_tables_dev &T, uint32_t r) const;
  __device__ __forceinline__ bool zmask(const r1cs_class_dev &c, const r1cs_tables_dev &T, uint32_t r) const {
    const uint32_t *tm = T.terms + c.term_off;
    return (trace[tm[r] & 0xFFFFu] & tm[c.count + r]) == 0u;
  }
  __device__ __forceinline__ bool small(uint32_t d, i128 &v) const {
    const uint32_t t = d & 0xFFFFu, k = (d >> 16) & 31u, kind = d >> 24;
    const uint32_t w = trace[t];
    if (kind == DK_BIT) v = (w >> k) & 1u;
    else if (kind == DK_W32) v = w;
    else if (kind == DK_W64) v = (i128)(((uint64_t)trace[t + 1] << 32) | w);
    else if (kind == DK_S64) v = (i128)(int64_t)(((uint64_t)trace[t + 1] << 32) | w);
    else return false;
    return true;
  }
  __device__ __forceinline__ fr_t field(uint32_t d) const {
    const uint32_t t = d & 0xFFFFu, kind = d >> 24;
    const int64_t x = (int64_t)(((uint64_t)trace[t + 1] << 32) | trace[t]);
    if (kind == DK_INV) return fr_inv_s64(x, *F);
    i128 v;
    small(d, v);
    return fr_from_s64((int64_t)v, F->p);      // FIELD rows only hold S64 / INV / bit terms
  }
};

// XORW row: for bits, 2xy = x + y - o  <=>  o = x ^ y; all 32 rows of the word triple at once.
__device__ __forceinline__ bool r1cs_xorw_row(const uint32_t *trace, const r1cs_class_dev &c, const r1cs_tables_dev &T, uint32_t r) {
  const uint32_t *tm = T.terms + c.term_off;
  const uint32_t dx = tm[r], dy = tm[c.count + r], dz = tm[2 * c.count + r];
  const uint32_t X = trace[dx & 0xFFFFu], Y = trace[dy & 0xFFFFu], O = trace[dz & 0xFFFFu];
  return (X ^ __funnelshift_r(Y, Y, (dy >> 16) & 31u)) == __funnelshift_r(O, O, (dz >> 16) & 31u);
}

__device__ __forceinline__ bool TraceSrc::xorw(const r1cs_class_dev &c, const r1cs_tables_dev &T, uint32_t r) const {
  return r1cs_xorw_row(trace, c, T, r);
}

// FIELD row (IsZero):  a * b == C.z  with single unit-coefficient A and B terms.
template <class Src>
__device__ __noinline__ bool r1cs_field_row(const Src &src, const r1cs_class_dev &c, const r1cs_tables_dev &T, uint32_t r) {
  const uint32_t *tm = T.terms + c.term_off;
  const fr_t a = src.field(tm[r]), b = src.field(tm[c.count + r]);
  const field_consts &F = *src.F;
  const fr_t ab = fr_montmul(fr_montmul(a, b, F.p, F.n0), F.r2, F.p, F.n0);
  i128 lc = 0;
  for (uint32_t t = 0; t < c.nC; t++) {
    i128 v;
    if (!src.small(tm[(2 + t) * c.count + r], v)) return false;
    lc += (i128)T.coef_lo[(c.flags & R1CS_FLAG_ROWCOEF) ? c.coef_off + (2 + t) * c.count + r : c.coef_off + 2 + t] * v;
  }
  const fr_t want = fr_from_s64((int64_t)lc, F.p);
  bool eq = true;
#pragma unroll
  for (int j = 0; j < 8; j++) eq = eq && (ab.l[j] == want.l[j]);
  return eq;
}

// ISZ row (fused set): A = the IsZero input (a small integer), no B, C.z must equal (in != 0).
template <class Src>
__device__ __forceinline__ bool r1cs_isz_row(const Src &src, const r1cs_class_dev &c, const r1cs_tables_dev &T, uint32_t r) {
  const uint32_t *tm = T.terms + c.term_off;
  i128 x, lc = 0;
  bool ok = src.small(tm[r], x);
  for (uint32_t t = 1; t <= c.nC; t++) {
    i128 v;
    ok = src.small(tm[t * c.count + r], v) && ok;
    lc += (i128)T.coef_lo[(c.flags & R1CS_FLAG_ROWCOEF) ? c.coef_off + t * c.count + r : c.coef_off + t] * v;
  }
  return ok && lc == (x != 0 ? 1 : 0);
}

// Integer rows, U rows per lane at a time (rows r0, r0 + 32, ...): the U accumulator chains are independent, so the
// descriptor / coefficient / value loads of U rows are in flight together -- the check is latency-bound, not issue-bound
// (profiles/: one row per lane left the fused nova kernel at 4.95 TB/s with 35 % of the issue slots used).  Rows past
// the end of the class are clamped to its last row and their verdict ignored, which keeps every load unconditional.
// Returns the smallest violated row id of this lane's group, or B3W_NO_ROW.
template <class Src, typename acc_t, int U>
__device__ __forceinline__ uint32_t r1cs_int_rows(const Src &src, const r1cs_class_dev &c, const r1cs_tables_dev &T, uint32_t r0) {
  const uint32_t *tm = T.terms + c.term_off;
  const uint32_t last = c.count - 1;
  uint32_t r[U];
  acc_t L[U][3];
  bool ok[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    r[u] = min(r0 + 32u * u, last);
    L[u][0] = L[u][1] = L[u][2] = 0;
    ok[u] = true;
  }
  const uint32_t n[3] = {c.nA, c.nB, c.nC};
  const bool rowcoef = (c.flags & R1CS_FLAG_ROWCOEF) != 0;
  uint32_t t = 0;
#pragma unroll
  for (int part = 0; part < 3; part++) {
    for (uint32_t j = 0; j < n[part]; j++, t++) {
      const uint32_t *col = tm + t * c.count;
      uint32_t d[U];
#pragma unroll
      for (int u = 0; u < U; u++) d[u] = col[r[u]];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t ci = rowcoef ? c.coef_off + t * c.count + r[u] : c.coef_off + t;
        const acc_t co = sizeof(acc_t) == 16 ? (acc_t)(((i128)T.coef_hi[ci] << 64) | (i128)(uint64_t)T.coef_lo[ci]) : (acc_t)T.coef_lo[ci];
        i128 v;
        ok[u] = src.small(d[u], v) && ok[u];
        L[u][part] += co * (acc_t)v;
      }
    }
  }
  uint32_t bad = B3W_NO_ROW;
#pragma unroll
  for (int u = U - 1; u >= 0; u--) {
    const bool holds = ok[u] && (c.nA == 0 ? L[u][2] == 0 : L[u][0] * L[u][1] == L[u][2]);
    if (!holds && r0 + 32u * u <= last) bad = c.row_off + r[u];
  }
  return bad;
}

// Evaluate every row of one instance with one warp.  Returns the smallest violated row id over the warp's lanes
// (B3W_NO_ROW if the instance satisfies the system); all lanes return the same value.
template <class Src>
__device__ __forceinline__ uint32_t r1cs_check_instance(const Src &src, const r1cs_tables_dev &T, int lane) {
  uint32_t bad = B3W_NO_ROW;
  for (uint32_t ci = 0; ci < T.n_classes; ci++) {
    const r1cs_class_dev c = T.cls[ci];
    if (c.flags & R1CS_FLAG_ZMASK) {
      for (uint32_t r = lane; r < c.count; r += 32)
        if (!src.zmask(c, T, r)) bad = min(bad, c.row_off + r);
    } else if (c.flags & R1CS_FLAG_ISZ) {
      for (uint32_t r = lane; r < c.count; r += 32)
        if (!r1cs_isz_row(src, c, T, r)) bad = min(bad, c.row_off + r);
    } else if (c.flags & (R1CS_FLAG_XORW | R1CS_FLAG_FIELD)) {
      for (uint32_t r = lane; r < c.count; r += 32) {
        const bool ok = (c.flags & R1CS_FLAG_XORW) ? src.xorw(c, T, r) : r1cs_field_row(src, c, T, r);
        if (!ok) bad = min(bad, c.row_off + r);
      }
    } else if (c.flags & R1CS_FLAG_WIDE) {
      for (uint32_t r = lane; r < c.count; r += 64) bad = min(bad, r1cs_int_rows<Src, i128, 2>(src, c, T, r));
    } else {
      for (uint32_t r = lane; r < c.count; r += 128) bad = min(bad, r1cs_int_rows<Src, int64_t, 4>(src, c, T, r));
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) bad = min(bad, __shfl_xor_sync(0xffffffffu, bad, o));
  return bad;
}
