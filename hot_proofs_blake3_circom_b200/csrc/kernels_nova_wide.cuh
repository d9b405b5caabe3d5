// kernels_nova_wide.cuh -- the nova step circuits on their FULL input domain: any 32 field elements, as the reference's
// wasm takes them (`normalize`, witness_calculator.js:319-323).  Included by blake3wit.cu only, after kernels_witness.cuh.
//
// As built, Blake3Nova constrains very little of its input (probed with the reference wasm):
//   * CheckDepth only sees leaf_depth - depth, which must lie in [1, 256] (two Num2Bits(9) inside LessThan(8) /
//     GreaterEqThan(8) and `exceed_depth.out === 0`, circuits/blake3_nova.circom:31-44 as built); depth itself may be ANY
//     field element, and so may total_depth, n_blocks and block_count (they only meet IsEqual);
//   * chunk_idx_low + 2^32 * chunk_idx_high must be < 2^65 (Num2Bits(65), :56-59) -- the two parts individually need not be
//     small (e.g. low = p - 2^32, high = 1);
//   * h, chunk_idx_low/high (leaf steps), b and the message words then meet the range constraints of the embedded
//     Blake3Compression exactly as in wide_domain.h: u32 for everything that passes a ToBits(32), a 34-bit sum window
//     for the message words; on parent steps h and m[8..15] do not reach the compression at all.
// None of this occurs in the reference's drivers (every input is a u32 there: the hot kernels' domain).  This kernel is
// the general form, used by b3w_witness_batch_fr when a batch holds an input outside u32: one warp per instance, the 32
// inputs as field elements in shared memory, the nova-level logic in Fr (lane 0; it is a few dozen additions and
// comparisons), then the ordinary u32 trace of the embedded compression with the wide-message correction, the ordinary
// expansion, and finally an OVERRIDE pass that rewrites every slot whose value is a function of a field-valued input
// (the slot list is derived from the descriptor table at context set-up: kinds W32 / W64 / S64 / INV on the trace words
// below).  Throughput is irrelevant here (it still runs at hundreds of thousands of witnesses per second).
#pragma once
#include "nova_wide_logic.h"

#define NW_FIN_WORDS 256u          /* 32 inputs x 8 limbs */
#define NW_EXTRA_WORDS (NW_FIN_WORDS + 16u /* ext */ + 64u /* selectors */ + 16u /* scalars */)
#define NW_STRIDE (NOVA_TRACE_STRIDE + NW_EXTRA_WORDS)
#define NW_WARPS 4
#define DK_WIDE_BIT64 7u           /* pseudo descriptor kind of the override list: Num2Bits(65).out[64] (O1 build only) */

struct nova_wide_args {
  const uint8_t *in_fr;            // n x 32 x 32 bytes, canonical (already reduced mod p)
  const uint2 *wslots;             // override list {slot, descriptor}, grouped by slot % 32
  const uint32_t *lane_off;        // 33 offsets into wslots: lane l owns [lane_off[l], lane_off[l + 1])
  const field_consts *F;
};

// per-warp scratch behind the trace
struct nw_view {
  uint32_t *trace;
  uint32_t *fin;                   // [32][8]
  uint32_t *ext;                   // [16] sign-extended
  uint8_t *sel;                    // [4][16]: which input (or NW_SEL_ZERO) tmp_down / m_is_parent / tmp_is_par / out_m select
  uint32_t *sc;                    // scalars: 0 not_parent, 1 decr, 2..4 S limbs 0..2
  __device__ __forceinline__ fr_t in(uint32_t k) const {
    fr_t v;
#pragma unroll
    for (int j = 0; j < 8; j++) v.l[j] = fin[8 * k + j];
    return v;
  }
};

// The value of an overridden slot: descriptor kind + trace word -> the field element the circuit holds there.
__device__ __noinline__ fr_t nova_wide_value(const nw_view &w, uint32_t dsc, const field_consts &F) {
  const uint32_t t = dsc & 0xFFFFu, kind = dsc >> 24;
  const fr_t &p = F.p;
  fr_t v = fr_zero();
  if (kind == DK_WIDE_BIT64) return fr_from_u64(w.sc[4] & 1u);
  if (t >= NV_IN && t < NV_IN + 32) {
    if (kind == DK_W64 && t == NV_IN + 10) { v.l[0] = w.sc[2]; v.l[1] = w.sc[3]; v.l[2] = w.sc[4]; }     // chunk_idx = low + 2^32 high
    else v = w.in(t - NV_IN);
  } else if (t == NV_LDM1) v = nw_add_small(w.in(12), -1, p);
  else if (t == NV_DP1) v = nw_add_small(w.in(14), 1, p);
  else if (t == NV_DEPTH_OUT) v = nw_add_small(w.in(14), -(int)w.sc[1], p);                             // depth - decr_depth (:262)
  else if (t == NV_NEG_DEPTH) v = fr_neg(w.in(14), p);
  else if (t == NV_NEG_BC) v = fr_neg(w.in(1), p);
  else if (t == NV_NBM1) v = nw_add_small(w.in(0), -1, p);
  else if (t == NV_BC_DIFF) v = nw_sub(nw_add_small(w.in(0), -1, p), w.in(1), p);
  else if (t == NV_BC_OUT) v = nw_add_small(w.in(1), (int)w.sc[0], p);                                  // block_count + (1 - is_parent) (:251)
  else if (t >= NV_EQ_IN1 && t < NV_EQ_IN1 + 128) v = nw_add_small(w.in(13), -(int)((t - NV_EQ_IN1) / 2) - 2, p);
  else if (t >= NV_EQ_D && t < NV_EQ_D + 128) v = nw_sub(nw_add_small(w.in(13), -(int)((t - NV_EQ_D) / 2) - 2, p), w.in(14), p);
  else {
    uint32_t s = NW_SEL_ZERO;
    if (t >= NV_TMP_DOWN && t < NV_TMP_DOWN + 16) s = w.sel[t - NV_TMP_DOWN];
    else if (t >= NV_M_IS_PAR && t < NV_M_IS_PAR + 16) s = w.sel[16 + t - NV_M_IS_PAR];
    else if (t >= NV_TMP_IS_PAR && t < NV_TMP_IS_PAR + 16) s = w.sel[32 + t - NV_TMP_IS_PAR];
    else if (t >= TR_IN + 8 && t < TR_IN + 24) s = w.sel[48 + t - (TR_IN + 8)];
    if (s != NW_SEL_ZERO) v = w.in(s);
  }
  if (kind == DK_INV && !fr_is_zero(v)) {
    // IsZero.inv.  Most of these differences stay small integers even when one input is a genuine field element (e.g. only
    // n_blocks is): +-k with k < 256 comes from the table, like on the hot path; only the rest pays a Fermat inversion
    const fr_t neg = fr_neg(v, p);
    if (nw_fits(v, 8)) v = F.inv_small[v.l[0]];
    else if (nw_fits(neg, 8)) v = fr_neg(F.inv_small[neg.l[0]], p);
    else v = fr_inv(v, F.p, F.r2, F.n0);
  }
  return v;
}

// Nova-level logic on field elements.  Fills the u32 trace exactly where nova_trace() does (small values, bits, the low
// words of what the override pass rewrites), the selectors and ext.  Returns false when a constraint fails.
__device__ __forceinline__ bool nova_trace_wide(const nw_view &w, const field_consts &F, int lane) {
  uint32_t *trace = w.trace;
  const fr_t &p = F.p;
  // ---- lane 0: the scalar part (CheckDepth, GetFlag, DownLeftPath), shared with the host's assert replay ----
  uint32_t ok = 1u;
  if (lane == 0) {
    const nova_wide_scalars s = nova_wide_scalar_logic(nw_in_words{w.fin}, p);
    trace[NV_V1] = s.v1; trace[NV_V2] = s.v2;
    trace[NV_LDM1] = w.fin[8 * 12] - 1u; trace[NV_DP1] = w.fin[8 * 14] + 1u;
    trace[NV_IS_PARENT] = s.is_parent; trace[NV_EXCEED] = 0u; trace[NV_IS_ROOT] = s.is_root;
    trace[NV_NOT_ROOT] = s.not_root; trace[NV_NOT_PARENT] = s.not_parent;
    trace[NV_BC_FIRST] = s.first; trace[NV_BC_LAST] = s.last; trace[NV_IS_LAST] = s.is_last; trace[NV_FIRST_SET] = s.first_set;
    trace[NV_URF_TMP] = s.urf_tmp; trace[NV_URF] = s.urf; trace[NV_DLP] = s.dlp;
    trace[NV_CDD] = s.cdd; trace[NV_DECR] = s.decr; trace[NV_DEPTH_OUT] = nw_add_small(w.in(14), -(int)s.decr, p).l[0];
    trace[NV_EQ_OUT] = (uint32_t)s.eq; trace[NV_EQ_OUT + 1] = (uint32_t)(s.eq >> 32);
    trace[NV_BAD] = (uint32_t)s.bad; trace[NV_BAD + 1] = (uint32_t)(s.bad >> 32);
    trace[NV_BC_OUT] = nw_add_small(w.in(1), (int)s.not_parent, p).l[0]; trace[NV_BC_OUT + 1] = 0u;
    trace[TR_IN + 27] = s.dflags;
    w.sc[0] = s.not_parent; w.sc[1] = s.decr; w.sc[2] = s.S[0]; w.sc[3] = s.S[1]; w.sc[4] = s.S[2];
    w.sc[5] = s.is_parent; w.sc[6] = s.dlp; w.sc[7] = s.fail == NW_FAIL_NONE ? 1u : 0u;
  }
  __syncwarp();
  ok = w.sc[7];
  const uint32_t is_parent = w.sc[5], dlp = w.sc[6], not_parent = w.sc[0];
  // ---- the 32 input words; the bits of chunk_idx are read from words 10 / 11, which therefore hold low + 2^32 high ----
  trace[NV_IN + lane] = lane == 10 ? w.sc[2] : lane == 11 ? w.sc[3] : w.fin[8 * lane];
  // ---- Blake3GetFinal_m (:86-120): every value is one of the inputs or 0 ----
  if (lane < 16) {
    const nova_wide_sel q = nova_wide_select((uint32_t)lane, is_parent, dlp);
    const uint32_t td = q.tmp_down, mp = q.m_is_parent, tp = q.tmp_is_par, om = q.out_m;
    w.sel[lane] = (uint8_t)td; w.sel[16 + lane] = (uint8_t)mp; w.sel[32 + lane] = (uint8_t)tp; w.sel[48 + lane] = (uint8_t)om;
    trace[NV_TMP_DOWN + lane] = td == NW_SEL_ZERO ? 0u : w.fin[8 * td];
    trace[NV_M_IS_PAR + lane] = w.fin[8 * mp];
    trace[NV_TMP_IS_PAR + lane] = tp == NW_SEL_ZERO ? 0u : w.fin[8 * tp];
    // out_m -> the compression's message word: a signed integer ext * 2^32 + lo with ext in [-2, 3], or it asserts
    const fr_t v = w.in(om);
    uint32_t lo = v.l[0];
    int e = 0;
    bool fits = true;
    if (nw_fits(v, 34)) e = (int)v.l[1];
    else {
      fr_t k;
      fr_raw_sub(k, p, v);
      const uint64_t k64 = ((uint64_t)k.l[1] << 32) | k.l[0];
      if (!nw_fits(k, 34) || k64 > (1ull << 33)) fits = false;
      const uint64_t xx = 0ull - k64;
      lo = (uint32_t)xx;
      e = (int)(int32_t)(uint32_t)(xx >> 32);
    }
    trace[TR_IN + 8 + lane] = lo;
    w.ext[lane] = fits ? (uint32_t)e : (uint32_t)B3W_EXT_ASSERT;
  }
  if (lane < 8) {                                                             // h_compression (:229-233): ToBits(32) later on
    const uint32_t IVc[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    uint32_t ivw = IVc[0];
#pragma unroll
    for (int j = 1; j < 8; j++) ivw = lane == j ? IVc[j] : ivw;
    trace[NV_TMPIV + lane] = ivw * is_parent;
    trace[TR_IN + lane] = is_parent ? ivw : w.fin[8 * (2 + lane)];
    if (!is_parent && !nw_fits(w.in(2 + lane), 32)) ok = 0u;
  }
  if (lane == 8 || lane == 9) {                                               // t = chunk_idx_low / high * (1 - is_parent) (:244-245)
    trace[TR_IN + 24 + (lane - 8)] = not_parent ? w.fin[8 * (10 + lane - 8)] : 0u;
    if (not_parent && !nw_fits(w.in(10 + lane - 8), 32)) ok = 0u;
  }
  if (lane == 10) {
    trace[TR_IN + 26] = w.fin[8 * 31];
    if (!nw_fits(w.in(31), 32)) ok = 0u;                                      // b
  }
  // the signed-64 words (NEG_DEPTH, ..., EQ_IN1, EQ_D) are only read through S64 / INV descriptors: all overridden
  return __all_sync(0xffffffffu, ok != 0u);
}

// the 15 outputs z_{i+1}, low 32 bits each (the full values are witness slots 1..15)
__device__ __forceinline__ uint32_t nova_wide_public_output(const nw_view &w, int lane) {
  if (lane == 0) return w.fin[0];
  if (lane == 1) return w.trace[NV_BC_OUT];
  if (lane < 10) return w.trace[TR_OUT + lane - 2];
  if (lane == 10) return w.fin[8 * 13];
  if (lane == 11) return w.trace[NV_DEPTH_OUT];
  if (lane == 12) return w.fin[8 * 10];
  if (lane == 13) return w.fin[8 * 11];
  return w.fin[8 * 12];
}

// list == NULL: instances [0, n) of in_fr.  list != NULL: the instances list[1 .. list[0]] (k_fr_to_rows_nova's wide list; n is
// ignored) -- in_fr, out, status, pub, sums are indexed by the INSTANCE, so the listed witnesses land where the hot kernel
// left their places empty.  sums (may be NULL): per-instance witness checksum, written (not accumulated) here.
__global__ void __launch_bounds__(NW_WARPS * 32)
k_blake3_nova_witness_wide(const nova_wide_args wa, const uint32_t *__restrict__ list, uint64_t n, const uint32_t *__restrict__ desc, uint32_t ws,
                           uint8_t *__restrict__ out, uint8_t *__restrict__ status, uint32_t *__restrict__ pub,
                           unsigned long long *__restrict__ sums) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  nw_view w;
  w.trace = s_dyn + wib * NW_STRIDE;
  w.fin = w.trace + NOVA_TRACE_STRIDE;
  w.ext = w.fin + NW_FIN_WORDS;
  w.sel = reinterpret_cast<uint8_t *>(w.ext + 16);
  w.sc = w.ext + 16 + 16;
  const field_consts &F = *wa.F;
  const lane_sched ls = load_lane_sched(lane);
  const uint64_t warp = (uint64_t)blockIdx.x * NW_WARPS + wib, nwarps = (uint64_t)gridDim.x * NW_WARPS;
  const uint64_t count = list ? (uint64_t)list[0] : n;
  for (uint64_t it = warp; it < count; it += nwarps) {
    const uint64_t i = list ? (uint64_t)list[1 + it] : it;
    __syncwarp();
    for (uint32_t k = lane; k < NOVA_TRACE_STRIDE; k += 32) w.trace[k] = 0u;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(wa.in_fr + i * 1024);
    for (uint32_t k = lane; k < NW_FIN_WORDS; k += 32) w.fin[k] = __ldg(src + k);
    __syncwarp();
    if (lane == 0) w.trace[TR_ONE] = 1u;
    bool ok = nova_trace_wide(w, F, lane);
    __syncwarp();
    if (ok) {
      compression_trace(w.trace, lane, ls);
      __syncwarp();
      ok = !wide_carries_at<true>(w.trace, w.ext, lane);
      __syncwarp();
    }
    if (status && lane == 0) status[i] = ok ? 0 : B3W_CIRCOM_ASSERT;
    if (pub && lane < 15) pub[i * 15 + lane] = ok ? nova_wide_public_output(w, lane) : 0u;
    if (!ok) {                                               // the reference throws "Assert Failed.": no witness exists
      if (sums && lane == 0) sums[i] = 0ull;
      continue;
    }
    uint8_t *dst = out + i * (uint64_t)ws * 32;
    uint64_t acc = expand_slots<true, true>(w.trace, desc, 0, ws, dst, lane, wa.F, nullptr, 0);     // S64 / INV slots: left to the override pass
    __syncwarp();
    // override pass: lane l rewrites the slots with slot % 32 == l -- the lane that wrote them above (expand_slots starts
    // at slot 0), so both stores to an address come from one thread, in program order
    for (uint32_t j = wa.lane_off[lane]; j < wa.lane_off[lane + 1]; j++) {
      const uint2 e = __ldg(wa.wslots + j);
      const fr_t v = nova_wide_value(w, e.y, F);
      st_slot_fr(dst + (size_t)e.x * 32, v.l);
      // checksum: take back what expand_slots accounted for this slot (nothing for the S64 / INV kinds it skips)
      const uint32_t kind = e.y >> 24, t = e.y & 0xFFFFu;
      if (kind == DK_WIDE_BIT64) acc -= sum_small_slot(e.x, 0u, 0u);
      else if (kind < DK_S64) acc -= sum_small_slot(e.x, w.trace[t], kind == DK_W64 ? w.trace[t + 1] : 0u);
      acc += sum_field_slot(e.x, v);
    }
    if (sums) {
#pragma unroll
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) sums[i] = (unsigned long long)(acc * B3W_SUM_K);
    }
  }
}
