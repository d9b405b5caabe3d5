// nova_wide_logic.h -- the nova-level logic of Blake3Nova(0) on FIELD ELEMENTS, shared by the wide nova kernel
// (kernels_nova_wide.cuh, lane 0) and the host-side replay that produces the reference's "Assert Failed." trace
// (wide_domain.h).  circuits/blake3_nova.circom:13-167 as built (without :25-30).
#pragma once
#include <stdint.h>
#include "fr.cuh"
#include "nova_trace.h"

FR_HD bool nw_fits(const fr_t &v, int bits) {          // bits <= 96
  uint32_t hi = v.l[3] | v.l[4] | v.l[5] | v.l[6] | v.l[7];
  if (bits <= 32) return (hi | v.l[2] | v.l[1]) == 0 && (bits == 32 || (v.l[0] >> bits) == 0);
  if (bits <= 64) return (hi | v.l[2]) == 0 && (bits == 64 || (v.l[1] >> (bits - 32)) == 0);
  return hi == 0 && (v.l[2] >> (bits - 64)) == 0;
}
FR_HD fr_t nw_sub(const fr_t &a, const fr_t &b, const fr_t &p) { return fr_add(a, fr_neg(b, p), p); }
FR_HD fr_t nw_add_small(const fr_t &a, int k, const fr_t &p) { return fr_add(a, fr_from_s64(k, p), p); }

// which constraint of the nova level fails first, in the wasm's execution order (check_depth runs when leaf_depth arrives,
// :201 as built; final_m -- and inside it down_left_path's Num2Bits(65) -- when chunk_idx arrives, :221)
#define NW_FAIL_NONE 0u
#define NW_FAIL_V1 1u        /* check_parent  = LessThan(8)(depth, leaf_depth - 1): Num2Bits(9) */
#define NW_FAIL_V2 2u        /* exceed_depth  = GreaterEqThan(8)(depth, leaf_depth): Num2Bits(9) */
#define NW_FAIL_EXCEED 3u    /* exceed_depth.out === 0 */
#define NW_FAIL_N2B65 4u     /* Num2Bits(65)(chunk_idx_low + 2^32 chunk_idx_high) */

struct nova_wide_scalars {
  uint32_t fail;
  uint32_t v1, v2, is_parent, is_root, not_root, not_parent, first, last, is_last, first_set, urf_tmp, urf, dlp, cdd, decr, dflags;
  uint32_t S[3];             // chunk_idx = low + 2^32 high (< 2^65 when fail == 0)
  uint64_t eq, bad;          // bit i = eqs[i].out / bit_at_depth[i]
};

// input accessors: 32 canonical field elements as an fr_t array / as 8-limb rows of u32
struct nw_in_array {
  const fr_t *v;
  FR_HD fr_t operator()(uint32_t k) const { return v[k]; }
};
struct nw_in_words {
  const uint32_t *w;
  FR_HD fr_t operator()(uint32_t k) const {
    fr_t r;
    for (int j = 0; j < 8; j++) r.l[j] = w[8 * k + j];
    return r;
  }
};

// in(k): the k-th input as a canonical field element, declaration order (n_blocks block_count h[8] chunk_idx_low
// chunk_idx_high leaf_depth total_depth depth m[16] b)
template <class In>
FR_HD nova_wide_scalars nova_wide_scalar_logic(const In &in, const fr_t &p) {
  nova_wide_scalars s;
  const fr_t n_blocks = in(0), bc = in(1), low = in(10), high = in(11), leaf = in(12), total = in(13), depth = in(14);
  const fr_t v1 = nw_add_small(nw_sub(depth, leaf, p), 257, p);            // depth + 256 - (leaf_depth - 1)      (:31-33)
  const fr_t v2 = nw_add_small(nw_sub(leaf, depth, p), 255, p);            // leaf_depth + 256 - (depth + 1)      (:41-43)
  s.fail = NW_FAIL_NONE;
  if (!nw_fits(v1, 9)) s.fail = NW_FAIL_V1;
  else if (!nw_fits(v2, 9)) s.fail = NW_FAIL_V2;
  else if (((v2.l[0] >> 8) & 1u) == 0) s.fail = NW_FAIL_EXCEED;
  s.v1 = v1.l[0]; s.v2 = v2.l[0];
  s.is_parent = 1u - ((v1.l[0] >> 8) & 1u);
  s.is_root = fr_is_zero(depth) ? 1u : 0u;
  s.not_root = 1u - s.is_root; s.not_parent = 1u - s.is_parent;
  s.first = fr_is_zero(bc) ? 1u : 0u;                                     // Blake3GetFlag (:122-167)
  s.last = fr_is_zero(nw_sub(nw_add_small(n_blocks, -1, p), bc, p)) ? 1u : 0u;
  s.is_last = s.last & s.not_parent; s.first_set = s.first & s.not_parent;
  s.urf_tmp = s.is_parent | s.last; s.urf = s.urf_tmp & s.is_root;
  s.dflags = s.first_set + 2u * s.is_last + 8u * s.urf + 4u * s.is_parent;
  // Blake3GetDownLeftPath (:47-84): eqs[i].out = (total_depth - i - 2 == depth)  <=>  total_depth - depth == i + 2
  const fr_t x = nw_sub(total, depth, p);
  s.eq = 0;
  if (nw_fits(x, 7) && x.l[0] >= 2u && x.l[0] <= 65u) s.eq = 1ull << (x.l[0] - 2u);
  fr_t h32 = high;                                                        // 2^32 * high by 32 doublings
  for (int i = 0; i < 32; i++) h32 = fr_add(h32, h32, p);
  const fr_t S = fr_add(low, h32, p);
  if (s.fail == NW_FAIL_NONE && !nw_fits(S, 65)) s.fail = NW_FAIL_N2B65;
  s.S[0] = S.l[0]; s.S[1] = S.l[1]; s.S[2] = S.l[2];
  const uint64_t s64 = ((uint64_t)S.l[1] << 32) | S.l[0];
  const uint64_t mask = s.eq & ~s64;                                      // (1 - n2b.out[i]) * eqs[i].out: at most one bit
  s.bad = mask ? ~((mask & (0 - mask)) - 1) : 0;                          // prefix sums bit_at_depth[i]            (:65,:70)
  s.dlp = s.not_parent + s.is_parent * (uint32_t)(s.bad >> 63);           // out (:79)
  s.cdd = s.is_last | s.is_parent; s.decr = s.cdd & s.not_root;           // :254-258
  return s;
}

// Blake3GetFinal_m (:86-120) and the inputs of the embedded compression: every value is one of the 32 inputs or 0.
// Input indices: h[j] = 2 + j, m[j] = 15 + j.  NW_SEL_ZERO = the value 0.
#define NW_SEL_ZERO 255u
struct nova_wide_sel { uint32_t tmp_down, m_is_parent, tmp_is_par, out_m; };
FR_HD nova_wide_sel nova_wide_select(uint32_t j, uint32_t is_parent, uint32_t dlp) {
  const uint32_t hs = 2u + (j & 7u), ms = 15u + j, mo = 15u + (j & 7u);
  nova_wide_sel r;
  r.tmp_down = j < 8 ? (dlp ? hs : NW_SEL_ZERO) : (dlp ? NW_SEL_ZERO : hs);       // h * dlp | h * (1 - dlp)
  r.m_is_parent = j < 8 ? (dlp ? hs : ms) : (dlp ? mo : hs);                       // m (1 - dlp) + tmp_down | m[j-8] dlp + tmp_down
  r.tmp_is_par = is_parent ? r.m_is_parent : NW_SEL_ZERO;
  r.out_m = is_parent ? r.m_is_parent : ms;
  return r;
}
