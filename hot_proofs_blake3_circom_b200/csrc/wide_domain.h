// wide_domain.h -- host side of the FULL input domain of blake3_compression.  Included by blake3wit.cu only.
//
// The reference accepts any field element for every input (`normalize`, witness_calculator.js:319-323); what happens next
// is decided by the circuit's own constraints:
//   * h, t, b, d each pass through a ToBits(32) whose recomposition constraint (circuits/blake3_common.circom:142-153)
//     fails for a canonical value >= 2^32: "Assert Failed.";
//   * the message words are never range-checked (circuits/blake3_compression.circom:169-170); m[j] only enters the sums
//     add1 = Bits34(v[a] + v[b] + m[j]) (:89), which assert iff the sum, as a canonical field element, is >= 2^34
//     (circuits/blake3_common.circom:182-203).  m[0] = 2^32 or m[0] = p - 1 therefore give VALID witnesses.
// So a satisfying input has u32 h, t, b, d and message words that are signed integers m = ext * 2^32 + lo with
// ext in [-2, 3]; that is the form the WIDE kernels take (kernels_witness.cuh).  This file holds
//   (1) the conversion from Fr256 inputs to that form, marking the instances that certainly assert, and
//   (2) a replay of the circuit's range constraints in the wasm's execution order, which yields the per-template trace
//       the reference appends to "Assert Failed." (printErrorMessage, witness_calculator.js:40-43) -- error text only:
//       witness values always come from the GPU.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "fr.cuh"
#include "nova_wide_logic.h"

static inline fr_t wd_load_reduced(const uint8_t *le32, const fr_t &p) {
  fr_t v;
  memcpy(v.l, le32, 32);
  while (fr_gte(v, p)) fr_raw_sub(v, v, p);            // p > 2^253: a 256-bit value needs at most a few rounds
  return v;
}
static inline bool wd_fits(const fr_t &v, int bits) {   // bits in (32, 64)
  uint32_t hi = v.l[2] | v.l[3] | v.l[4] | v.l[5] | v.l[6] | v.l[7];
  if (bits == 32) return (hi | v.l[1]) == 0;
  return hi == 0 && (v.l[1] >> (bits - 32)) == 0;
}

// One instance: 28 reduced inputs -> u32 row + ext[16].  Returns 0 = plain u32 instance, 1 = wide message words,
// 2 = asserts for certain (ext[0] = B3W_EXT_ASSERT tells the kernel).
static inline int wd_convert_compression(const fr_t in[28], const fr_t &p, uint32_t row[28], int8_t ext[16]) {
  bool wide = false, dead = false;
  for (int k = 0; k < 28; k++) row[k] = in[k].l[0];
  memset(ext, 0, 16);
  for (int k = 0; k < 28; k++) {
    if (k >= 8 && k < 24) continue;
    if (!wd_fits(in[k], 32)) dead = true;              // h, t, b, d: ToBits(32) cannot hold it
  }
  for (int j = 0; j < 16 && !dead; j++) {
    const fr_t &v = in[8 + j];
    if (wd_fits(v, 34)) {
      ext[j] = (int8_t)v.l[1];
    } else {
      fr_t k;                                          // v = p - k: the integer -k, if k <= 2^33
      fr_raw_sub(k, p, v);
      const uint64_t k64 = ((uint64_t)k.l[1] << 32) | k.l[0];
      if (!wd_fits(k, 34) || k64 > (1ull << 33)) { dead = true; break; }
      const uint64_t x = 0ull - k64;
      row[8 + j] = (uint32_t)x;
      ext[j] = (int8_t)(int32_t)(uint32_t)(x >> 32);   // -1 or -2
    }
    if (ext[j]) wide = true;
  }
  if (dead) {
    memset(ext, 0, 16);
    ext[0] = B3W_EXT_ASSERT;
    return 2;
  }
  return wide ? 1 : 0;
}

// ---- the first failing constraint, in the wasm's execution order ---------------------------------------------------
struct wd_fail { int kind, g, half, round; };   // kind 0 none | 1 add1 (Bits34) | 2 rxor2's ToBits(v[d]) | 3 rxor4's ToBits(v[b]) | 4 XorWord2's ToBits(h[i])

static inline uint32_t wd_rotr(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }

// Components run when their last input arrives, which for these templates is program order: rounds in sequence
// (circuits/blake3_compression.circom:194,207), GS[0..7] in sequence (:155-158), half1 then half2 (:116,:121), and inside a
// HalfFunG add1 (:89), rxor2 (:91), add3 (:92), rxor4 (:94); the output XORs come last (:213-227).
static inline wd_fail wd_first_assert_compression(const fr_t in[28], const fr_t &p) {
  static const uint8_t GI[8][4] = {{0, 4, 8, 12}, {1, 5, 9, 13}, {2, 6, 10, 14}, {3, 7, 11, 15},
                                   {0, 5, 10, 15}, {1, 6, 11, 12}, {2, 7, 8, 13}, {3, 4, 9, 14}};
  static const uint8_t PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};   // circuits/blake3_common.circom:20-24
  static const uint32_t IV[4] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au};
  fr_t v[16];
  for (int i = 0; i < 8; i++) v[i] = in[i];
  for (int i = 0; i < 4; i++) v[8 + i] = fr_from_u64(IV[i]);
  for (int i = 0; i < 4; i++) v[12 + i] = in[24 + i];
  uint8_t sched[16];
  for (int j = 0; j < 16; j++) sched[j] = (uint8_t)j;
  for (int r = 0; r < 7; r++) {
    for (int g = 0; g < 8; g++) {
      const int a = GI[g][0], b = GI[g][1], c = GI[g][2], d = GI[g][3];
      for (int half = 0; half < 2; half++) {
        const int R1 = half ? 8 : 16, R2 = half ? 7 : 12;
        const fr_t sum = fr_add(fr_add(v[a], v[b], p), in[8 + sched[2 * g + half]], p);
        if (!wd_fits(sum, 34)) return wd_fail{1, g, half, r};
        if (!wd_fits(v[d], 32)) return wd_fail{2, g, half, r};
        const uint32_t a2 = sum.l[0], d2 = wd_rotr(v[d].l[0] ^ a2, R1), c2 = v[c].l[0] + d2;   // v[c] is always a u32
        if (!wd_fits(v[b], 32)) return wd_fail{3, g, half, r};
        const uint32_t b2 = wd_rotr(v[b].l[0] ^ c2, R2);
        v[a] = fr_from_u64(a2); v[b] = fr_from_u64(b2); v[c] = fr_from_u64(c2); v[d] = fr_from_u64(d2);
      }
    }
    uint8_t nx[16];
    for (int j = 0; j < 16; j++) nx[j] = sched[PERM[j]];
    memcpy(sched, nx, 16);
  }
  for (int i = 0; i < 8; i++)
    if (!wd_fits(in[i], 32)) return wd_fail{4, 0, 0, 0};
  return wd_fail{0, 0, 0, 0};
}

// Template instance numbers and line numbers as compiled into the committed wasm files (circom 2.1.6 numbers template
// instances in instantiation order; observed through the reference's own error path).  Stand-alone blake3_compression:
// id_off = 0, line_off = 0, no suffix.  Inside the nova step circuits the same templates are instances 13 higher
// (Bits34_14, ToBits_16, RotXorWordBits_18 / _20, HalfFunG_21.., SingleRound_49, XorWord2_52, Blake3Compression_53), the
// Blake3Compression lines are one higher (the nova wasm files were compiled from a blake3_compression.circom with one more
// line above the template body), and the stack ends in Blake3Nova_54 line 239 (`blake3Compression.t[0] <== ...`, the
// component's last input); identical in all three nova builds.
static inline void wd_assert_text_compression(const wd_fail &f, char *buf, size_t cap, int id_off = 0, int line_off = 0,
                                              const char *suffix = "") {
  static const int HALF1[8] = {8, 15, 18, 21, 24, 27, 30, 33}, HALF2[8] = {13, 16, 19, 22, 25, 28, 31, 34},
                   MIX[8] = {14, 17, 20, 23, 26, 29, 32, 35};
  if (!cap) return;
  buf[0] = 0;
  if (f.kind == 0) return;
  if (f.kind == 4) {
    snprintf(buf, cap, "Error in template ToBits_%d line: 153\nError in template XorWord2_%d line: 66\n"
                       "Error in template Blake3Compression_%d line: %d\n%s", 3 + id_off, 39 + id_off, 40 + id_off, 224 + line_off, suffix);
    return;
  }
  const int hid = (f.half ? HALF2[f.g] : HALF1[f.g]) + id_off;
  char head[200];
  if (f.kind == 1) snprintf(head, sizeof head, "Error in template Bits34_%d line: 201\nError in template HalfFunG_%d line: 89\n", 1 + id_off, hid);
  else if (f.kind == 2)
    snprintf(head, sizeof head, "Error in template ToBits_%d line: 153\nError in template RotXorWordBits_%d line: 62\n"
                                "Error in template HalfFunG_%d line: 91\n", 3 + id_off, 5 + id_off, hid);
  else
    snprintf(head, sizeof head, "Error in template ToBits_%d line: 153\nError in template RotXorWordBits_%d line: 62\n"
                                "Error in template HalfFunG_%d line: 94\n", 3 + id_off, 7 + id_off, hid);
  // rounds[0].inp <== init (:194) / rounds[i].out ==> rounds[i + 1].inp (:207)
  snprintf(buf, cap, "%sError in template MixFunG_%d line: %d\nError in template SingleRound_%d line: 156\n"
                     "Error in template Blake3Compression_%d line: %d\n%s", head, MIX[f.g] + id_off, f.half ? 121 : 116, 36 + id_off,
           40 + id_off, (f.round == 0 ? 194 : 207) + line_off, suffix);
}

// ---- nova step circuits: first failing constraint for ANY 32 field elements, and its text ---------------------------
// Order (components run when their last input arrives): check_depth (:201 as built), final_m with down_left_path's
// Num2Bits(65) (:221), then the embedded Blake3Compression (:239), which sees h_compression, final_m.out_m,
// t = chunk_idx * (1 - is_parent), b and comp_d.out (circuits/blake3_nova.circom:222-245).
static inline int wd_assert_text_nova(const fr_t in[32], const fr_t &p, char *buf, size_t cap) {
  if (cap) buf[0] = 0;
  const nova_wide_scalars s = nova_wide_scalar_logic(nw_in_array{in}, p);
  const char *msg = nullptr;
  if (s.fail == NW_FAIL_V1)
    msg = "Error in template Num2Bits_2 line: 38\nError in template LessThan_3 line: 96\n"
          "Error in template Blake3NovaTreePath_CheckDepth_5 line: 27\nError in template Blake3Nova_54 line: 201\n";
  else if (s.fail == NW_FAIL_V2)
    msg = "Error in template Num2Bits_2 line: 38\nError in template LessThan_3 line: 96\nError in template GreaterEqThan_4 line: 138\n"
          "Error in template Blake3NovaTreePath_CheckDepth_5 line: 37\nError in template Blake3Nova_54 line: 201\n";
  else if (s.fail == NW_FAIL_EXCEED)
    msg = "Error in template Blake3NovaTreePath_CheckDepth_5 line: 38\nError in template Blake3Nova_54 line: 201\n";
  else if (s.fail == NW_FAIL_N2B65)
    msg = "Error in template Num2Bits_11 line: 38\nError in template Blake3GetDownLeftPath_12 line: 53\n"
          "Error in template Blake3GetFinal_m_13 line: 96\nError in template Blake3Nova_54 line: 221\n";
  if (msg) {
    if (cap) snprintf(buf, cap, "%s", msg);
    return 4;
  }
  static const uint32_t IV8[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
  fr_t c[28];
  for (int j = 0; j < 8; j++) c[j] = s.is_parent ? fr_from_u64(IV8[j]) : in[2 + j];                     // h_compression (:229-233)
  for (uint32_t j = 0; j < 16; j++) {
    const uint32_t om = nova_wide_select(j, s.is_parent, s.dlp).out_m;
    c[8 + j] = om == NW_SEL_ZERO ? fr_zero() : in[om];
  }
  c[24] = s.not_parent ? in[10] : fr_zero();                                                            // t[0], t[1] (:244-245)
  c[25] = s.not_parent ? in[11] : fr_zero();
  c[26] = in[31];
  c[27] = fr_from_u64(s.dflags);
  const wd_fail f = wd_first_assert_compression(c, p);
  if (f.kind == 0) return 0;
  wd_assert_text_compression(f, buf, cap, 13, 1, "Error in template Blake3Nova_54 line: 239\n");
  return 4;
}
