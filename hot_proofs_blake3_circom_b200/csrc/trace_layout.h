/* trace_layout.h -- layout of the compact per-instance trace (u32 words) shared by the kernels and the
 * offline table generator (tools/circuit_model.py; keep the two in sync).
 *
 * A witness of the reference circuits is ~24 k field elements, but all of them are bits / words /
 * carries of < 1000 distinct 32-bit values.  The kernel computes those values once per instance (the
 * "trace") and then expands them into 32-byte witness slots through a per-slot descriptor table.
 */
#pragma once
#include <stdint.h>

#define TR_ZERO 0u        /* constant 0 */
#define TR_ONE 1u         /* constant 1 (witness slot 0) */
#define TR_IN 2u          /* compression inputs h[8] m[16] t[2] b d  (circuits/blake3_compression.circom:172-176) */
#define TR_OUT 30u        /* out[16] (blake3_compression.circom:213-227) */
#define TR_HG 48u         /* 112 half-G records x 8 words, index ((round*8 + g)*2 + half):
                             +0 a' = lo32(a+b+xy)   +1 carries (bit0 u, bit1 v)   +2 d   +3 d' = rotr(d^a', R1)
                             +4 c' = lo32(c+d')     +5 carry (bit0 u)             +6 b   +7 b' = rotr(b^c', R2)
                             (blake3_compression.circom:83-99) */
#define TR_NOVA 944u      /* nova-only words */
#define TR_WORDS_MAX 1280u

/* descriptor: bits 0..15 trace index | 16..20 bit index | 24..26 kind */
#define DK_BIT 0u
#define DK_W32 1u
#define DK_W64 2u
#define DK_S64 3u         /* signed 64-bit integer trace[t] | trace[t+1] << 32, as a field element */
#define DK_INV 4u         /* inverse mod p of that signed integer (0 for 0): circomlib IsZero.inv */
