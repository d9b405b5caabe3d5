// fr.cuh -- 256-bit prime-field layer (BN254 scalar field r, and the Pallas scalar field = circom's
// "--prime vesta"), 8 x 32-bit limbs, Montgomery arithmetic.  Replaces the part of the wasm's Fr_* library
// that the nova witness needs: negation of small integers and IsZero's inverse
// (circomlib IsZero: inv <-- in != 0 ? 1/in : 0; reference call sites circuits/blake3_nova.circom:19,63,69,139).
// Everything is __host__ __device__ so that the context set-up (tables) and the kernels share one code path.
#pragma once
#include <stdint.h>

struct fr_t { uint32_t l[8]; };

struct field_consts {
  fr_t p;            // modulus
  fr_t r2;           // 2^512 mod p
  uint32_t n0;       // -p^-1 mod 2^32
  uint32_t pad[7];
  fr_t inv_small[256];   // canonical inverses of 0..255 (inv_small[0] = 0)
};

#define FR_HD __host__ __device__ __forceinline__

FR_HD fr_t fr_zero() { fr_t r; for (int i = 0; i < 8; i++) r.l[i] = 0; return r; }
FR_HD fr_t fr_from_u64(uint64_t x) { fr_t r = fr_zero(); r.l[0] = (uint32_t)x; r.l[1] = (uint32_t)(x >> 32); return r; }

// r = a + b, returns carry
FR_HD uint32_t fr_raw_add(fr_t &r, const fr_t &a, const fr_t &b) {
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (uint64_t)a.l[i] + b.l[i]; r.l[i] = (uint32_t)c; c >>= 32; }
  return (uint32_t)c;
}
// r = a - b, returns borrow
FR_HD uint32_t fr_raw_sub(fr_t &r, const fr_t &a, const fr_t &b) {
  uint32_t br = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t d = (uint64_t)a.l[i] - b.l[i] - br;
    r.l[i] = (uint32_t)d;
    br = (uint32_t)(d >> 63);
  }
  return br;
}
FR_HD bool fr_gte(const fr_t &a, const fr_t &b) {
  for (int i = 7; i >= 0; i--) if (a.l[i] != b.l[i]) return a.l[i] > b.l[i];
  return true;
}
FR_HD bool fr_is_zero(const fr_t &a) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.l[i]; return o == 0; }

FR_HD fr_t fr_add(const fr_t &a, const fr_t &b, const fr_t &p) {
  fr_t r;
  uint32_t c = fr_raw_add(r, a, b);
  if (c || fr_gte(r, p)) fr_raw_sub(r, r, p);
  return r;
}
FR_HD fr_t fr_neg(const fr_t &a, const fr_t &p) {      // p - a (0 stays 0)
  if (fr_is_zero(a)) return a;
  fr_t r;
  fr_raw_sub(r, p, a);
  return r;
}

// Montgomery product a*b*2^-256 mod p (CIOS, 32-bit limbs)
FR_HD fr_t fr_montmul(const fr_t &a, const fr_t &b, const fr_t &p, uint32_t n0) {
  uint32_t t[10];
#pragma unroll
  for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll 1
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { c += (uint64_t)a.l[j] * b.l[i] + t[j]; t[j] = (uint32_t)c; c >>= 32; }
    c += t[8]; t[8] = (uint32_t)c; t[9] = (uint32_t)(c >> 32);
    uint32_t m = t[0] * n0;
    c = (uint64_t)m * p.l[0] + t[0]; c >>= 32;
#pragma unroll
    for (int j = 1; j < 8; j++) { c += (uint64_t)m * p.l[j] + t[j]; t[j - 1] = (uint32_t)c; c >>= 32; }
    c += t[8]; t[7] = (uint32_t)c; t[8] = t[9] + (uint32_t)(c >> 32);
  }
  fr_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = t[i];
  if (t[8] || fr_gte(r, p)) fr_raw_sub(r, r, p);
  return r;
}

// canonical inverse of a canonical non-zero a: a^(p-2) (Fermat), square-and-multiply in Montgomery form
__host__ __device__ __noinline__ static fr_t fr_inv(const fr_t &a, const fr_t &p, const fr_t &r2, uint32_t n0) {
  fr_t e = p, two = fr_from_u64(2);
  fr_raw_sub(e, e, two);
  fr_t am = fr_montmul(a, r2, p, n0);                 // a*R
  fr_t one = fr_from_u64(1);
  fr_t acc = fr_montmul(one, r2, p, n0);              // R
#pragma unroll 1
  for (int i = 255; i >= 0; i--) {
    acc = fr_montmul(acc, acc, p, n0);
    if ((e.l[i >> 5] >> (i & 31)) & 1u) acc = fr_montmul(acc, am, p, n0);
  }
  return fr_montmul(acc, one, p, n0);                 // leave Montgomery form
}

// The field element of a signed 64-bit integer: x mod p.
FR_HD fr_t fr_from_s64(int64_t x, const fr_t &p) {
  if (x >= 0) return fr_from_u64((uint64_t)x);
  return fr_neg(fr_from_u64((uint64_t)0 - (uint64_t)x), p);
}

// IsZero.inv of a signed 64-bit integer: 0 -> 0, else (x mod p)^-1.  |x| < 256 comes from the table.
FR_HD fr_t fr_inv_s64(int64_t x, const field_consts &F) {
  if (x == 0) return fr_zero();
  uint64_t a = x < 0 ? (uint64_t)0 - (uint64_t)x : (uint64_t)x;
  fr_t r = a < 256 ? F.inv_small[a] : fr_inv(fr_from_u64(a), F.p, F.r2, F.n0);
  return x < 0 ? fr_neg(r, F.p) : r;
}

// host-side set-up of the constants for a prime given as 8 little-endian u32 limbs
static inline void field_consts_init(field_consts &F, const uint32_t p[8]) {
  for (int i = 0; i < 8; i++) F.p.l[i] = p[i];
  uint32_t inv = 1;
  for (int i = 0; i < 5; i++) inv *= 2u - p[0] * inv;   // p^-1 mod 2^32 (Newton)
  F.n0 = 0u - inv;
  fr_t r = fr_from_u64(1);
  for (int i = 0; i < 512; i++) r = fr_add(r, r, F.p);  // 2^512 mod p
  F.r2 = r;
  for (int i = 0; i < 7; i++) F.pad[i] = 0;
  F.inv_small[0] = fr_zero();
  for (uint32_t k = 1; k < 256; k++) F.inv_small[k] = fr_inv(fr_from_u64(k), F.p, F.r2, F.n0);
}
