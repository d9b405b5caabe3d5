// kernels_r1cs_fast.cuh -- the stand-alone R1CS check of witnesses RESIDENT IN HBM (b3w_r1cs_check_device):
//   (A.z) * (B.z) == C.z  for every row of the constraint system, exact over Fr, at the speed the witnesses can be read.
// What the reference's consumers do on the CPU with the vector they are handed: bellpepper's enforce of every row
// (rust_fold/src/utils.rs:78-85) and circom_tester's expectPass (test/blake3_hash.test.ts:36).
// Included by blake3wit.cu only, after r1cs_rows.cuh.
//
// One CTA per instance, several CTAs per SM.  The instance's witness (771 KB) is streamed from HBM exactly once with
// 256-bit loads into a COMPACT shared-memory copy: per slot one bit "holds 0 or 1" and one bit "its value", plus the
// tagged 8-byte value of every other slot in a side table found by rank -- 21 KB instead of 771 KB, because 97 % of the
// slots of these witnesses are bits.  The rows are then evaluated from that copy by a program COMPILED on the host from
// the constraint system (fp_compile, r1cs_load.h), in which the system's regularity is spent once instead of per row:
//   * booleanity rows  (a x)(b x - b w0) = 0  -> a mask of slots that must be bits: 32 rows = one AND with the is-bit map;
//   * XOR rows  (a x)(b y) = k x + k y - k o, a b = 2 k  -> runs of consecutive (x, y, o) triples: up to 32 rows = three
//     bit-field extractions from the value map and one compare;
//   * every other row -> a short list of ITEMS per linear combination: a scalar wire, or a RUN of up to 32 consecutive
//     bit wires whose coefficients double (sum 2^i b_i: a Num2Bits / Bits34 recomposition is one or two items instead of
//     33-35 terms), stored item-major in tiles of 32 rows so that a warp's loads coalesce and its trip counts are uniform;
//     tiles whose terms provably stay below 2^57 (FP_TILE_FAST) are summed in plain 64-bit arithmetic; warps take tiles
//     from a shared counter, dearest first;
//   * VIRTUAL BITS: where circom's O2 pass substituted a bit by w - sum 2^i b_i, that combination is evaluated once per
//     witness into a bit slot of its own and the rows that carried it become booleanity / XOR / short rows (r1cs_load.h);
//   * a product with ONE field-valued factor (IsZero's in * inv = 1 - out) is one 64 x 256-bit multiply-reduce; ONE value
//     beyond 2^62 alone on the right-hand side (the nova circuits' 64-bit chunk index) is eight word compares.
// Arithmetic is exact: signed 128-bit integers with a bit-length bound per term and per product; a row that cannot be
// decided that way (a genuine field element such as IsZero's inverse, a run over slots that are not all bits, a bound
// exceeded) is re-evaluated in Fr (Montgomery) by the same lane.  Rows the compiler does not take (coefficients that are
// arbitrary field elements) stay with the general class/block evaluator of r1cs_rows.cuh as a RESIDUAL set; the built-in
// systems compile completely.  Any satisfied system is accepted and the smallest violated row id reported, whatever the
// witness holds; a slot >= p is reported as B3W_NOT_CANONICAL.
// The kernel's code is kept SMALL on purpose (cold paths out of line, one loop body for A, B and C): the four-to-eight
// CTAs of an SM are in different phases, and once their hot code exceeded the 32 KB instruction cache ncu showed 6.6
// warps per issue waiting for instructions (profiles/r02m).
// History (profiles/): 8 bytes per slot, one CTA per SM 1.65 M witnesses/s (compression, 24 544 rows) -> compact copy +
// row blocks 2.33 M/s (0.27 of the read roofline; row arithmetic issue-bound) -> compiled program 7.8 M/s (nova O2 5.1)
// -> virtual bits, fast tiles, dynamic hand-out, 128-thread CTAs: 8.2 M/s (0.96), nova O2 7.9 M/s (0.90), O1 6.8 (0.82)
// -> rotating-register streaming loop, instances from a global counter, side-table places from the circuit's layout
// instead of a shared-memory counter, the copy addressed through the shared array by name at constant offsets:
// 8.8 M/s (1.03; 12.4 M/s from compressible buffers), nova O2 8.3 (0.95), O1 7.5 (0.90); nine CTAs per SM on a 16-bit rank
// table: nova O2 8.5 (0.97); the nova circuits' 64-bit chunk index decided by a compare (fp_big_equals) instead of the Fr
// evaluator, tiles with signed / field-valued scalars (by slot kind) compiled for the exact path at once: O2 8.6 (0.985), O1 8.2 (0.99).
// Measured and dropped on the way (profiles/r02z_*): L2 prefetch of the next instance, staggered CTA starts, 9 and 10 CTAs
// per SM (the program tables lose their L1), an unrolled tile loop (instruction cache), table entries and tile headers
// loaded one step ahead in the row pass (the extra live registers spill: +1..2 %).  Kept: the booleanity / XOR loops without
// a call inside (unsettled steps are noted in a mask and redone out of line afterwards): -1.5 % on the nova systems.
#pragma once

#ifndef FPK_EXP
#define FPK_EXP 0                 /* experiment builds only: 1 = stream + classify, no rows */
#endif
#ifndef FPK_THREADS
#define FPK_THREADS 128           /* 128-thread CTAs: same warps as 256-thread ones at half the count, the phases of more instances interleave (profiles/r02t) */
#endif
#ifndef FPK_CTAS_PER_SM
#define FPK_CTAS_PER_SM 9         /* 56 registers per thread; with the 16-bit rank table nine copies need no larger shared-memory carve-out
                                     than eight did (132 KB for compression / nova O2, 164 KB for O1), so the program tables keep their L1:
                                     nova O2 -2.9 % time, the others +-0 (profiles/r02z_r1cs_sweep_9ctas.jsonl; before, 9 and 10 CTAs lost) */
#endif
#ifndef FPK_INFLIGHT
#define FPK_INFLIGHT 4            /* 256-bit loads per lane kept in flight by the streaming loop (rotating registers; 3..5) */
#endif

struct fp_item {                  // 16 bytes
  uint32_t wire;                  // scalar: the wire; run: its first wire
  uint32_t meta;                  // bits 0..5 run length (0 = scalar, 2..32 = run) | 8..15 shift | 16..23 bit length of |coef| + shift
                                  // | 24..29 scalars: log2 bound of the value the 64-bit tile path assumes (1 or FP_FAST_VBITS)
  long long coef;                 // value contributes  (coef * v) << shift;  a run's v = sum 2^j bit_j
};
struct fp_tile {                  // 32 rows of identical item counts; lane = row
  uint32_t item_off;              // items at item_off + k * 32 + lane, k < nA + nB + nC
  uint32_t row_off;               // row ids at row_ids[row_off + lane]
  uint16_t nA, nB, nC, rows;      // rows: 1..32 | FP_TILE_FAST
};
// FP_TILE_FAST (set by fp_compile): with every scalar below 2^FP_FAST_VBITS and every run over bits, no term of any row of
// the tile reaches 2^57 and no linear combination has more than 16 items -- plain 64-bit sums are exact, no bounds tracked
#define FP_TILE_FAST 0x8000u
#define FP_FAST_VBITS 36
struct fp_xor { uint32_t x, y, o, len_id; };      // len = len_id & 63 rows (x + j, y + j, o + j); ids at xor_ids[(len_id >> 6) + j]
struct fastprog_dev {
  uint32_t *bool_mask;            // (ws + 31) / 32 + 1 words: slots with a booleanity row
  uint32_t *bool_row;             // per slot: id of that row (read on failure only)
  fp_xor *xors;
  uint32_t *xor_ids;
  fp_tile *tiles;
  fp_tile *vtiles;                // virtual-bit definitions, 32 per group: nA items each, then one item holding the unit (r1cs_load.h)
  fp_item *items;
  uint32_t *row_ids;
  uint32_t *side_rank;            // per 32-slot word (+ virtual-bit words + 1, like the maps): side-table entries the circuit's slot kinds reserve before it
  uint32_t n_xors, n_tiles, n_vtiles, n_rows;      // n_rows: rows the program covers (all kinds)
  uint32_t side_total;            // entries of the side table = non-bit slots of the circuit's witness layout
};

// 32-byte slot -> tagged 8-byte value; BIG = "genuine field element", payload = slot index
__device__ __forceinline__ uint64_t cpt_classify(const uint32_t x[8], uint32_t slot, const fr_t &p, bool &noncanon) {
  if ((x[2] | x[3] | x[4] | x[5] | x[6] | x[7]) == 0 && (x[1] >> 30) == 0) return ((uint64_t)x[1] << 32) | x[0];
  fr_t v, d;
#pragma unroll
  for (int j = 0; j < 8; j++) v.l[j] = x[j];
  const uint32_t borrow = fr_raw_sub(d, p, v);               // p - x: a small negative integer stored canonically?
  noncanon = noncanon || borrow || fr_is_zero(d);            // x >= p: not a canonical field element
  if (!borrow && (d.l[2] | d.l[3] | d.l[4] | d.l[5] | d.l[6] | d.l[7]) == 0 && (d.l[1] >> 30) == 0)
    return STG_TAG_NEG | ((uint64_t)d.l[1] << 32) | d.l[0];
  return STG_TAG_BIG | slot;
}
// out-of-line form for the streaming loop (3 % of the slots take it): keeps the loop's register footprint small, so that
// several 256-bit loads per lane can be in flight.  Returns the tagged value; bit 0 of *flags is set for a slot >= p.
__device__ __noinline__ uint64_t cpt_classify_slow(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, uint32_t x5, uint32_t x6,
                                                   uint32_t x7, uint32_t slot, const field_consts *__restrict__ F, uint32_t *flags) {
  const uint32_t x[8] = {x0, x1, x2, x3, x4, x5, x6, x7};
  bool noncanon = false;
  const uint64_t v = cpt_classify(x, slot, F->p, noncanon);
  if (noncanon) atomicOr(flags, 1u);
  return v;
}
// one witness slot with a single 256-bit load (SASS LDG.E.256); streamed: read once, kept out of L1
__device__ __forceinline__ void ld_slot_stream(const uint8_t *p, uint32_t x[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]) : "l"(p));
}
// the same under a predicate (x keeps its contents when `on` is false): no branch around the load, the registers stay registers
__device__ __forceinline__ void ld_slot_stream_if(const uint8_t *p, uint32_t (&x)[8], bool on) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %9, 0;\n\t@q ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t}"
               : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]) : "l"(p), "r"((uint32_t)on));
}
__device__ __forceinline__ void ld_slot(const uint8_t *p, uint32_t x[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p)), b = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}

// ---- the kernel's shared memory: the compact copy (dynamic) and four control words.  Every access goes through fp_smem
// BY NAME: a pointer handed to an out-of-line function is a generic address to the compiler (LD.E / ST.E with 64-bit
// address arithmetic instead of LDS / STS -- the first compiled row pass had 619 generic loads and 13 LDS) ----
extern __shared__ __align__(16) uint8_t fp_smem[];
__shared__ uint32_t fp_s_bad, fp_s_flags, fp_s_tile;
__shared__ unsigned long long fp_s_it;
#define FP_MAPW 784u               /* words per map: >= (slots + 31) / 32 + virtual-bit words + 1 of every system (checked on the host); a
                                      CONSTANT so that no accessor needs a value from the (stack-resident) argument structs */
#define FP_SIDE_OFF (FP_MAPW * 10u) /* bytes: two maps of 32-bit words, one of 16-bit ranks */
struct fp_copy {                  // isbit | bitval (FP_MAPW words each) | rank (16 bits per word: the launch's constant side_rank) | side (8-byte entries)
  uint32_t *isbit, *bitval;
  uint16_t *rank;
  uint64_t *side;
  __device__ __forceinline__ fp_copy() {
    isbit = reinterpret_cast<uint32_t *>(fp_smem);
    bitval = isbit + FP_MAPW;
    rank = reinterpret_cast<uint16_t *>(bitval + FP_MAPW);
    side = reinterpret_cast<uint64_t *>(fp_smem + FP_SIDE_OFF);
  }
};

struct CompactSrc {
  // several CTAs per SM walk the same tables: let them live in L1
  static __device__ __forceinline__ uint32_t ld_table(const uint32_t *p) { return __ldg(p); }
  // the copy is found in fp_smem by name, at constant offsets (see fp_copy)
  static __device__ __forceinline__ const uint32_t *isbit() { return reinterpret_cast<const uint32_t *>(fp_smem); }      // one bit per slot: holds 0 or 1
  static __device__ __forceinline__ const uint32_t *bitval() { return isbit() + FP_MAPW; }                               // ... its value
  static __device__ __forceinline__ const uint16_t *rank() { return reinterpret_cast<const uint16_t *>(isbit() + 2u * FP_MAPW); }   // side-table base of each 32-slot word
  static __device__ __forceinline__ const uint64_t *side() { return reinterpret_cast<const uint64_t *>(fp_smem + FP_SIDE_OFF); }
  // false: an irregular instance (non-bit slots where the circuit's layout has none): such values are read from HBM
  static __device__ __forceinline__ bool side_ok() { return (fp_s_flags & 4u) == 0; }
  const uint8_t *wit;                        // this instance's witness in HBM
  const field_consts *F;
  __device__ __forceinline__ uint64_t get(uint32_t s) const {
    const uint32_t w = s >> 5, b = s & 31u, m = isbit()[w];
    if ((m >> b) & 1u) return (bitval()[w] >> b) & 1u;
    if (side_ok()) return side()[rank()[w] + __popc(~m & ((1u << b) - 1u))];
    bool nc = false;                         // (a non-canonical slot was already reported by the streaming pass)
    uint32_t x[8];
    ld_slot(wit + (size_t)s * 32, x);
    return cpt_classify(x, s, F->p, nc);
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
      ld_slot(wit + (size_t)s * 32, r.l);
      return r;
    }
    r = fr_from_u64(x & STG_PAYLOAD);
    return (x & STG_TAG_NEG) ? fr_neg(r, F->p) : r;
  }
  // `len` (1..32) consecutive bits of a bitmap starting at slot s (the maps carry one padding word)
  static __device__ __forceinline__ uint32_t field_of(const uint32_t *map, uint32_t s, uint32_t len) {
    const uint32_t w = s >> 5, v = __funnelshift_r(map[w], map[w + 1], s & 31u);
    return len >= 32u ? v : v & ((1u << len) - 1u);
  }
  __device__ __forceinline__ bool run_is_bits(uint32_t s, uint32_t len) const {
    return field_of(isbit(), s, len) == (len >= 32u ? 0xFFFFFFFFu : (1u << len) - 1u);
  }
  __device__ __forceinline__ uint32_t run_value(uint32_t s, uint32_t len) const { return field_of(bitval(), s, len); }
};

__device__ __forceinline__ int fp_bitlen64(uint64_t x) { return 64 - __clzll((long long)x); }
__device__ __forceinline__ int fp_bitlen128(i128 x) {
  const unsigned __int128 m = x < 0 ? (unsigned __int128)(-x) : (unsigned __int128)x;
  const uint64_t hi = (uint64_t)(m >> 64);
  return hi ? 64 + fp_bitlen64(hi) : fp_bitlen64((uint64_t)m);
}

// (coef << shift) as a field element; |coef| < 2^62 and bit length + shift <= 250 (guaranteed by fp_compile)
__device__ __forceinline__ fr_t fp_coef_fr(long long coef, uint32_t shift, const fr_t &p) {
  const uint64_t m = coef < 0 ? (uint64_t)0 - (uint64_t)coef : (uint64_t)coef;
  fr_t r = fr_zero();
  const uint32_t q = shift >> 5, b = shift & 31u;
  const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
  const uint32_t w0 = lo << b, w1 = b ? (hi << b) | (lo >> (32 - b)) : hi, w2 = b ? hi >> (32 - b) : 0u;
#pragma unroll
  for (int j = 0; j < 8; j++) r.l[j] = (uint32_t)j == q ? w0 : (uint32_t)j == q + 1 ? w1 : (uint32_t)j == q + 2 ? w2 : 0u;
  return coef < 0 ? fr_neg(r, p) : r;
}

// acc + (coef << shift) * v in Fr; unit coefficients (most rows that end up here: IsZero's in * inv = 1 - out) cost no product.
// Out of line, like everything on the Fr path: the instruction cache (32 KB before L2) has to hold the streaming loop and
// the row pass of the four CTAs of an SM at the same time -- ncu showed the kernel stalled on instruction fetch
// (no_instruction 6.6 warps per issue) when these bodies were inlined at every call site.
__device__ __noinline__ fr_t fp_montmul(const fr_t &a, const fr_t &b, const field_consts &F) { return fr_montmul(a, b, F.p, F.n0); }
__device__ __noinline__ fr_t fp_acc_fr(const fr_t &acc, long long coef, uint32_t shift, const fr_t &v, const field_consts &F) {
  if (shift == 0 && coef == 1) return fr_add(acc, v, F.p);
  if (shift == 0 && coef == -1) return fr_add(acc, fr_neg(v, F.p), F.p);
  const fr_t co = fp_coef_fr(coef, shift, F.p);
  return fr_add(acc, fp_montmul(fp_montmul(co, F.r2, F), v, F), F.p);
}
// a * b in Fr for canonical (non-Montgomery) operands
__device__ __forceinline__ fr_t fp_mul_fr(const fr_t &a, const fr_t &b, const field_consts &F) {
  return fp_montmul(fp_montmul(a, b, F), F.r2, F);
}

// exact evaluation of one compiled row in Fr (slow path)
__device__ __noinline__ bool fp_row_fr(const CompactSrc &src, const fp_item *__restrict__ items, uint32_t item_off, uint32_t lane,
                                       uint32_t nA, uint32_t nB, uint32_t nC) {
  const field_consts &F = *src.F;
  const uint32_t nt = nA + nB + nC;
  fr_t L[3] = {fr_zero(), fr_zero(), fr_zero()};
#pragma unroll 1
  for (uint32_t k = 0; k < nt; k++) {
    const fp_item it = items[item_off + k * 32u + lane];
    const uint32_t len = it.meta & 63u, shift = (it.meta >> 8) & 255u;
    if (it.coef == 0) continue;
    const bool as_run = len && src.run_is_bits(it.wire, len);   // a run over bits is one value; else wire by wire: its slots may hold anything
    const uint32_t cnt = (len && !as_run) ? len : 1u;
    const int part = k < nA ? 0 : k < nA + nB ? 1 : 2;
#pragma unroll 1
    for (uint32_t e = 0; e < cnt; e++) {
      const fr_t v = as_run ? fr_from_u64(src.run_value(it.wire, len)) : src.field(it.wire + e);
      const fr_t acc = fp_acc_fr(part == 0 ? L[0] : part == 1 ? L[1] : L[2], it.coef, shift + e, v, F);
      if (part == 0) L[0] = acc; else if (part == 1) L[1] = acc; else L[2] = acc;
    }
  }
  fr_t lhs = fr_zero();
  if (nA && nB) lhs = fp_mul_fr(L[0], L[1], F);
  bool eq = true;
#pragma unroll
  for (int j = 0; j < 8; j++) eq = eq && (lhs.l[j] == L[2].l[j]);
  return eq;
}

// 2 x y == x + y - o in Fr (an XOR row whose slots are not all bits)
__device__ __noinline__ bool fp_xor_fr(const CompactSrc &src, uint32_t x, uint32_t y, uint32_t o) {
  const field_consts &F = *src.F;
  const fr_t vx = src.field(x), vy = src.field(y), vo = src.field(o);
  fr_t xy = fp_mul_fr(vx, vy, F);
  xy = fr_add(xy, xy, F.p);
  const fr_t rhs = fr_add(fr_add(vx, vy, F.p), fr_neg(vo, F.p), F.p);
  bool eq = true;
#pragma unroll
  for (int j = 0; j < 8; j++) eq = eq && (xy.l[j] == rhs.l[j]);
  return eq;
}

// x (x - w0) == 0 in Fr (a booleanity row when wire 0 does not hold 1, or x is not a bit)
__device__ __noinline__ bool fp_bool_fr(const CompactSrc &src, uint32_t x) {
  const field_consts &F = *src.F;
  const fr_t vx = src.field(x), w0 = src.field(0);
  const fr_t d = fr_add(vx, fr_neg(w0, F.p), F.p);
  return fr_is_zero(fp_montmul(vx, d, F));
}

// the exact integer value of one item, (coef * v) << shift; `undecided` is raised when it has none in 128 bits (a run over
// non-bits, a bound past 2^118); a scalar that holds a genuine field element gives 0 and is reported in `big` as
// 1 + (coef == 1 ? 0 : coef == -1 ? 1 : 2) -- the caller may know what to do with a lone unit-coefficient one
__device__ __forceinline__ i128 fp_term(const CompactSrc &src, const uint4 raw, bool &undecided, uint32_t &big) {
  const uint32_t wire = raw.x, meta = raw.y, len = meta & 63u, shift = (meta >> 8) & 255u, cbits = (meta >> 16) & 255u;
  const long long coef = (long long)(((uint64_t)raw.w << 32) | raw.z);
  long long v;
  int vbits;
  big = 0;
  if (len) {
    undecided = undecided || !src.run_is_bits(wire, len);
    v = (long long)src.run_value(wire, len);
    vbits = (int)len;
  } else {
    const uint64_t x = src.get(wire);
    if (x & STG_TAG_BIG) {
      big = (shift == 0 && coef == 1) ? 1u : (shift == 0 && coef == -1) ? 2u : 3u;
      return 0;
    }
    const uint64_t mag = x & STG_PAYLOAD;
    v = (x & STG_TAG_NEG) ? -(long long)mag : (long long)mag;
    vbits = fp_bitlen64(mag);
  }
  const int bits = (int)cbits + vbits;
  if (bits <= 62) return (i128)((coef * v) << shift);
  undecided = undecided || bits > 118;                          // <= 255 items of < 2^118 each stay below 2^126
  return ((i128)coef * (i128)v) << shift;
}

// The virtual bits of one instance (r1cs_load.h: fp_find_virtuals): every definition's linear combination L is evaluated
// exactly and must be 0 or its unit u; the bit goes into the maps at slot 32 * words + index, where the rewritten rows
// read it like any other bit.  A definition that is neither (or cannot be decided in integers), or a witness whose slot 0
// is not 1, raises bit 1 of *flags: the caller then evaluates the instance with the program compiled without virtual bits.
__device__ __noinline__ void fp_eval_virtuals(const CompactSrc &src, const fastprog_dev &P, uint32_t words) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const fp_copy m;
  uint32_t *const isbit = m.isbit, *const bitval = m.bitval, *const flags = &fp_s_flags;
  if (tid == 0 && src.get(0) != 1ull) atomicOr(flags, 2u);
  for (uint32_t g = tid >> 5; g < P.n_vtiles; g += FPK_THREADS / 32) {
    const uint4 h = __ldg(reinterpret_cast<const uint4 *>(P.vtiles) + g);
    const uint32_t n = h.z & 0xFFFFu, rows = h.w >> 16;
    const fp_item *__restrict__ it0 = P.items + h.x + lane;
    i128 acc = 0;
    bool undecided = false;
#pragma unroll 1
    for (uint32_t k = 0; k < n; k++) {
      uint32_t big;
      acc += fp_term(src, __ldg(reinterpret_cast<const uint4 *>(it0 + k * 32u)), undecided, big);
      undecided = undecided || big != 0;
    }
    const uint4 ur = __ldg(reinterpret_cast<const uint4 *>(it0 + n * 32u));
    const i128 unit = (i128)(long long)(((uint64_t)ur.w << 32) | ur.z) << ((ur.y >> 8) & 255u);
    const bool in = lane < rows, one = in && !undecided && acc == unit, zero = !undecided && acc == 0;
    const uint32_t mv = __ballot_sync(0xffffffffu, one);
    const bool valid = __all_sync(0xffffffffu, !in || one || zero);
    if (lane == 0) {
      isbit[words + g] = 0xFFFFFFFFu;
      bitval[words + g] = mv;
      if (!valid) atomicOr(flags, 2u);
    }
  }
}

// an XOR run in which some row is violated or holds non-bits, row by row (cold: kept out of the row pass's code)
__device__ __noinline__ uint32_t fp_xor_run_slow(const CompactSrc &src, const fastprog_dev &P, const uint4 e) {
  const uint32_t len = e.w & 63u;
  uint32_t bad = B3W_NO_ROW;
  for (uint32_t j = 0; j < len; j++) {
    const uint64_t vx = src.get(e.x + j), vy = src.get(e.y + j), vo = src.get(e.z + j);
    const bool holds = (vx | vy | vo) < 2ull ? vo == (vx ^ vy) : fp_xor_fr(src, e.x + j, e.y + j, e.z + j);
    if (!holds) bad = min(bad, __ldg(P.xor_ids + (e.w >> 6) + j));
  }
  return bad;
}
// the booleanity rows of one mask word that the bit map does not settle (cold)
__device__ __noinline__ uint32_t fp_bool_word_slow(const CompactSrc &src, const fastprog_dev &P, uint32_t wd, uint32_t viol, bool one_ok) {
  uint32_t bad = B3W_NO_ROW;
  while (viol) {
    const uint32_t s = wd * 32u + (uint32_t)__ffs((int)viol) - 1u;
    viol &= viol - 1u;
    if (one_ok || !fp_bool_fr(src, s)) bad = min(bad, __ldg(P.bool_row + s));
  }
  return bad;
}

// A FP_TILE_FAST tile in 64-bit arithmetic.  *ok = false when some value of the lane's row is not what the flag assumes (a
// scalar of magnitude >= 2^36 or field-valued, a run over non-bits): the caller then re-evaluates the tile exactly.
__device__ __forceinline__ uint32_t fp_eval_tile_fast(const CompactSrc &src, const fastprog_dev &P, const fp_tile t, uint32_t lane, bool *ok_out) {
  const fp_item *__restrict__ it0 = P.items + t.item_off + lane;
  const uint32_t nA = t.nA, nAB = (uint32_t)t.nA + t.nB, nt = nAB + t.nC;
  long long L[3] = {0, 0, 0};
  bool ok = true;
  uint4 nxt = nt ? __ldg(reinterpret_cast<const uint4 *>(it0)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
  for (uint32_t k = 0; k < nt; k++) {                           // one loop body for the three linear combinations (code size)
    const uint4 raw = nxt;
    if (k + 1 < nt) nxt = __ldg(reinterpret_cast<const uint4 *>(it0 + (k + 1) * 32u));
    const uint32_t wire = raw.x, len = raw.y & 63u, shift = (raw.y >> 8) & 255u;
    const long long coef = (long long)((((uint64_t)raw.w << 32) | raw.z) << shift);
    long long v;
    if (len) {
      ok = ok && src.run_is_bits(wire, len);
      v = (long long)src.run_value(wire, len);
    } else {
      // a small value of either sign (differences such as depth - leaf_depth - k are negative in every witness); its
      // magnitude must stay below the bound fp_compile assumed for this wire, and it must not be a genuine field element
      const uint64_t x = src.get(wire), mag = x & STG_PAYLOAD;
      ok = ok && !(x & STG_TAG_BIG) && (mag >> ((raw.y >> 24) & 63u)) == 0;
      v = (x & STG_TAG_NEG) ? -(long long)mag : (long long)mag;
    }
    const long long term = coef * v;
    if (k < nA) L[0] += term; else if (k < nAB) L[1] += term; else L[2] += term;
  }
  *ok_out = ok;
  const bool holds = (t.nA && t.nB) ? (i128)L[0] * (i128)L[1] == (i128)L[2] : L[2] == 0;
  return (holds || lane >= (t.rows & 63u)) ? B3W_NO_ROW : P.row_ids[t.row_off + lane];
}

// a * W == c (mod p) for the field element W in witness slot `wire` (canonical: the streaming pass has checked) and small
// integers a, c:  T = |a| W + E with E = -+c mod p is a multiple of p  <=>  two Montgomery steps leave 0 or p.
// 40 multiply-adds instead of the two full Montgomery products of the general Fr evaluator.
__device__ __noinline__ bool fp_small_times_big(const CompactSrc &src, uint32_t wire, long long a, long long c) {
  const field_consts &F = *src.F;
  uint32_t w[8];
  ld_slot(src.wit + (size_t)wire * 32, w);
  const uint64_t ma = a < 0 ? (uint64_t)0 - (uint64_t)a : (uint64_t)a;
  const long long e = a < 0 ? c : -c;                         // a W = c  <=>  |a| W + e = 0  (a > 0: e = -c;  a < 0: e = c)
  fr_t E = fr_from_u64(e < 0 ? (uint64_t)0 - (uint64_t)e : (uint64_t)e);
  if (e < 0) E = fr_neg(E, F.p);
  uint32_t t[11];
#pragma unroll
  for (int j = 0; j < 8; j++) t[j] = E.l[j];
  t[8] = t[9] = t[10] = 0;
#pragma unroll
  for (int i = 0; i < 2; i++) {                                 // T += |a| W
    const uint32_t ai = (uint32_t)(ma >> (32 * i));
    uint64_t cy = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { cy += (uint64_t)ai * w[j] + t[i + j]; t[i + j] = (uint32_t)cy; cy >>= 32; }
#pragma unroll
    for (int j = i + 8; j < 11; j++) { cy += t[j]; t[j] = (uint32_t)cy; cy >>= 32; }
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {                                 // T += m p 2^(32 i), clearing limb i
    const uint32_t m = t[i] * F.n0;
    uint64_t cy = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { cy += (uint64_t)m * F.p.l[j] + t[i + j]; t[i + j] = (uint32_t)cy; cy >>= 32; }
#pragma unroll
    for (int j = i + 8; j < 11; j++) { cy += t[j]; t[j] = (uint32_t)cy; cy >>= 32; }
  }
  // T / 2^64 < 2 p: a multiple of p is 0 or p
  uint32_t any = t[10], diff = t[10];
#pragma unroll
  for (int j = 0; j < 8; j++) { any |= t[2 + j]; diff |= t[2 + j] ^ F.p.l[j]; }
  return any == 0 || diff == 0;
}

// W == v (mod p) for the field element W in witness slot `wire` (canonical: the streaming pass has checked) and a signed
// 128-bit integer v: eight word compares.  Decides a row whose ONE value beyond 2^62 stands alone with a unit coefficient
// on the right-hand side -- the 64-bit chunk index of the nova circuits in `low + 2^32 high = idx` and in Num2Bits(65)'s
// recomposition -- which the general Fr evaluator settled with two Montgomery products per item.
__device__ __noinline__ bool fp_big_equals(const CompactSrc &src, uint32_t wire, i128 v) {
  uint32_t w[8];
  ld_slot(src.wit + (size_t)wire * 32, w);
  const bool neg = v < 0;
  const unsigned __int128 m = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
  fr_t e = fr_zero();
  e.l[0] = (uint32_t)m; e.l[1] = (uint32_t)(m >> 32); e.l[2] = (uint32_t)(m >> 64); e.l[3] = (uint32_t)(m >> 96);
  if (neg && m != 0) e = fr_neg(e, src.F->p);
  bool eq = true;
#pragma unroll
  for (int j = 0; j < 8; j++) eq = eq && w[j] == e.l[j];
  return eq;
}

// one tile: lane = row.  Returns the lane's violated row id or B3W_NO_ROW.
// A term (coef * v) << shift whose bit-length bound stays <= 62 is formed in 64 bits, the others -- 2^32.. coefficients
// on wide values, Num2Bits(65)'s top run -- in 128; the sums are 128-bit.
__device__ __noinline__ uint32_t fp_eval_tile(const CompactSrc &src, const fastprog_dev &P, const fp_tile t, uint32_t lane) {
  const bool active = lane < (t.rows & 63u);
  const fp_item *__restrict__ it0 = P.items + t.item_off + lane;
  i128 L[3] = {0, 0, 0};
  bool undecided = false;
  uint32_t nbig = 0, big_wire = 0, big_kind = 0;
  const uint32_t nA = t.nA, nAB = (uint32_t)t.nA + t.nB, nt = nAB + t.nC;
  uint4 nxt = nt ? __ldg(reinterpret_cast<const uint4 *>(it0)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
  for (uint32_t k = 0; k < nt; k++) {                           // one loop body for the three linear combinations (code size)
    const uint4 raw = nxt;
    if (k + 1 < nt) nxt = __ldg(reinterpret_cast<const uint4 *>(it0 + (k + 1) * 32u));      // next item in flight (the tables live in L1 / L2)
    uint32_t big;
    const i128 term = fp_term(src, raw, undecided, big);
    const uint32_t part = k < nA ? 0u : k < nAB ? 1u : 2u;
    if (big) {                                                  // remember one field-valued operand: wire, sign, which side
      nbig++;
      big_wire = raw.x;
      big_kind = big | (part << 4);
    }
    if (part == 0) L[0] += term; else if (part == 1) L[1] += term; else L[2] += term;
  }
  if (!active) return B3W_NO_ROW;
  bool holds;
  if (nbig) {
    // IsZero's  in * inv = 1 - out  and its kin: one factor IS a field-valued slot (unit coefficient, nothing else on its
    // side), the other factor and the right-hand side are small integers -> one 64 x 256-bit product (fp_small_times_big)
    const uint32_t side = big_kind >> 4;
    const i128 other = side == 0 ? L[1] : L[0];
    const bool lone = !undecided && nbig == 1 && (big_kind & 15u) <= 2u && side <= 1u && t.nA && t.nB && (side == 0 ? L[0] : L[1]) == 0 &&
                      fp_bitlen128(other) <= 62 && fp_bitlen128(L[2]) <= 62;
    if (lone) return fp_small_times_big(src, big_wire, (big_kind & 15u) == 2u ? -(long long)other : (long long)other, (long long)L[2])
                         ? B3W_NO_ROW : P.row_ids[t.row_off + lane];
    // ... or it stands on the right-hand side with a unit coefficient:  A B = C' +- W  <=>  W = +-(A B - C')
    const bool quad = t.nA && t.nB;
    if (!undecided && nbig == 1 && (big_kind & 15u) <= 2u && side == 2u && (!quad || fp_bitlen128(L[0]) + fp_bitlen128(L[1]) <= 125)) {
      const i128 rest = (quad ? L[0] * L[1] : (i128)0) - L[2];
      return fp_big_equals(src, big_wire, (big_kind & 15u) == 1u ? rest : -rest) ? B3W_NO_ROW : P.row_ids[t.row_off + lane];
    }
    undecided = true;
  }
  if (t.nA == 0 || t.nB == 0) {
    holds = !undecided && L[2] == 0;
  } else {
    undecided = undecided || fp_bitlen128(L[0]) + fp_bitlen128(L[1]) > 125;
    holds = !undecided && L[0] * L[1] == L[2];
  }
  if (undecided) holds = fp_row_fr(src, P.items, t.item_off, lane, t.nA, t.nB, t.nC);
  return holds ? B3W_NO_ROW : P.row_ids[t.row_off + lane];
}

// rows the compiler did not take (coefficients that are arbitrary field elements): the general class / block evaluator
__device__ __noinline__ uint32_t fp_eval_residual(const CompactSrc &src, const r1cs_tables_dev &T, bool one_ok) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  constexpr uint32_t NW = FPK_THREADS / 32;
  uint32_t bad = B3W_NO_ROW;
  for (uint32_t ci = 0; ci < T.n_classes; ci++) {
    const r1cs_class_dev c = T.cls[ci];
    const uint32_t nb = T.cls_blocks[ci], hw = 2u + 2u * (c.nA + c.nB + c.nC);
    const bool fast = one_ok && (c.flags & R1CS_FLAG_FAST64);
    if (c.flags & R1CS_FLAG_MATRIX) {
      for (uint32_t r = tid; r < ((c.count + 31u) & ~31u); r += FPK_THREADS)
        bad = min(bad, fast ? staged_matrix_row<true>(src, c, T, r) : staged_matrix_row<false>(src, c, T, r));
    } else {
      for (uint32_t b = warp; b < nb; b += NW) {
        const uint32_t *hdr = T.terms + c.term_off + (size_t)b * hw;
        bad = min(bad, fast ? staged_block<true>(src, c, T, hdr, lane) : staged_block<false>(src, c, T, hdr, lane));
      }
    }
  }
  return bad;
}

// every row of one instance from the compact copy (all threads of the CTA; each returns its own smallest violated row id).
// Out of line: the streaming loop of the kernel keeps its registers for loads in flight.
__device__ __noinline__ uint32_t fp_eval_rows(const CompactSrc &src, const fastprog_dev &P, const r1cs_tables_dev &T, uint32_t words) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const uint32_t *isbit = src.isbit();
  uint32_t bad = B3W_NO_ROW;
  const bool one_ok = src.get(0) == 1ull;                 // wire 0 holds the constant 1
  // The two loops below contain NO call: a step that does not settle (a violation, or slots that are not bits) is only
  // noted in a mask and redone out of line afterwards, so that the compiler is free to unroll and to have the table loads
  // of several steps in flight (with the slow-path call inside, every step was a load -> use chain of an L1 / L2 latency).
  // ---- booleanity rows: the slots of the mask must be bits ----
  {
    const uint32_t nw = words + P.n_vtiles;                 // (virtual bits: a word per group of definitions; <= FP_MAPW: 7 steps)
    const uint32_t *__restrict__ bm = P.bool_mask;
    uint32_t redo = 0;
#pragma unroll 4
    for (uint32_t j = 0, wd = tid; wd < nw; j++, wd += FPK_THREADS) {
      const uint32_t need = __ldg(bm + wd);
      const uint32_t viol = one_ok ? need & ~isbit[wd] : need;       // x (x - w0) = 0 with w0 != 1: decide every row exactly
      if (viol) redo |= 1u << j;
    }
    for (; redo; redo &= redo - 1u) {                       // cold
      const uint32_t wd = tid + ((uint32_t)__ffs((int)redo) - 1u) * FPK_THREADS, need = __ldg(bm + wd);
      bad = min(bad, fp_bool_word_slow(src, P, wd, one_ok ? need & ~isbit[wd] : need, one_ok));
    }
  }
  // ---- XOR rows: runs of consecutive (x, y, o) triples ----
  {
    const uint32_t nx = P.n_xors, nf = min(nx, 32u * FPK_THREADS);   // (the mask has 32 steps; the systems in use have 4 to 7)
    const uint4 *__restrict__ xs = reinterpret_cast<const uint4 *>(P.xors);
    uint32_t redo = 0;
#pragma unroll 1
    for (uint32_t j = 0, b = tid; b < nf; j++, b += FPK_THREADS) {
      const uint4 e = __ldg(xs + b);
      const uint32_t len = e.w & 63u;
      const bool bits = src.run_is_bits(e.x, len) && src.run_is_bits(e.y, len) && src.run_is_bits(e.z, len);
      if (!(bits && (src.run_value(e.x, len) ^ src.run_value(e.y, len)) == src.run_value(e.z, len))) redo |= 1u << j;
    }
    for (; redo; redo &= redo - 1u)                         // cold: some row of the run is violated or holds non-bits: row by row
      bad = min(bad, fp_xor_run_slow(src, P, __ldg(xs + tid + ((uint32_t)__ffs((int)redo) - 1u) * FPK_THREADS)));
    for (uint32_t b = tid + 32u * FPK_THREADS; b < nx; b += FPK_THREADS)      // a loaded system with more than 4 096 runs: the rest, run by run
      bad = min(bad, fp_xor_run_slow(src, P, __ldg(xs + b)));
  }
  // ---- every other compiled row: tiles of 32 rows, one per warp step, handed out dynamically (fp_compile orders them
  // by decreasing cost: a warp that meets rows for the Fr path does not end up holding the CTA's barrier alone) ----
  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(&fp_s_tile, 1u);           // (0 at entry)
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= P.n_tiles) break;
    const uint4 h = __ldg(reinterpret_cast<const uint4 *>(P.tiles) + t);
    fp_tile tl;
    tl.item_off = h.x; tl.row_off = h.y;
    tl.nA = (uint16_t)(h.z & 0xFFFFu); tl.nB = (uint16_t)(h.z >> 16); tl.nC = (uint16_t)(h.w & 0xFFFFu); tl.rows = (uint16_t)(h.w >> 16);
    if (tl.rows & FP_TILE_FAST) {
      bool ok;
      const uint32_t r = fp_eval_tile_fast(src, P, tl, lane, &ok);
      if (__all_sync(0xffffffffu, ok)) { bad = min(bad, r); continue; }
    }
    bad = min(bad, fp_eval_tile(src, P, tl, lane));
  }
  if (T.n_classes) bad = min(bad, fp_eval_residual(src, T, one_ok));
  return bad;
}

struct fp_stream_args {
  const uint8_t *w;               // this instance's witness in HBM
  const field_consts *F;
  uint32_t ws, words;
};
// one 32-slot word of the streaming pass from the registers x (lane = slot); then the load of word wu + ahead goes into x
__device__ __forceinline__ void fp_stream_word(const fp_stream_args &c, const fp_copy &m, uint32_t lane, uint32_t (&x)[8], uint32_t wu,
                                               uint32_t ahead) {
  if (wu >= c.words) return;                                  // warp-uniform
  const uint32_t x0 = x[0], x1 = x[1], hi = x[2] | x[3] | x[4] | x[5] | x[6] | x[7];
  const uint32_t s = wu * 32u + lane;
  const bool in = s < c.ws;
  const bool bit = in && (x1 | hi) == 0 && x0 < 2u;
  const uint32_t mb = __ballot_sync(0xffffffffu, bit || !in);  // slots past the end count as bits (value 0)
  const uint32_t mv = __ballot_sync(0xffffffffu, bit && x0 == 1u);
  if (~mb) {                                                  // warp-uniform: the word holds non-bit slots
    // Their side-table entries start where the CIRCUIT's slot kinds put this word's (m.rank, constant per launch): no
    // counter, no atomic, no exchange between the warps.  A word with more non-bit slots than the circuit's layout has
    // there (never a witness of this circuit) makes the instance IRREGULAR: its rows read such values from HBM instead.
    const uint32_t r0 = m.rank[wu], cap = (uint32_t)m.rank[wu + 1u] - r0;
    const bool fits = (uint32_t)__popc(~mb) <= cap;
    if (!((mb >> lane) & 1u)) {
      // small non-negative integers (every word, sum and carry of these circuits) need no field arithmetic
      uint64_t v;
      if (hi == 0 && (x1 >> 30) == 0) v = ((uint64_t)x1 << 32) | x0;
      else v = cpt_classify_slow(x0, x1, x[2], x[3], x[4], x[5], x[6], x[7], s, c.F, &fp_s_flags);
      if (fits) m.side[r0 + __popc(~mb & ((1u << lane) - 1u))] = v;
    }
    if (!fits && lane == 0) atomicOr(&fp_s_flags, 4u);
  }
  if (lane == 0) { m.isbit[wu] = mb; m.bitval[wu] = mv; }
  const uint32_t wn = wu + ahead;                             // the registers are free: the next load goes out now
  ld_slot_stream_if(c.w + (size_t)min(wn * 32u + lane, c.ws - 1u) * 32, x, wn < c.words);
}
// The streaming pass of one instance: lane = slot inside a 32-slot word (1 KiB), the CTA's warps take the words round-robin.
// ROTATING registers: a word's slots are consumed and the load of the word FPK_INFLIGHT steps ahead goes into the same
// registers at once, so FPK_INFLIGHT - 1 .. FPK_INFLIGHT KiB per warp stay in flight THROUGH the classification work
// instead of draining to zero before the next burst (compressible buffers: 3.59 -> 2.95 ms per 2^15 compression witnesses,
// ordinary memory 4.06 -> 3.87 with the dynamic hand-out; profiles/r02z).  Out of line for a register allocation of its
// own: inside the kernel body ptxas kept the slot registers on the stack, next to the values that live across the row pass.
__device__ __noinline__ void fp_stream(const fp_stream_args c) {
  constexpr uint32_t NW = FPK_THREADS / 32, AH = FPK_INFLIGHT * NW;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const fp_copy m;
#define FPK_SLOT_REGS(x, k) \
  uint32_t x[8] = {0, 0, 0, 0, 0, 0, 0, 0}; \
  ld_slot_stream_if(c.w + (size_t)min((warp + (k) * NW) * 32u + lane, c.ws - 1u) * 32, x, warp + (k) * NW < c.words)
  FPK_SLOT_REGS(xa, 0);
  FPK_SLOT_REGS(xb, 1);
  FPK_SLOT_REGS(xc, 2);
#if FPK_INFLIGHT >= 4
  FPK_SLOT_REGS(xd, 3);
#endif
#if FPK_INFLIGHT >= 5
  FPK_SLOT_REGS(xe, 4);
#endif
#undef FPK_SLOT_REGS
#pragma unroll 1
  for (uint32_t wd = warp; wd < c.words; wd += AH) {
    fp_stream_word(c, m, lane, xa, wd, AH);
    fp_stream_word(c, m, lane, xb, wd + NW, AH);
    fp_stream_word(c, m, lane, xc, wd + 2 * NW, AH);
#if FPK_INFLIGHT >= 4
    fp_stream_word(c, m, lane, xd, wd + 3 * NW, AH);
#endif
#if FPK_INFLIGHT >= 5
    fp_stream_word(c, m, lane, xe, wd + 4 * NW, AH);
#endif
  }
}

__global__ void __launch_bounds__(FPK_THREADS, FPK_CTAS_PER_SM)
k_r1cs_check_fast(const uint8_t *__restrict__ wit, const uint32_t *__restrict__ list /* NULL, or {count, instances...}: see below */, uint64_t n,
                  uint32_t ws, const fastprog_dev P,
                  const fastprog_dev P0 /* the same rows compiled without virtual bits (= P when P has none) */,
                  const r1cs_tables_dev T /* residual rows */, const field_consts *__restrict__ F, uint8_t *__restrict__ status,
                  uint32_t *__restrict__ first_bad, unsigned long long *__restrict__ counter /* NULL: static round-robin; else 0 at launch */,
                  uint32_t skip_asserted /* leave instances whose status says "Assert Failed." alone: no witness exists for them */) {
  const uint32_t words = (ws + 31u) >> 5, mw = words + P.n_vtiles + 1u;      // virtual-bit words, one padding word per map (field_of reads w + 1)
  const fp_copy m;
  uint32_t *const isbit = m.isbit, *const bitval = m.bitval;
  const uint32_t tid = threadIdx.x;
  // list != NULL: check the instances list[1 .. list[0]] (n is ignored)
  const uint64_t count = list ? (uint64_t)list[0] : n;
  for (uint32_t k = tid; k < mw; k += FPK_THREADS) m.rank[k] = (uint16_t)__ldg(P.side_rank + k);      // (the loop's first barrier publishes it)
  // Instances are handed out through a counter in global memory (one atomic per 770 KB read): CTAs that finish early take
  // the remainder, and the phases of the CTAs (stream / rows) drift apart instead of staying in the step the launch put
  // them in (-3 % on all three systems, profiles/r02z).  counter == NULL: static round-robin.
  for (uint64_t it = blockIdx.x;; it += gridDim.x) {
    __syncthreads();                                          // the previous instance's rows are done with the copy
    if (tid == 0) {
      if (counter) fp_s_it = atomicAdd(counter, 1ull);
      fp_s_bad = B3W_NO_ROW; fp_s_flags = 0; fp_s_tile = 0; isbit[mw - 1u] = 0xFFFFFFFFu; bitval[mw - 1u] = 0u;
    }
    __syncthreads();
    if (counter) it = fp_s_it;
    if (it >= count) break;
    const uint64_t i = list ? (uint64_t)list[1 + it] : it;
    if (skip_asserted && status && status[i] == B3W_CIRCOM_ASSERT) {      // CTA-uniform: no witness exists for this instance
      if (first_bad && tid == 0) first_bad[i] = B3W_NO_ROW;
      continue;
    }
    const uint8_t *w = wit + i * (uint64_t)ws * 32;
    // ---- stream the witness once into the compact copy ----
    fp_stream(fp_stream_args{w, F, ws, words});
    __syncthreads();
    uint32_t bad = B3W_NO_ROW;
    if (!(fp_s_flags & 1u) && FPK_EXP == 0) {
      const CompactSrc src{w, F};
      if (P.n_vtiles) {                                       // CTA-uniform
        fp_eval_virtuals(src, P, words);
        __syncthreads();
      }
      bad = (fp_s_flags & 2u) ? fp_eval_rows(src, P0, T, words) : fp_eval_rows(src, P, T, words);
    }
    if (bad != B3W_NO_ROW) atomicMin(&fp_s_bad, bad);
    __syncthreads();
    if (tid == 0) {
      const uint32_t verdict = (fp_s_flags & 1u) ? B3W_NOT_CANONICAL : fp_s_bad;
      if (status) status[i] = verdict == B3W_NO_ROW ? 0 : B3W_R1CS_VIOLATION;
      if (first_bad) first_bad[i] = verdict;
    }
  }
}
