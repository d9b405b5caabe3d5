// kernels_witness.cuh -- the witness kernels: trace phase (native u32 BLAKE3 + the nova step logic), expansion phase
// (slot descriptors -> 256-bit streaming stores), dynamic work-item scheduling, fused R1CS check on checker warps.
// Included by blake3wit.cu only (one translation unit).
#pragma once
// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rotr32(uint32_t x, int r) { return __funnelshift_r(x, x, r); }

// Cache policy of the witness stores: written once and never re-read by this kernel, so they bypass L1 and are marked
// evict-first in L2 (keeps the descriptor / field tables resident there; +5 % on the nova kernel, +0.3 % on compression).
// B3W_ST_HINT is an experiment switch; 1 is what ships.
#ifndef B3W_ST_HINT
#define B3W_ST_HINT 1
#endif
#if B3W_ST_HINT == 0
#define B3W_ST_QUAL ".L1::no_allocate"
#elif B3W_ST_HINT == 1
#define B3W_ST_QUAL ".L1::no_allocate.L2::evict_first"
#else
#define B3W_ST_QUAL ".cs"
#endif
// 256-bit streaming store of one witness slot {w0..w7}: written once, never re-read by this kernel.
__device__ __forceinline__ void st_slot(void *p, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4,
                                        uint32_t w5, uint32_t w6, uint32_t w7) {
  asm volatile("st.global" B3W_ST_QUAL ".v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w0), "r"(w1), "r"(w2),
               "r"(w3), "r"(w4), "r"(w5), "r"(w6), "r"(w7)
               : "memory");
}

// A full 8-limb field element goes out as two 128-bit stores.  (ptxas 12.9 mis-handles the live ranges of a
// v8.b32 store whose eight operands are all computed values inside a non-inlined function and keeps only the first
// limb -- seen in SASS as a 32-bit STG; the {lo, hi, 0...} form above is not affected.  tests/test_gpu_nova.py pins it.)
__device__ __forceinline__ void st_slot_fr(void *p, const uint32_t *l) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0+16], {%1,%2,%3,%4};" ::"l"(p), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
}

// BLAKE3 message schedule: MSG_SCHED[r][j] = index into the original m[] of the word that round r
// sees at position j, i.e. sigma applied r times (circuits/blake3_common.circom:20-24,
// circuits/blake3_compression.circom:198-209).
__constant__ uint8_t MSG_SCHED[7][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
    {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8},
    {3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1},
    {10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6},
    {12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4},
    {9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7},
    {11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13}};

// One HalfFunG (circuits/blake3_compression.circom:72-100) on this lane's (a,b,c,d); lanes 0..3 record it.
template <int R1, int R2>
__device__ __forceinline__ void half_g(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d, uint32_t xy, uint32_t *rec,
                                       bool writer) {
  uint32_t s = a + b;
  uint32_t hi1 = (s < a);
  uint32_t s2 = s + xy;
  hi1 += (s2 < s);                       // add1 = Bits34(v[a]+v[b]+xy): carries u = bit0, v = bit1   (:83,:88)
  uint32_t d_old = d;
  d = rotr32(d ^ s2, R1);                // rxor2 = RotXorWordBits(R1)(v[d], add1.out_bits)            (:89-90)
  uint32_t t = c + d;
  uint32_t hi3 = (t < c);                // add3 = Bits33(v[c] + rxor2.out_word)                        (:91)
  uint32_t b_old = b;
  b = rotr32(b ^ t, R2);                 // rxor4 = RotXorWordBits(R2)(v[b], add3.out_bits)            (:92-93)
  a = s2;
  c = t;
  if (writer) {
    *reinterpret_cast<uint4 *>(rec) = make_uint4(a, hi1, d_old, d);
    *reinterpret_cast<uint4 *>(rec + 4) = make_uint4(c, hi3, b_old, b);
  }
}

// This lane's slice of the message schedule, packed for registers: word r holds the four m[] indices that lane
// q = lane & 3 needs in round r (columns: msg[2q], msg[2q+1]; diagonals: msg[8+2q], msg[9+2q]), one byte each.
struct lane_sched { uint32_t w[7]; };
__device__ __forceinline__ lane_sched load_lane_sched(int lane) {
  const int q = lane & 3;
  lane_sched ls;
#pragma unroll
  for (int r = 0; r < 7; r++)
    ls.w[r] = (uint32_t)MSG_SCHED[r][2 * q] | ((uint32_t)MSG_SCHED[r][2 * q + 1] << 8) | ((uint32_t)MSG_SCHED[r][8 + 2 * q] << 16) |
              ((uint32_t)MSG_SCHED[r][9 + 2 * q] << 24);
  return ls;
}

// Phase 1 for the compression circuit.  trace[TR_IN..TR_IN+28) must already hold h,m,t,b,d.
// All 32 lanes execute (8 redundant groups of 4); lanes 0..3 write.
__device__ __forceinline__ void compression_trace(uint32_t *trace, int lane, const lane_sched &ls) {
  const int q = lane & 3;
  const bool writer = lane < 4;
  const uint32_t IVq[4] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au};
  uint32_t a = trace[TR_IN + q];                 // v[q]     = h[q]
  uint32_t b = trace[TR_IN + 4 + q];             // v[4+q]   = h[4+q]
  uint32_t c = q == 0 ? IVq[0] : q == 1 ? IVq[1] : q == 2 ? IVq[2] : IVq[3];   // v[8+q] = IV[q]
  uint32_t d = trace[TR_IN + 24 + q];            // v[12+q]  = t0,t1,b,d               (:184-187)
  const uint32_t h_lo = a, h_hi = b;
  const uint32_t *m = trace + TR_IN + 8;
#pragma unroll
  for (int r = 0; r < 7; r++) {
    uint32_t *rec = trace + TR_HG + 128 * r + 16 * q;
    const uint32_t sw = ls.w[r];
    const uint32_t m0 = m[sw & 15u], m1 = m[(sw >> 8) & 15u], m2 = m[(sw >> 16) & 15u], m3 = m[sw >> 24];
    // columns: G(q, 4+q, 8+q, 12+q) with msg[2q], msg[2q+1]                              (:145-148)
    half_g<16, 12>(a, b, c, d, m0, rec, writer);
    half_g<8, 7>(a, b, c, d, m1, rec + 8, writer);
    // diagonals: lane q takes b from column q+1, c from q+2, d from q+3                  (:150-153)
    b = __shfl_sync(0xffffffffu, b, (q + 1) & 3, 4);
    c = __shfl_sync(0xffffffffu, c, (q + 2) & 3, 4);
    d = __shfl_sync(0xffffffffu, d, (q + 3) & 3, 4);
    half_g<16, 12>(a, b, c, d, m2, rec + 64, writer);
    half_g<8, 7>(a, b, c, d, m3, rec + 72, writer);
    b = __shfl_sync(0xffffffffu, b, (q + 3) & 3, 4);
    c = __shfl_sync(0xffffffffu, c, (q + 2) & 3, 4);
    d = __shfl_sync(0xffffffffu, d, (q + 1) & 3, 4);
  }
  if (writer) {                                  // out[i] = v[i]^v[i+8], out[i+8] = v[i+8]^h[i]  (:213-227)
    trace[TR_OUT + q] = a ^ c;
    trace[TR_OUT + 4 + q] = b ^ d;
    trace[TR_OUT + 8 + q] = c ^ h_lo;
    trace[TR_OUT + 12 + q] = d ^ h_hi;
  }
}

// Phase 1 for the nova step circuit Blake3Nova(0) (circuits/blake3_nova.circom:169-267, as built: without
// the Num2Bits(8) range checks of :25-30).  trace[NV_IN..NV_IN+32) holds the 32 inputs.  Computes every
// nova-level value (trace indices: nova_trace.h, generated from tools/circuit_model.py) and the inputs of
// the embedded compression (TR_IN..).  Returns false when a constraint fails ("Assert Failed.").
__device__ __forceinline__ bool nova_trace(uint32_t *trace, int lane) {
  const uint32_t *in = trace + NV_IN;
  const uint32_t n_blocks = in[0], block_count = in[1], low = in[10], high = in[11];
  const uint32_t leaf_depth = in[12], total_depth = in[13], depth = in[14], bb = in[31];
  // Blake3NovaTreePath_CheckDepth (:13-45)
  const int64_t v1 = (int64_t)depth + 256 - ((int64_t)leaf_depth - 1);      // check_parent = LessThan(8)(depth, leaf_depth-1)
  const int64_t v2 = (int64_t)leaf_depth + 256 - ((int64_t)depth + 1);      // exceed_depth = GreaterEqThan(8)(depth, leaf_depth)
  // Num2Bits(9) recomposition must hold for both, and exceed_depth.out === 0 (:44)
  if (v1 < 0 || v1 >= 512 || v2 < 0 || v2 >= 512 || ((v2 >> 8) & 1) == 0) return false;
  const uint32_t is_parent = 1u - (uint32_t)((v1 >> 8) & 1);
  const uint32_t is_root = depth == 0;
  // Blake3GetFlag (:122-167)
  const uint32_t not_root = 1u - is_root, not_parent = 1u - is_parent;
  const uint32_t first = block_count == 0;
  const uint32_t last = (int64_t)block_count == (int64_t)n_blocks - 1;
  const uint32_t is_last = last & not_parent, first_set = first & not_parent;
  const uint32_t urf_tmp = is_parent | last, urf = urf_tmp & is_root;
  const uint32_t dflags = first_set + 2u * is_last + 8u * urf + 4u * is_parent;
  // Blake3GetDownLeftPath (:47-84): eqs[i] = IsEqual(depth, total_depth - i - 2), i = lane and lane + 32
  const int64_t in1a = (int64_t)total_depth - lane - 2, in1b = in1a - 32;
  const int64_t da = in1a - depth, db = in1b - depth;
  uint2 *eq_in1 = reinterpret_cast<uint2 *>(trace + NV_EQ_IN1), *eq_d = reinterpret_cast<uint2 *>(trace + NV_EQ_D);
  eq_in1[lane] = make_uint2((uint32_t)in1a, (uint32_t)((uint64_t)in1a >> 32));
  eq_in1[lane + 32] = make_uint2((uint32_t)in1b, (uint32_t)((uint64_t)in1b >> 32));
  eq_d[lane] = make_uint2((uint32_t)da, (uint32_t)((uint64_t)da >> 32));
  eq_d[lane + 32] = make_uint2((uint32_t)db, (uint32_t)((uint64_t)db >> 32));
  const uint32_t eq_lo = __ballot_sync(0xffffffffu, da == 0), eq_hi = __ballot_sync(0xffffffffu, db == 0);
  // bit_at_depth[i] = sum_{j<=i} (1 - n2b.out[j]) * eqs[j].out: at most one term is non-zero  (:65,:70)
  const uint64_t mask = (((uint64_t)eq_hi << 32) | eq_lo) & ~(((uint64_t)high << 32) | low);
  const uint64_t bad = mask ? ~((mask & (0 - mask)) - 1) : 0;
  const uint32_t dlp = not_parent + is_parent * (uint32_t)(bad >> 63);      // out (:79); boolean by construction (:81)
  // Blake3GetFinal_m (:86-120)
  if (lane < 16) {
    const uint32_t hw = in[2 + (lane & 7)], mw = in[15 + lane], mo = in[15 + (lane & 7)];
    const uint32_t td = lane < 8 ? hw * dlp : hw * (1u - dlp);
    const uint32_t mp = (lane < 8 ? mw * (1u - dlp) : mo * dlp) + td;
    const uint32_t tp = mp * is_parent;
    trace[NV_TMP_DOWN + lane] = td;
    trace[NV_M_IS_PAR + lane] = mp;
    trace[NV_TMP_IS_PAR + lane] = tp;
    trace[TR_IN + 8 + lane] = mw * not_parent + tp;                          // out_m -> compression m
  }
  if (lane < 8) {                                                           // :229-233
    const uint32_t IVc[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    uint32_t ivw = IVc[0];
#pragma unroll
    for (int j = 1; j < 8; j++) ivw = lane == j ? IVc[j] : ivw;
    const uint32_t tiv = ivw * is_parent;
    trace[NV_TMPIV + lane] = tiv;
    trace[TR_IN + lane] = in[2 + lane] * not_parent + tiv;                   // h_compression
  }
  if (lane == 0) {
    const uint32_t cdd = is_last | is_parent, decr = cdd & not_root;        // :254-258
    const int64_t neg_depth = -(int64_t)depth, neg_bc = -(int64_t)block_count;
    const int64_t nbm1 = (int64_t)n_blocks - 1, bc_diff = nbm1 - block_count;
    const uint64_t bc_out = (uint64_t)block_count + not_parent;             // :251
    trace[NV_V1] = (uint32_t)v1; trace[NV_V2] = (uint32_t)v2;
    trace[NV_LDM1] = leaf_depth - 1u; trace[NV_DP1] = depth + 1u;
    trace[NV_IS_PARENT] = is_parent; trace[NV_EXCEED] = 0u; trace[NV_IS_ROOT] = is_root;
    trace[NV_NOT_ROOT] = not_root; trace[NV_NOT_PARENT] = not_parent;
    trace[NV_BC_FIRST] = first; trace[NV_BC_LAST] = last; trace[NV_IS_LAST] = is_last; trace[NV_FIRST_SET] = first_set;
    trace[NV_URF_TMP] = urf_tmp; trace[NV_URF] = urf; trace[NV_DLP] = dlp;
    trace[NV_CDD] = cdd; trace[NV_DECR] = decr; trace[NV_DEPTH_OUT] = depth - decr;   // :262
    trace[NV_NEG_DEPTH] = (uint32_t)neg_depth; trace[NV_NEG_DEPTH + 1] = (uint32_t)((uint64_t)neg_depth >> 32);
    trace[NV_NEG_BC] = (uint32_t)neg_bc; trace[NV_NEG_BC + 1] = (uint32_t)((uint64_t)neg_bc >> 32);
    trace[NV_NBM1] = (uint32_t)nbm1; trace[NV_NBM1 + 1] = (uint32_t)((uint64_t)nbm1 >> 32);
    trace[NV_BC_DIFF] = (uint32_t)bc_diff; trace[NV_BC_DIFF + 1] = (uint32_t)((uint64_t)bc_diff >> 32);
    trace[NV_BC_OUT] = (uint32_t)bc_out; trace[NV_BC_OUT + 1] = (uint32_t)(bc_out >> 32);
    trace[NV_EQ_OUT] = eq_lo; trace[NV_EQ_OUT + 1] = eq_hi;
    trace[NV_BAD] = (uint32_t)bad; trace[NV_BAD + 1] = (uint32_t)(bad >> 32);
    trace[TR_IN + 24] = low * not_parent;                                   // t[0] (:245)
    trace[TR_IN + 25] = high * not_parent;                                  // t[1] (:244)
    trace[TR_IN + 26] = bb;
    trace[TR_IN + 27] = dflags;                                             // comp_d.out (:161-165)
  }
  return true;
}

// Slow path of phase 2 (nova only): a slot that holds a true field element.  Kept out of line so that the hot
// loop keeps its small register footprint.
__device__ __forceinline__ fr_t store_field_slot(uint8_t *p, uint32_t kind, uint32_t lo, uint32_t hi,
                                              const field_consts *__restrict__ F) {
  const int64_t x = (int64_t)(((uint64_t)hi << 32) | lo);
  fr_t v;
  if (kind == DK_S64) v = fr_from_s64(x, F->p);
  else v = fr_inv_s64(x, *F);
  st_slot_fr(p, v.l);
  return v;
}

// ---- per-instance witness checksum, computed from the values on their way to HBM (b3w_batch_extras.sums) -----------------
// b3w_checksum_device's definition: sum over slots s and 64-bit limbs j of (limb[s][j] + 1) * mix(4 s + j) mod 2^64 with
// mix(x) = (x + 1) * K.  K factors out, so a lane accumulates sum (limb + 1) * (4 s + j + 1) and the warp multiplies once.
// A hot-path slot is {v, 0, 0, 0} (v = lo | hi << 32): (v + 1)(4 s + 1) + (4 s + 2) + (4 s + 3) + (4 s + 4).
#define B3W_SUM_K 0x9E3779B97F4A7C15ull
__device__ __forceinline__ uint64_t sum_small_slot(uint32_t s, uint32_t lo, uint32_t hi) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (v + 1) * (4ull * s + 1) + 12ull * s + 9;
}
__device__ __forceinline__ uint64_t sum_field_slot(uint32_t s, const fr_t &v) {
  uint64_t acc = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) acc += ((((uint64_t)v.l[2 * j + 1] << 32) | v.l[2 * j]) + 1) * (4ull * s + j + 1);
  return acc;
}
// adds the warp's partial sum to sums[i] (zeroed before the launch; the parts of one witness are expanded by different warps)
__device__ __forceinline__ void sum_commit(unsigned long long *sums, uint64_t i, uint64_t acc, int lane) {
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) atomicAdd(sums + i, (unsigned long long)(acc * B3W_SUM_K));
}

// Phase 2: expand the trace into witness slots [s0, s1) at `dst` (32 B per slot).
// kinds BIT / W32 / W64 are the hot path (single 256-bit store, upper 6 words from RZ).  Nova's true field elements
// (kinds S64 / INV; 67 .. 260 slots per witness) are skipped here and written by a second pass over the list of field
// slots (`fslots`: {slot, descriptor} pairs), in which all 32 lanes do field arithmetic together instead of one lane
// diverging inside the hot loop.
// SUMS: also returns this lane's share of the checksum of what it stored (see sum_small_slot), else 0.
template <bool HAS_FIELD, bool SUMS = false>
__device__ __forceinline__ uint64_t expand_slots(const uint32_t *trace, const uint32_t *__restrict__ desc, uint32_t s0, uint32_t s1,
                                                 uint8_t *dst, int lane, const field_consts *__restrict__ F,
                                                 const uint2 *__restrict__ fslots, uint32_t n_fslots) {
  uint64_t acc = 0;
#pragma unroll 4
  for (uint32_t s = s0 + lane; s < s1; s += 32) {
    const uint32_t dsc = __ldg(desc + s);
    const uint32_t t = dsc & 0xFFFFu, k = (dsc >> 16) & 31u, kind = dsc >> 24;
    const uint32_t w = trace[t];
    uint32_t lo = kind == DK_BIT ? ((w >> k) & 1u) : w;
    uint32_t hi = kind == DK_W64 ? trace[t + 1] : 0u;
    if (!HAS_FIELD || kind < DK_S64) {
      st_slot(dst + (size_t)s * 32, lo, hi, 0u, 0u, 0u, 0u, 0u, 0u);
      if (SUMS) acc += sum_small_slot(s, lo, hi);
    }
  }
  if (HAS_FIELD) {
    for (uint32_t j = lane; j < n_fslots; j += 32) {
      const uint2 fs = __ldg(fslots + j);
      const uint32_t t = fs.y & 0xFFFFu;
      if (fs.x >= s0 && fs.x < s1) {
        const fr_t v = store_field_slot(dst + (size_t)fs.x * 32, fs.y >> 24, trace[t], trace[t + 1], F);
        if (SUMS) acc += sum_field_slot(fs.x, v);
      }
    }
  }
  return acc;
}

#define WARPS_PER_CTA 8
#define TRACE_STRIDE 960          // u32 words per warp, compression (>= 944, 16-byte multiple)
#define NOVA_TRACE_STRIDE 1344    // u32 words per warp, nova (>= NOVA_TRACE_WORDS)
static_assert(NOVA_TRACE_WORDS <= NOVA_TRACE_STRIDE, "nova trace does not fit its stride");
#define NOVA_SMEM(warps) ((warps) * NOVA_TRACE_STRIDE * 4)

// Optional extras of the checked kernel variants: the fused R1CS check (rows evaluated on the shared-memory trace,
// nothing re-read from HBM) and a fault-injection hook for its negative tests.
struct check_args {
  r1cs_tables_dev T;
  const field_consts *F;
  uint32_t *first_bad;       // per instance: smallest violated row id or B3W_NO_ROW (may be NULL)
  uint32_t fault_word;       // trace word to corrupt (B3W_NO_ROW = none) ...
  uint32_t fault_mask;       // ... by xor with this mask, after the trace phase
  unsigned long long *sums;  // SUMS kernel variants: per-instance witness checksum, zeroed before the launch (else unused)
};

// Work distribution.  A work item is one PART of one instance: slots [part * part_len, (part + 1) * part_len) of its
// witness.  Warps of the persistent grid take items from a global counter (dynamic scheduling): SMs do not all see the
// same HBM bandwidth, and with a static split the launch ends with a long tail of slow warps; measured on B200
// (2^16 compression instances) 5.9 TB/s static vs 7.2 TB/s dynamic.  A warp computes the (cheap) trace of the
// item's instance and expands only the item's slots; part 0 also writes status / public outputs / the check result.
#define SCHED_LANES 8             // sub-counters per launch: same-address atomics serialise in one L2 slice (~2.4 ns each)
#define SCHED_STRIDE 16           // u64 between sub-counters (128 B: one L2 line each)
struct sched_args {
  unsigned long long *counter;     // SCHED_LANES sub-counters, zeroed before the launch; sub-counter c hands out the
                                   // items {v * SCHED_LANES + c}
  unsigned int parts;              // items per instance
  unsigned int part_len;           // slots per item, a multiple of 32
};

// Software pipeline over work items: while item k is traced and expanded, the input row of item k+1 is already on its
// way from HBM and the counter grab for item k+2 is in flight, so neither latency sits between two expansions.
template <int N_IN>
struct item_pipe {
  const sched_args &sc;
  const uint32_t *__restrict__ in;
  uint64_t n, total;
  int lane;
  uint32_t sub, tries;                    // current sub-counter, exhausted sub-counters seen so far
  unsigned long long cur, nxt, grabbed;   // item ids: being processed / input row loading / grab in flight (lane 0)
  uint32_t cur_in, nxt_in;                // this lane's word of the input rows

  __device__ __forceinline__ unsigned long long grab() {
    return lane == 0 ? atomicAdd(sc.counter + sub * SCHED_STRIDE, 1ull) * SCHED_LANES + sub : 0ull;
  }
  // the grabbed id, or -- when its sub-counter has run dry -- an id from the next sub-counter that still has work
  // (a dry result that was grabbed before the last switch says nothing about the current sub-counter)
  __device__ __forceinline__ unsigned long long resolve(unsigned long long g) {
    unsigned long long id = __shfl_sync(0xffffffffu, g, 0);
    while (id >= total && tries < SCHED_LANES) {
      if (id % SCHED_LANES == sub) {
        tries++;
        sub = (sub + 1) % SCHED_LANES;
      }
      id = __shfl_sync(0xffffffffu, grab(), 0);
    }
    return id;
  }
  __device__ __forceinline__ uint32_t load_row(unsigned long long item) {
    return (item < total && lane < N_IN) ? __ldg(in + (item / sc.parts) * N_IN + lane) : 0u;
  }
  __device__ __forceinline__ item_pipe(const sched_args &sc_, const uint32_t *in_, uint64_t n_, int lane_, uint64_t gwarp)
      : sc(sc_), in(in_), n(n_), total(n_ * sc_.parts), lane(lane_), sub((uint32_t)(gwarp % SCHED_LANES)), tries(0) {
    cur = resolve(grab());
    nxt = resolve(grab());
    cur_in = load_row(cur);
    nxt_in = load_row(nxt);
    grabbed = grab();
  }
  __device__ __forceinline__ bool valid() const { return cur < total; }
  __device__ __forceinline__ uint64_t inst() const { return cur / sc.parts; }
  __device__ __forceinline__ uint32_t part() const { return (uint32_t)(cur % sc.parts); }
  // call once the current item's input word has been consumed: shifts the pipeline and refills its far end
  __device__ __forceinline__ void advance() {
    cur = nxt;
    cur_in = nxt_in;
    nxt = resolve(grabbed);
    nxt_in = load_row(nxt);
    grabbed = grab();
  }
};

// The *_checked variants add CHECK_WARPS "checker" warps to every CTA.  The 8 expansion warps run exactly the loop of the
// plain kernel (so the store stream keeps the shape that reaches the write roofline); the checker warps take whole
// instances from a second set of counters, recompute the trace and evaluate the R1CS rows on it, filling issue slots the
// store-bound expansion leaves idle.  A warp whose own queue has run dry helps with the other queue (phase 1), so the
// launch has no tail of one kind of work.  With the check inside the expansion warps (4 items per witness, every
// resident CTA) the fused kernels ran at 6.3 (compression) / 4.95 TB/s (nova); see profiles/.
#ifdef B3W_EXP_NOCHECK              /* experiment builds only: checker warps trace but do not evaluate rows */
#define B3W_EXP_CHECK(x) B3W_NO_ROW
#else
#define B3W_EXP_CHECK(x) (x)
#endif
#ifndef CHECK_WARPS
#define CHECK_WARPS 4          /* compression */
#endif
#ifndef NOVA_CHECK_WARPS
#define NOVA_CHECK_WARPS 4     /* nova: more rows per instance (1 124 vs 688) */
#endif

// ---- the FULL input domain of blake3_compression (WIDE kernel variants) -------------------------------------------
// The circuit range-checks every input except the message words (circuits/blake3_compression.circom:169-170 TODO): m[j]
// only ever enters a sum add1 = Bits34(v[a] + v[b] + m[j]), whose recomposition constraint holds whenever that sum, as a
// canonical field element, is below 2^34 (circuits/blake3_common.circom:182-203) -- e.g. m[0] = 2^32 or m[0] = p - 1 give
// valid witnesses in the reference.  A message word of a satisfying input is therefore a signed
// integer m = ext * 2^32 + lo with ext in [-2, 3], and such an instance differs from the u32 instance with the same lo
// words in exactly two places: the carry word of every half-G record that adds m[j] grows by ext[j] (and must stay in
// [0, 3], else "Assert Failed."), and the witness slots of m itself hold m mod p.  So the wide variants run the ordinary
// u32 trace and then apply that correction; the hot (u32-only) instantiations do not contain any of it.
#define TR_EXT 944u               /* 16 sign-extended ext words; free in the compression kernels' 960-word trace stride */
/* B3W_EXT_ASSERT (include/blake3wit.h) in m_ext[i][0]: the host already knows that instance i asserts; no carry survives + 127 */
static_assert(TR_EXT >= B3W_TRACE_WORDS_COMPRESSION && TR_EXT + 16 <= TRACE_STRIDE, "ext words do not fit the compression trace stride");
struct wide_args {
  const int8_t *m_ext;            // n x 16, or NULL (then the kernel must be a non-WIDE instantiation)
  const field_consts *F;
  uint32_t m_slot0;               // witness slot of m[0]; m[j] is slot m_slot0 + j
};

// Adds ext to the carry words of the 112 half-G records (APPLY) or only looks at them; true = some sum left [0, 2^34).
// ext: 16 sign-extended words (the compression kernels keep them at trace + TR_EXT).
template <bool APPLY>
__device__ __forceinline__ bool wide_carries_at(uint32_t *trace, const uint32_t *ext, int lane) {
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int rec = lane + 32 * k;                 // record ((round * 8 + g) * 2 + half) adds msg[2g + half] of that round
    if (rec < 112) {
      const int e = (int)ext[MSG_SCHED[rec >> 4][rec & 15]];
      const int c = (int)trace[TR_HG + 8 * rec + 1] + e;
      bad = bad || c < 0 || c > 3;
      if (APPLY) trace[TR_HG + 8 * rec + 1] = (uint32_t)c;
    }
  }
  return __any_sync(0xffffffffu, bad);
}
template <bool APPLY>
__device__ __forceinline__ bool wide_carries(uint32_t *trace, int lane) { return wide_carries_at<APPLY>(trace, trace + TR_EXT, lane); }

// The m slots of a wide instance inside [a, b): m mod p.  Each is rewritten by the lane that expand_slots used for it
// (a is a multiple of 32), so the two stores to one address are ordered by program order.
// Returns the change of this lane's checksum share (new slot content minus what expand_slots accounted for).
__device__ __forceinline__ uint64_t wide_patch_m(const uint32_t *trace, const wide_args &wd, uint32_t a, uint32_t b, uint8_t *dst, int lane) {
  const uint32_t j = ((uint32_t)lane - wd.m_slot0) & 31u, s = wd.m_slot0 + j;
  if (j < 16u && s >= a && s < b) {
    const int e = (int)trace[TR_EXT + j];
    const uint32_t lo = trace[TR_IN + 8 + j];
    if (e > 0) {
      st_slot(dst + (size_t)s * 32, lo, (uint32_t)e, 0u, 0u, 0u, 0u, 0u, 0u);
      return sum_small_slot(s, lo, (uint32_t)e) - sum_small_slot(s, lo, 0u);
    } else if (e < 0) {
      const fr_t v = fr_from_s64((int64_t)(((uint64_t)(uint32_t)e << 32) | lo), wd.F->p);
      st_slot_fr(dst + (size_t)s * 32, v.l);
      return sum_field_slot(s, v) - sum_small_slot(s, lo, 0u);
    }
  }
  return 0;
}

// k_blake3_comp_witness: compression circuit, one warp per work item (see above).
template <bool CHECK, bool WIDE, bool SUMS>
__global__ void __launch_bounds__((WARPS_PER_CTA + (CHECK ? CHECK_WARPS : 0)) * 32, CHECK ? 2 : 4)
k_blake3_comp_witness(const uint32_t *__restrict__ in, uint64_t n, const uint32_t *__restrict__ desc, uint32_t ws,
                      uint8_t *__restrict__ out, uint8_t *__restrict__ status, uint32_t *__restrict__ pub,
                      const check_args ck, const sched_args sc, const sched_args sck, const wide_args wd) {
  constexpr int WARPS = WARPS_PER_CTA + (CHECK ? CHECK_WARPS : 0);
  __shared__ __align__(16) uint32_t s_trace[WARPS][TRACE_STRIDE];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t *trace = s_trace[wib];
  if (lane == 0) { trace[TR_ZERO] = 0u; trace[TR_ONE] = 1u; }
  const lane_sched ls = load_lane_sched(lane);
  const bool checker = CHECK && wib >= WARPS_PER_CTA;
  const bool fault = CHECK && ck.fault_word != B3W_NO_ROW;
#pragma unroll 1
  for (int phase = 0; phase < (CHECK ? 2 : 1); phase++) {
    if (CHECK && checker == (phase == 0)) {
      // ---- check items: one instance each ----
      for (item_pipe<28> pipe(sck, in, n, lane, (uint64_t)blockIdx.x * WARPS + wib); pipe.valid();) {
        const uint64_t i = pipe.inst();
        __syncwarp();
        if (lane < 28) trace[TR_IN + lane] = pipe.cur_in;
        if (WIDE && lane < 16) trace[TR_EXT + lane] = (uint32_t)(int)wd.m_ext[i * 16 + lane];
        pipe.advance();
        __syncwarp();
        compression_trace(trace, lane, ls);
        __syncwarp();
        if (fault && lane == 0) trace[ck.fault_word] ^= ck.fault_mask;
        __syncwarp();
        // WIDE: the rows are evaluated on the u32 instance with the same low words (they hold for one iff for the other:
        // the ext terms cancel), the carry range decides "Assert Failed."
        const uint32_t bad = B3W_EXP_CHECK(r1cs_check_instance(TraceSrc{trace, ck.F}, ck.T, lane));
        const bool asserted = WIDE && wide_carries<false>(trace, lane);
        if (ck.first_bad && lane == 0) ck.first_bad[i] = asserted ? B3W_NO_ROW : bad;
        if (status && lane == 0) status[i] = asserted ? B3W_CIRCOM_ASSERT : bad != B3W_NO_ROW ? B3W_R1CS_VIOLATION : 0;
      }
    } else {
      // ---- expansion items: 1/parts of one witness each ----
      for (item_pipe<28> pipe(sc, in, n, lane, (uint64_t)blockIdx.x * WARPS + wib); pipe.valid();) {
        const uint64_t i = pipe.inst();
        const uint32_t part = pipe.part();
        const uint32_t a = part * sc.part_len, b = a + sc.part_len < ws ? a + sc.part_len : ws;
        __syncwarp();                               // the previous expansion has finished reading the trace
        if (lane < 28) trace[TR_IN + lane] = pipe.cur_in;
        if (WIDE && lane < 16) trace[TR_EXT + lane] = (uint32_t)(int)wd.m_ext[i * 16 + lane];
        pipe.advance();
        __syncwarp();
        compression_trace(trace, lane, ls);
        __syncwarp();
        if (WIDE) {
          const bool asserted = wide_carries<true>(trace, lane);
          __syncwarp();
          if (asserted) {                           // the reference throws "Assert Failed.": no witness exists
            if (part == 0) {
              if (!CHECK && status && lane == 0) status[i] = B3W_CIRCOM_ASSERT;
              if (pub && lane < 16) pub[i * 16 + lane] = 0u;
            }
            continue;
          }
        }
        if (part == 0) {                            // this warp owns the instance's head
          if (pub && lane < 16) pub[i * 16 + lane] = trace[TR_OUT + lane];
          // u32 inputs can never violate a constraint of this circuit (which the fused check confirms row by row)
          if (!CHECK && status && lane == 0) status[i] = 0;
        }
        if (fault) {                                // keep the injected fault visible in the witness
          if (lane == 0) trace[ck.fault_word] ^= ck.fault_mask;
          __syncwarp();
        }
        uint64_t acc = expand_slots<false, SUMS>(trace, desc, a, b, out + i * (uint64_t)ws * 32, lane, nullptr, nullptr, 0);
        if (WIDE) acc += wide_patch_m(trace, wd, a, b, out + i * (uint64_t)ws * 32, lane);
        if (SUMS) sum_commit(ck.sums, i, acc, lane);
      }
    }
  }
}

// ---- experiment (b3w_debug_set_store_mode(ctx, 1)): expanded tiles staged in shared memory, written by the TMA engine -------
// What BASELINE's north_star sketches ("stage each instance's witness in shared memory and write it back with ... TMA bulk
// stores").  Each warp expands 128 slots (4 KiB) at a time into one of TMA_NBUF private shared-memory tiles and one elected
// lane hands the tile to the TMA engine (cp.async.bulk.global.shared::cta, SASS UBLKCP); the next tile is expanded while
// that copy drains.  The two 16-byte halves of a slot are written in an order that keeps the shared-memory stores free of
// bank conflicts (lanes 0-3 / 4-7 of every quarter-warp start with opposite halves).  blake3_compression, plain u32 inputs
// only: a nova witness has field slots that a second generic-proxy store would race the async-proxy copy for.
// Measured against the direct 256-bit stores in profiles/r02_store_mode.jsonl; the default stays whichever is faster there.
#define TMA_TILE_SLOTS 128u
#define TMA_TILE_BYTES (TMA_TILE_SLOTS * 32u)
#define TMA_NBUF 2
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 2)
k_blake3_comp_witness_tma(const uint32_t *__restrict__ in, uint64_t n, const uint32_t *__restrict__ desc, uint32_t ws,
                          uint8_t *__restrict__ out, uint8_t *__restrict__ status, uint32_t *__restrict__ pub, const sched_args sc) {
  __shared__ __align__(16) uint32_t s_trace[WARPS_PER_CTA][TRACE_STRIDE];
  extern __shared__ __align__(128) uint8_t s_tiles[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t *trace = s_trace[wib];
  uint8_t *tiles = s_tiles + (size_t)wib * TMA_NBUF * TMA_TILE_BYTES;
  if (lane == 0) { trace[TR_ZERO] = 0u; trace[TR_ONE] = 1u; }
  const lane_sched ls = load_lane_sched(lane);
  const bool flip = (lane >> 2) & 1;               // which half of its slot this lane writes first
  uint32_t buf = 0;
  for (item_pipe<28> pipe(sc, in, n, lane, (uint64_t)blockIdx.x * WARPS_PER_CTA + wib); pipe.valid();) {
    const uint64_t i = pipe.inst();
    const uint32_t part = pipe.part();
    const uint32_t a = part * sc.part_len, b = a + sc.part_len < ws ? a + sc.part_len : ws;
    __syncwarp();
    if (lane < 28) trace[TR_IN + lane] = pipe.cur_in;
    pipe.advance();
    __syncwarp();
    compression_trace(trace, lane, ls);
    __syncwarp();
    if (part == 0) {
      if (pub && lane < 16) pub[i * 16 + lane] = trace[TR_OUT + lane];
      if (status && lane == 0) status[i] = 0;
    }
    uint8_t *dst = out + i * (uint64_t)ws * 32;
    for (uint32_t s0 = a; s0 < b; s0 += TMA_TILE_SLOTS) {
      const uint32_t cnt = min(TMA_TILE_SLOTS, b - s0);
      // the copy that read this buffer TMA_NBUF tiles ago has finished reading it
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(TMA_NBUF - 1) : "memory");
      __syncwarp();
      uint8_t *tile = tiles + (size_t)buf * TMA_TILE_BYTES;
#pragma unroll 4
      for (uint32_t s = s0 + lane; s < s0 + cnt; s += 32) {
        const uint32_t dsc = __ldg(desc + s);
        const uint32_t t = dsc & 0xFFFFu, k = (dsc >> 16) & 31u, kind = dsc >> 24;
        const uint32_t w = trace[t];
        const uint32_t lo = kind == DK_BIT ? ((w >> k) & 1u) : w;
        const uint32_t hi = kind == DK_W64 ? trace[t + 1] : 0u;
        uint4 *p = reinterpret_cast<uint4 *>(tile + (size_t)(s - s0) * 32);
        const uint4 h0 = make_uint4(lo, hi, 0u, 0u), h1 = make_uint4(0u, 0u, 0u, 0u);
        p[flip ? 1 : 0] = flip ? h1 : h0;
        p[flip ? 0 : 1] = flip ? h0 : h1;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // this thread's tile writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)s0 * 32),
                     "r"((uint32_t)__cvta_generic_to_shared(tile)), "r"(cnt * 32u)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      buf = (buf + 1) % TMA_NBUF;
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the tiles stay valid until the engine is done with them
}

// The 15 outputs z_{i+1} of a nova step, one per lane < 15: n_blocks_out, block_count_out, h_out[8], total_depth_out,
// depth_out, chunk_idx_low/high_out, leaf_depth_out (circuits/blake3_nova.circom:195-202).
__device__ __forceinline__ uint32_t nova_public_output(const uint32_t *trace, int lane) {
  if (lane == 0) return trace[NV_IN + 0];
  if (lane == 1) return trace[NV_BC_OUT];
  if (lane < 10) return trace[TR_OUT + lane - 2];
  if (lane == 10) return trace[NV_IN + 13];
  if (lane == 11) return trace[NV_DEPTH_OUT];
  if (lane == 12) return trace[NV_IN + 10];
  if (lane == 13) return trace[NV_IN + 11];
  return trace[NV_IN + 12];
}

// k_blake3_nova_witness: the nova step circuit (all three committed builds share it; they differ in the
// slot table and the prime).  pub = z_{i+1} = the 15 outputs (low 32 bits each).
template <bool CHECK, bool SUMS>
__global__ void __launch_bounds__((WARPS_PER_CTA + (CHECK ? NOVA_CHECK_WARPS : 0)) * 32, CHECK ? 2 : 3)
k_blake3_nova_witness(const uint32_t *__restrict__ in, uint64_t n, const uint32_t *__restrict__ desc, uint32_t ws,
                      const field_consts *__restrict__ F, const uint2 *__restrict__ fslots, uint32_t n_fslots,
                      uint8_t *__restrict__ out, uint8_t *__restrict__ status, uint32_t *__restrict__ pub,
                      const check_args ck, const sched_args sc, const sched_args sck) {
  constexpr int WARPS = WARPS_PER_CTA + (CHECK ? NOVA_CHECK_WARPS : 0);
  extern __shared__ __align__(16) uint32_t s_dyn[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t *trace = s_dyn + wib * NOVA_TRACE_STRIDE;
  if (lane == 0) { trace[TR_ZERO] = 0u; trace[TR_ONE] = 1u; }
  const lane_sched ls = load_lane_sched(lane);
  const bool checker = CHECK && wib >= WARPS_PER_CTA;
  const bool fault = CHECK && ck.fault_word != B3W_NO_ROW;
#pragma unroll 1
  for (int phase = 0; phase < (CHECK ? 2 : 1); phase++) {
    if (CHECK && checker == (phase == 0)) {
      for (item_pipe<32> pipe(sck, in, n, lane, (uint64_t)blockIdx.x * WARPS + wib); pipe.valid();) {
        const uint64_t i = pipe.inst();
        __syncwarp();
        trace[NV_IN + lane] = pipe.cur_in;
        pipe.advance();
        __syncwarp();
        if (!nova_trace(trace, lane)) {           // the reference throws "Assert Failed.": no witness exists
          if (status && lane == 0) status[i] = B3W_CIRCOM_ASSERT;
          if (ck.first_bad && lane == 0) ck.first_bad[i] = B3W_NO_ROW;
          continue;
        }
        __syncwarp();
        compression_trace(trace, lane, ls);
        __syncwarp();
        if (fault && lane == 0) trace[ck.fault_word] ^= ck.fault_mask;
        __syncwarp();
        const uint32_t bad = B3W_EXP_CHECK(r1cs_check_instance(TraceSrc{trace, ck.F}, ck.T, lane));
        if (ck.first_bad && lane == 0) ck.first_bad[i] = bad;
        if (status && lane == 0) status[i] = bad != B3W_NO_ROW ? B3W_R1CS_VIOLATION : 0;
      }
    } else {
      for (item_pipe<32> pipe(sc, in, n, lane, (uint64_t)blockIdx.x * WARPS + wib); pipe.valid();) {
        const uint64_t i = pipe.inst();
        const uint32_t part = pipe.part();
        const uint32_t a = part * sc.part_len, b = a + sc.part_len < ws ? a + sc.part_len : ws;
        const bool head = part == 0;
        __syncwarp();
        trace[NV_IN + lane] = pipe.cur_in;
        pipe.advance();
        __syncwarp();
        if (!nova_trace(trace, lane)) {
          if (head) {
            if (!CHECK && status && lane == 0) status[i] = B3W_CIRCOM_ASSERT;
            if (pub && lane < 15) pub[i * 15 + lane] = 0u;
          }
          continue;
        }
        __syncwarp();
        compression_trace(trace, lane, ls);
        __syncwarp();
        if (head) {
          if (!CHECK && status && lane == 0) status[i] = 0;
          if (pub && lane < 15) pub[i * 15 + lane] = nova_public_output(trace, lane);
        }
        if (fault) {
          if (lane == 0) trace[ck.fault_word] ^= ck.fault_mask;
          __syncwarp();
        }
        const uint64_t acc = expand_slots<true, SUMS>(trace, desc, a, b, out + i * (uint64_t)ws * 32, lane, F, fslots, n_fslots);
        if (SUMS) sum_commit(ck.sums, i, acc, lane);
      }
    }
  }
}

