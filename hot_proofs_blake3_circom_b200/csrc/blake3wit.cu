// blake3wit.cu -- C ABI of libblake3wit.so (see include/blake3wit.h) and the host side around the sm_100a kernels, which
// live in kernels_witness.cuh (trace + expansion + fused check), kernels_chain.cuh (BLAKE3 tree, step rows) and
// kernels_aux.cuh (compact witnesses, HBM check, checksum, calibration).
//
// Replaces the reference's wasm witness programs (build/**/**.wasm driven by
// blake3_nova_js/witness_calculator.js:131-272).  Per instance:
//   phase 1 (trace):  native u32 BLAKE3 compression, 4 lanes = 4 G functions in parallel, every
//                     intermediate the circuit exposes is written to a ~4 KB shared-memory trace
//                     (layout: trace_layout.h; semantics: circuits/blake3_compression.circom:72-228);
//   phase 2 (expand): each lane turns one slot descriptor into one canonical 32-byte field element and
//                     writes it with a single 256-bit streaming store (STG.E.256), so one warp
//                     instruction covers 1 KiB of contiguous .wtns body.
// The kernel is HBM-write bound: 770 976 B written per compression witness vs 112 B read.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <new>
#include <exception>
#include <vector>
#include <string>
#include <algorithm>
#include <thread>
#include <mutex>
#include <chrono>
#include <ctype.h>
#include <emmintrin.h>
#include <nvtx3/nvToolsExt.h>
#ifdef __linux__
#include <unistd.h>
#include <sys/syscall.h>
#endif

#include "../../include/blake3wit.h"
#include "trace_layout.h"
#include "slot_tables.h"
#include "nova_trace.h"
#include "fr.cuh"
#include "r1cs_tables.h"
#include "r1cs.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512];
static int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return fail(B3W_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// The C ABI never lets a C++ exception cross it ("never abort"): entry points that allocate through the standard library
// or start threads run their bodies under this guard.
template <class F>
static int guarded(const char *what, F &&body) {
  try {
    return body();
  } catch (const std::bad_alloc &) {
    return fail(B3W_ERR_NOMEM, "%s: out of host memory", what);
  } catch (const std::exception &e) {
    return fail(B3W_ERR_INVALID, "%s: %s", what, e.what());
  } catch (...) {
    return fail(B3W_ERR_INVALID, "%s: unexpected exception", what);
  }
}

#include "kernels_witness.cuh"
#include "kernels_nova_wide.cuh"
#include "kernels_chain.cuh"
#include "kernels_aux.cuh"
#include "r1cs_rows.cuh"
#include "kernels_r1cs_fast.cuh"
#include "kernels_fr_input.cuh"
#include "r1cs_load.h"
#include "wide_domain.h"
#include "device_mem.h"

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct circuit_def {
  const char *name;
  bool nova;
  uint32_t ws, n_inputs, n_public, trace_words;
  const b3w_seg *segs;
  size_t n_segs;
  const uint8_t *prime;
  struct r1cs_set {
    const b3w_r1cs_class *cls; size_t ncls;
    const uint64_t (*coef)[2]; size_t ncoef;
    const b3w_seg *cols; size_t ncols;
    uint32_t rows, terms;
  } r_fused, r_slots;    // reduced trace-space set (fused check) / all rows in witness-slot space (O1 builds: template level; O2 builds: the O2-form system)
  int n_sig;
  struct { const char *name; uint32_t off, size; } sig[12];
};

static const uint8_t PRIME_BN254[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9,
                                        0x79, 0x48, 0xe8, 0x33, 0x28, 0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45,
                                        0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};

static const uint8_t PRIME_PALLAS_SCALAR[32] = {0x01, 0x00, 0x00, 0x00, 0x21, 0xeb, 0x46, 0x8c, 0xdd, 0xa8, 0x94, 0x09, 0xfc, 0x98, 0x46, 0x22, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x40};

#define NOVA_SIGS                                                                                              \
  10, {{"n_blocks", 0, 1}, {"block_count", 1, 1}, {"h", 2, 8}, {"chunk_idx_low", 10, 1}, {"chunk_idx_high", 11, 1}, \
       {"leaf_depth", 12, 1}, {"total_depth", 13, 1}, {"depth", 14, 1}, {"m", 15, 16}, {"b", 31, 1}}
#define SEGS(v) B3W_SEGS_##v, sizeof(B3W_SEGS_##v) / sizeof(b3w_seg)
#define R1CS1(v)                                                                                                  \
  {R1CS_CLASSES_##v, sizeof(R1CS_CLASSES_##v) / sizeof(b3w_r1cs_class), R1CS_COEFS_##v, sizeof(R1CS_COEFS_##v) / 16, \
   R1CS_COLS_##v, sizeof(R1CS_COLS_##v) / sizeof(b3w_seg), R1CS_ROWS_##v, R1CS_TERMS_##v}

static const circuit_def CIRCUITS[] = {
    {"blake3_compression", false, B3W_WS_COMPRESSION, 28, 16, B3W_TRACE_WORDS_COMPRESSION, SEGS(COMPRESSION), PRIME_BN254, R1CS1(COMPRESSION_FUSED), R1CS1(COMPRESSION_SLOTS), 5,
     {{"h", 0, 8}, {"m", 8, 16}, {"t", 24, 2}, {"b", 26, 1}, {"d", 27, 1}}},
    {"blake3_nova (bn128, O2)", true, B3W_WS_NOVA_BN_O2, 32, 15, B3W_TRACE_WORDS_NOVA_BN_O2, SEGS(NOVA_BN_O2), PRIME_BN254, R1CS1(NOVA_FUSED), R1CS1(NOVA_O2_SLOTS), NOVA_SIGS},
    {"blake3_nova_pasta (vesta prime = Pallas scalar, O2)", true, B3W_WS_NOVA_PASTA_O2, 32, 15, B3W_TRACE_WORDS_NOVA_PASTA_O2,
     SEGS(NOVA_PASTA_O2), PRIME_PALLAS_SCALAR, R1CS1(NOVA_FUSED), R1CS1(NOVA_O2_SLOTS), NOVA_SIGS},
    {"blake3_nova (bn128, O1)", true, B3W_WS_NOVA_BN_O1, 32, 15, B3W_TRACE_WORDS_NOVA_BN_O1, SEGS(NOVA_BN_O1), PRIME_BN254, R1CS1(NOVA_FUSED), R1CS1(NOVA_BN_O1_SLOTS), NOVA_SIGS},
};
static const int N_CIRCUITS = sizeof(CIRCUITS) / sizeof(CIRCUITS[0]);

#define B3W_DEFAULT_CHUNK 4096u    // instances per ring slot: 8.4 M witnesses/s at 1 024, 10.5 at 2 048, 10.75 from 4 096 on (profiles/r02z_chunk_sweep.jsonl)
#define N_SCHED_COUNTERS 64       // launches in flight on different streams each need their own counter set
#define SCHED_SET_U64 (SCHED_LANES * SCHED_STRIDE)
#define B3W_RING_SLOTS 2
struct b3w_ctx {
  const circuit_def *def = nullptr;
  int device = 0;
  int sm_count = 0;
  uint32_t chunk = 0;
  int ctas_per_sm = 0;          // resident CTAs of this circuit's kernel (occupancy query)
  int ctas_per_sm_checked = 0;  // ... of its *_checked variant
  uint32_t *d_desc = nullptr;
  field_consts *d_field = nullptr;
  uint2 *d_fslots = nullptr;    // {slot, descriptor} of every slot that holds a true field element (nova)
  uint32_t n_fslots = 0;
  uint32_t *h_desc = nullptr;   // host copy of the per-slot descriptors
  field_consts *h_field = nullptr;   // host copy of the field constants (b3w_unpack_host)
  uint32_t flags = 0;
  // One host-buffer call at a time per context (include/blake3wit.h, "Ownership / threading"): the ring slots, streams and
  // scratch below are shared mutable state, so every entry point that touches them holds this lock for its duration.
  std::mutex mu;
  // R1CS tables (built on first use)
  bool r1cs_ready = false;
  struct r1cs_dev { r1cs_class_dev *cls; int64_t *lo, *hi; uint32_t *terms; uint32_t ncls; fr_t *coef_fr; uint32_t *row_ids, *nblk; uint32_t rows; };
  r1cs_dev r_fused = {}, r_slots = {};
  bool r1cs_loaded = false;     // r_slots comes from b3w_r1cs_load, not from the built-in tables
  fastprog_dev fp = {};         // compiled form of the slot-space rows (kernels_r1cs_fast.cuh); r_slots holds the residual rows only
  fastprog_dev fp0 = {};        // the same rows compiled without virtual bits (all-zero when fp has none: fp is then used for both)
  uint32_t slot_rows = 0;       // rows of the slot-space system (compiled + residual); 0 = none installed
  uint32_t fault_word = B3W_NO_ROW, fault_mask = 0;
  int ctas_limit = 0;           // tuning hook: cap on resident CTAs per SM (0 = occupancy limit)
  uint32_t sched_parts = 0;     // work items per instance (0 = default)
  int store_mode = 0;           // tuning hook: 0 = direct 256-bit stores, 1 = shared-memory tiles + TMA bulk stores
  // rotating pool of work-item counters: launches on different streams may overlap, so a counter pair is only handed out
  // again once the launch that used it last has finished (the new launch's stream waits on that launch's event)
  unsigned long long *d_counters = nullptr;
  uint32_t next_counter = 0;
  cudaEvent_t ctr_ev[N_SCHED_COUNTERS / 2] = {};
  // staging for host-buffer batches: 2 ring slots
  cudaStream_t st[2] = {};
  cudaEvent_t ev[2] = {};
  cudaEvent_t ev_k0[2] = {}, ev_k1[2] = {};    // timing: around the witness kernel of the chunk in flight on each slot
  bool ev_pending[2] = {};
  uint8_t *d_ring[2] = {};
  uint32_t *d_in[2] = {};
  uint8_t *d_status[2] = {};
  uint32_t *d_pub[2] = {};
  int8_t *d_ext[2] = {};         // compression only: m_ext of the chunk (wide batches)
  uint64_t *d_sums[2] = {};      // per-instance witness checksums of the chunk (b3w_batch_extras.sums)
  uint32_t *d_fbad[2] = {};      // fused check: first violated row of the chunk's instances
  uint8_t *d_fr[2] = {};         // Fr256 input rows of the chunk (b3w_witness_batch_fr), chunk x n_inputs x 32 bytes
  uint32_t *d_wlist[2] = {};     // nova: indices (inside the chunk) of the instances that hold a field-valued input; [0] = count
  bool ring_ready = false, fr_ready = false;
  uint32_t ring_cap = 0;         // instances per ring slot as allocated: grows with the largest batch seen, up to `chunk`
  uint32_t m_slot0 = 0;          // compression only: witness slot of m[0] (the 16 m slots are consecutive)
  // nova only, built on first use: the wide (field-element input) kernel's override list
  uint2 *d_wslots = nullptr;
  uint32_t *d_lane_off = nullptr;
  bool nw_ready = false;
  // compressible device memory (device_mem.h): driver entry points + the blocks handed out by b3w_device_alloc / the ring
  vmm_api vmm = {};
  std::vector<vmm_block> *blocks = nullptr;
  void *cs_ptr[12] = {};         // chain driver scratch (grow-only)
  size_t cs_cap[12] = {};
  // staging for host-buffer batches of PACKED witnesses: 2 slots
  cudaStream_t pk_st[2] = {};
  uint32_t *pk_in[2] = {}, *pk_buf[2] = {}, *pk_pub[2] = {};
  uint8_t *pk_status[2] = {};
  bool pk_ready = false;
  uint32_t *hy_host[2] = {};     // hybrid export: pinned host staging of the packed chunks
  bool hy_ready = false;
  b3w_timing timing = {};        // what b3w_last_timing() reports: the last host-buffer call of this context
};

// Every entry point runs on the context's device and leaves the CALLER's current device as it found it.
struct dev_guard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit dev_guard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) {
      err = cudaSetDevice(dev);
      switched = err == cudaSuccess;
    }
  }
  ~dev_guard() {
    if (switched) cudaSetDevice(prev);
  }
};
#define ON_DEVICE(c)                                                                                         \
  dev_guard dev_guard_(c->device);                                                                           \
  if (dev_guard_.err != cudaSuccess) return fail(B3W_ERR_CUDA, "cudaSetDevice(%d): %s", c->device, cudaGetErrorString(dev_guard_.err))
#define LOCKED(c) std::lock_guard<std::mutex> lock_guard_(c->mu)

// NVTX ranges around the stages of the host-buffer calls (tracing): visible in nsys / ncu timelines,
// free when no tool is attached.
struct nvtx_range {
  explicit nvtx_range(const char *name) { nvtxRangePushA(name); }
  ~nvtx_range() { nvtxRangePop(); }
};
// joins on every path out of the scope: a std::thread that is destroyed while joinable terminates the process
struct thread_group {
  std::vector<std::thread> th;
  void join() { for (auto &t : th) if (t.joinable()) t.join(); }
  ~thread_group() { join(); }
};
static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" int b3w_version(void) { return B3W_VERSION; }
extern "C" const char *b3w_last_error(void) { return g_err; }

static int b3w_create_impl(const b3w_config *cfg, b3w_ctx **out) {
  if (!cfg || !out) return fail(B3W_ERR_INVALID, "b3w_create: null argument");
  if (cfg->circuit >= (uint32_t)N_CIRCUITS) return fail(B3W_ERR_UNSUPPORTED, "b3w_create: circuit %u not built", cfg->circuit);
  if (cfg->flags & ~(uint32_t)(B3W_FLAG_FUSED_CHECK | B3W_FLAG_COMPRESSIBLE_RING | B3W_FLAG_PLAIN_RING | B3W_FLAG_REFERENCE_SIBLINGS | B3W_FLAG_BYTE_CHECK))
    return fail(B3W_ERR_INVALID, "b3w_create: unknown flags 0x%x", cfg->flags);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(B3W_ERR_CUDA, "no CUDA device: %s (libblake3wit has no CPU path)", cudaGetErrorString(e));
  int dev = cfg->device;
  if (dev < 0) CK(cudaGetDevice(&dev));
  if (dev >= ndev) return fail(B3W_ERR_INVALID, "device %d out of range (%d devices)", dev, ndev);
  b3w_ctx *c = new (std::nothrow) b3w_ctx();
  if (!c) return fail(B3W_ERR_NOMEM, "out of host memory");
  c->def = &CIRCUITS[cfg->circuit];
  c->device = dev;
  dev_guard dg(dev);                              // the caller's current device is restored on every path out
  if (dg.err != cudaSuccess) { delete c; return fail(B3W_ERR_CUDA, "cudaSetDevice(%d): %s", dev, cudaGetErrorString(dg.err)); }
  c->chunk = cfg->chunk ? cfg->chunk : B3W_DEFAULT_CHUNK;
  c->flags = cfg->flags;
  c->fault_word = B3W_NO_ROW;
  cudaError_t e1 = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e1 != cudaSuccess) { delete c; return fail(B3W_ERR_CUDA, "device attribute query: %s", cudaGetErrorString(e1)); }
  // expand the run-length table to one descriptor per slot and upload it
  const circuit_def *d = c->def;
  uint32_t *h = (uint32_t *)malloc((size_t)d->ws * 4);
  if (!h) { delete c; return fail(B3W_ERR_NOMEM, "out of host memory"); }
  uint32_t pos = 0;
  for (size_t i = 0; i < d->n_segs; i++)
    for (uint32_t j = 0; j < d->segs[i].count; j++) h[pos++] = d->segs[i].desc0 + j * d->segs[i].delta;
  if (pos != d->ws) { free(h); delete c; return fail(B3W_ERR_INVALID, "slot table of %s is corrupt", d->name); }
  e1 = cudaMalloc(&c->d_desc, (size_t)d->ws * 4);
  if (e1 == cudaSuccess) e1 = cudaMemcpy(c->d_desc, h, (size_t)d->ws * 4, cudaMemcpyHostToDevice);
  c->h_desc = h;
  if (e1 != cudaSuccess) { b3w_destroy(c); return fail(B3W_ERR_CUDA, "descriptor upload: %s", cudaGetErrorString(e1)); }
  if (!d->nova) {
    // the wide-domain kernels rewrite the witness slots of m[0..15]: exactly 16 consecutive W32 slots read TR_IN + 8 + j
    uint32_t first = 0, found = 0;
    for (uint32_t sl = 0; sl < d->ws; sl++) {
      const uint32_t t = h[sl] & 0xFFFFu;
      if ((h[sl] >> 24) == DK_W32 && t >= TR_IN + 8 && t < TR_IN + 24) {
        if (found == 0) first = sl;
        if (sl != first + found || t != TR_IN + 8 + found) { found = 99; break; }
        found++;
      }
    }
    if (found != 16) { b3w_destroy(c); return fail(B3W_ERR_INVALID, "slot table of %s: the m slots are not where the wide path expects them", d->name); }
    c->m_slot0 = first;
  }
  {
    std::vector<uint2> fs;
    for (uint32_t sl = 0; sl < d->ws; sl++)
      if ((h[sl] >> 24) >= DK_S64) fs.push_back(make_uint2(sl, h[sl]));
    c->n_fslots = (uint32_t)fs.size();
    if (c->n_fslots) {
      e1 = cudaMalloc(&c->d_fslots, fs.size() * sizeof(uint2));
      if (e1 == cudaSuccess) e1 = cudaMemcpy(c->d_fslots, fs.data(), fs.size() * sizeof(uint2), cudaMemcpyHostToDevice);
      if (e1 != cudaSuccess) { b3w_destroy(c); return fail(B3W_ERR_CUDA, "field slot list upload: %s", cudaGetErrorString(e1)); }
    }
  }
  {
    field_consts *F = new (std::nothrow) field_consts();
    if (!F) { b3w_destroy(c); return fail(B3W_ERR_NOMEM, "out of host memory"); }
    uint32_t pl[8];
    memcpy(pl, d->prime, 32);
    field_consts_init(*F, pl);
    e1 = cudaMalloc(&c->d_field, sizeof(field_consts));
    if (e1 == cudaSuccess) e1 = cudaMemcpy(c->d_field, F, sizeof(field_consts), cudaMemcpyHostToDevice);
    c->h_field = F;
    if (e1 != cudaSuccess) { b3w_destroy(c); return fail(B3W_ERR_CUDA, "field table upload: %s", cudaGetErrorString(e1)); }
  }
  const int bs_plain = WARPS_PER_CTA * 32, bs_checked = (WARPS_PER_CTA + (d->nova ? NOVA_CHECK_WARPS : CHECK_WARPS)) * 32;
  if (d->nova) {
    e1 = cudaFuncSetAttribute(k_blake3_nova_witness<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NOVA_SMEM(WARPS_PER_CTA + NOVA_CHECK_WARPS));
    if (e1 == cudaSuccess)
      e1 = cudaFuncSetAttribute(k_blake3_nova_witness<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NOVA_SMEM(WARPS_PER_CTA + NOVA_CHECK_WARPS));
    if (e1 == cudaSuccess)
      e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->ctas_per_sm, k_blake3_nova_witness<false, true>, bs_plain, NOVA_SMEM(WARPS_PER_CTA));
    if (e1 == cudaSuccess)
      e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->ctas_per_sm_checked, k_blake3_nova_witness<true, true>, bs_checked,
                                                         NOVA_SMEM(WARPS_PER_CTA + NOVA_CHECK_WARPS));
  } else {
    e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->ctas_per_sm, k_blake3_comp_witness<false, false, true>, bs_plain, 0);
    if (e1 == cudaSuccess)
      e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->ctas_per_sm_checked, k_blake3_comp_witness<true, false, true>, bs_checked, 0);
  }
  if (e1 != cudaSuccess || c->ctas_per_sm < 1 || c->ctas_per_sm_checked < 1) { b3w_destroy(c); return fail(B3W_ERR_CUDA, "occupancy query: %s", cudaGetErrorString(e1)); }
  *out = c;
  return B3W_OK;
}
extern "C" int b3w_create(const b3w_config *cfg, b3w_ctx **out) {
  return guarded("b3w_create", [&]() { return b3w_create_impl(cfg, out); });
}

// ---- device memory for witness buffers: ordinary or compressible (device_mem.h) ------------------------------------
static int device_alloc(b3w_ctx *c, size_t bytes, uint32_t flags, void **out) {
  *out = nullptr;
  if (flags & ~(uint32_t)B3W_MEM_COMPRESSIBLE) return fail(B3W_ERR_INVALID, "b3w_device_alloc: unknown flags 0x%x", flags);
  if (bytes == 0) return fail(B3W_ERR_INVALID, "b3w_device_alloc: zero bytes");
  if (!c->vmm.ok && !vmm_load(c->vmm)) return fail(B3W_ERR_UNSUPPORTED, "b3w_device_alloc: the driver's virtual-memory entry points are not available");
  if (!c->blocks) c->blocks = new std::vector<vmm_block>();
  vmm_block blk;
  const char *err;
  void *p = vmm_alloc(c->vmm, c->device, bytes, (flags & B3W_MEM_COMPRESSIBLE) != 0, blk, &err);
  if (!p) return fail(B3W_ERR_NOMEM, "b3w_device_alloc(%zu bytes): %s", bytes, err);
  c->blocks->push_back(blk);
  *out = p;
  return B3W_OK;
}
// B3W_OK when p was one of this context's blocks (and is now released), B3W_ERR_INVALID otherwise
static int device_free(b3w_ctx *c, void *p) {
  if (c->blocks)
    for (size_t i = 0; i < c->blocks->size(); i++)
      if ((void *)(*c->blocks)[i].va == p) {
        vmm_free(c->vmm, (*c->blocks)[i]);
        c->blocks->erase(c->blocks->begin() + i);
        return B3W_OK;
      }
  return B3W_ERR_INVALID;
}
extern "C" int b3w_device_alloc(b3w_ctx *c, size_t bytes, uint32_t flags, void **out, uint32_t *granted) {
  if (!c || !out) return fail(B3W_ERR_INVALID, "b3w_device_alloc: null argument");
  ON_DEVICE(c);
  LOCKED(c);
  return guarded("b3w_device_alloc", [&]() {
    const int rc = device_alloc(c, bytes, flags, out);
    if (rc == B3W_OK && granted) *granted = c->blocks->back().compressed ? B3W_MEM_COMPRESSIBLE : 0u;
    return rc;
  });
}
extern "C" int b3w_device_free(b3w_ctx *c, void *p) {
  if (!c) return fail(B3W_ERR_INVALID, "b3w_device_free: null argument");
  if (!p) return B3W_OK;
  ON_DEVICE(c);
  LOCKED(c);
  CK(cudaDeviceSynchronize());                               // nothing may still be using the mapping
  if (device_free(c, p) != B3W_OK) return fail(B3W_ERR_INVALID, "b3w_device_free: %p was not allocated by b3w_device_alloc of this context", p);
  return B3W_OK;
}

static void free_packed_ring(b3w_ctx *c);
static void free_r1cs_dev(b3w_ctx::r1cs_dev *r) {
  if (r->cls) cudaFree(r->cls);
  if (r->lo) cudaFree(r->lo);
  if (r->hi) cudaFree(r->hi);
  if (r->terms) cudaFree(r->terms);
  if (r->coef_fr) cudaFree(r->coef_fr);
  if (r->row_ids) cudaFree(r->row_ids);
  if (r->nblk) cudaFree(r->nblk);
  memset(r, 0, sizeof *r);
}
static void free_ring(b3w_ctx *c) {
  for (int k = 0; k < 2; k++) {
    if (c->d_ring[k] && device_free(c, c->d_ring[k]) != B3W_OK) cudaFree(c->d_ring[k]);
    for (void **q : {(void **)&c->d_in[k], (void **)&c->d_status[k], (void **)&c->d_pub[k], (void **)&c->d_ext[k], (void **)&c->d_sums[k],
                     (void **)&c->d_fbad[k], (void **)&c->d_fr[k], (void **)&c->d_wlist[k]}) {
      if (*q) cudaFree(*q);
      *q = nullptr;
    }
    if (c->st[k]) cudaStreamDestroy(c->st[k]);
    for (cudaEvent_t *e : {&c->ev[k], &c->ev_k0[k], &c->ev_k1[k]}) {
      if (*e) cudaEventDestroy(*e);
      *e = nullptr;
    }
    c->d_ring[k] = nullptr;
    c->st[k] = nullptr;
    c->ev_pending[k] = false;
  }
  c->ring_ready = false;
  c->fr_ready = false;
  c->ring_cap = 0;
}

static void free_fastprog(fastprog_dev *p);
extern "C" void b3w_destroy(b3w_ctx *c) {
  if (!c) return;
  {
    dev_guard dg(c->device);
    cudaDeviceSynchronize();                         // nothing of this context may still run (mapped blocks are unmapped below)
    free_ring(c);
    free_packed_ring(c);
    for (int i = 0; i < 12; i++)
      if (c->cs_ptr[i]) cudaFree(c->cs_ptr[i]);
    if (c->d_desc) cudaFree(c->d_desc);
    if (c->d_field) cudaFree(c->d_field);
    if (c->d_fslots) cudaFree(c->d_fslots);
    if (c->d_counters) cudaFree(c->d_counters);
    for (cudaEvent_t e : c->ctr_ev)
      if (e) cudaEventDestroy(e);
    if (c->blocks) {
      for (const vmm_block &b : *c->blocks) vmm_free(c->vmm, b);
      delete c->blocks;
    }
    if (c->d_wslots) cudaFree(c->d_wslots);
    if (c->d_lane_off) cudaFree(c->d_lane_off);
    for (b3w_ctx::r1cs_dev *r : {&c->r_slots, &c->r_fused}) free_r1cs_dev(r);
    free_fastprog(&c->fp);
    free_fastprog(&c->fp0);
    free(c->h_desc);
    delete c->h_field;
  }
  delete c;
}

static const circuit_def *find_def(uint32_t circuit) {
  if (circuit >= (uint32_t)N_CIRCUITS) {
    fail(B3W_ERR_UNSUPPORTED, "circuit %u not built", circuit);
    return nullptr;
  }
  return &CIRCUITS[circuit];
}

extern "C" int b3w_circuit_info(uint32_t circuit, b3w_info *info) {
  const circuit_def *d = find_def(circuit);
  if (!d) return B3W_ERR_UNSUPPORTED;
  if (!info) return fail(B3W_ERR_INVALID, "b3w_circuit_info: null argument");
  info->witness_size = d->ws;
  info->n_inputs = d->n_inputs;
  info->n32 = 8;
  info->n_public = d->n_public;
  info->version[0] = 2; info->version[1] = 1; info->version[2] = 6;
  memcpy(info->prime, d->prime, 32);
  return B3W_OK;
}

extern "C" int b3w_wtns_header(uint32_t circuit, uint8_t hdr[76]) {
  const circuit_def *d = find_def(circuit);
  if (!d) return B3W_ERR_UNSUPPORTED;
  if (!hdr) return fail(B3W_ERR_INVALID, "b3w_wtns_header: null argument");
  // witness_calculator.js:214-262 with n32 = 8
  uint32_t w[19];
  memcpy(&w[0], "wtns", 4);
  w[1] = 2;                     // version
  w[2] = 2;                     // sections
  w[3] = 1;                     // section 1 id
  w[4] = 40; w[5] = 0;          // section 1 length (u64) = 8 + n8
  w[6] = 32;                    // n8
  memcpy(&w[7], d->prime, 32);
  w[15] = d->ws;
  w[16] = 2;                    // section 2 id
  uint64_t len2 = 32ull * d->ws;
  w[17] = (uint32_t)len2; w[18] = (uint32_t)(len2 >> 32);
  memcpy(hdr, w, 76);
  return B3W_OK;
}

extern "C" int b3w_input_signal(uint32_t circuit, const char *name, uint32_t *offset, uint32_t *size) {
  const circuit_def *d = find_def(circuit);
  if (!d) return B3W_ERR_UNSUPPORTED;
  if (!name) return fail(B3W_ERR_INVALID, "b3w_input_signal: null argument");
  for (int i = 0; i < d->n_sig; i++)
    if (strcmp(d->sig[i].name, name) == 0) {
      if (offset) *offset = d->sig[i].off;
      if (size) *size = d->sig[i].size;
      return B3W_OK;
    }
  return fail(B3W_ERR_INVALID, "Signal %s not found", name);
}

// The text the reference's wasm emits through printErrorMessage when an assert of the nova step circuit fires
// (witness_calculator.js:21-43 appends it to "Assert Failed.\n"): one line per template on the call stack, innermost
// first.  Line numbers are those of the circuit AS BUILT into the committed wasm files (circuits/blake3_nova.circom
// without :25-30, circomlib 2.0.5 bitify/comparators); identical in all three nova builds.  Host-only, needs no GPU.
extern "C" int b3w_assert_trace(uint32_t circuit, const uint32_t *in, char *buf, size_t cap) {
  const circuit_def *d = find_def(circuit);
  if (!d) return B3W_ERR_UNSUPPORTED;
  if (!in || (!buf && cap)) return fail(B3W_ERR_INVALID, "b3w_assert_trace: null argument");
  if (cap) buf[0] = 0;
  if (!d->nova) return B3W_OK;                   // u32 inputs cannot violate a constraint of Blake3Compression
  const int64_t leaf_depth = in[12], depth = in[14];
  const int64_t v1 = depth + 256 - (leaf_depth - 1);     // check_parent = LessThan(8)(depth, leaf_depth - 1)   (:31-33 -> :27 as built)
  const int64_t v2 = leaf_depth + 256 - (depth + 1);     // exceed_depth = GreaterEqThan(8)(depth, leaf_depth)  (:41-43 -> :37)
  const char *msg = nullptr;
  if (v1 < 0 || v1 >= 512)
    msg = "Error in template Num2Bits_2 line: 38\nError in template LessThan_3 line: 96\n"
          "Error in template Blake3NovaTreePath_CheckDepth_5 line: 27\nError in template Blake3Nova_54 line: 201\n";
  else if (v2 < 0 || v2 >= 512)
    msg = "Error in template Num2Bits_2 line: 38\nError in template LessThan_3 line: 96\nError in template GreaterEqThan_4 line: 138\n"
          "Error in template Blake3NovaTreePath_CheckDepth_5 line: 37\nError in template Blake3Nova_54 line: 201\n";
  else if (((v2 >> 8) & 1) == 0)                         // exceed_depth.out === 0                              (:44 -> :38)
    msg = "Error in template Blake3NovaTreePath_CheckDepth_5 line: 38\nError in template Blake3Nova_54 line: 201\n";
  if (!msg) return B3W_OK;
  if (cap) snprintf(buf, cap, "%s", msg);
  return B3W_CIRCOM_ASSERT;
}

// Expand the class / coefficient / column tables of one R1CS row set (r1cs_tables.h): host only.
struct r1cs_expanded {
  std::vector<r1cs_class_dev> cls;
  std::vector<int64_t> lo, hi;
  std::vector<uint32_t> terms;           // [class][term][row] matrices
};
static bool expand_r1cs_set(const circuit_def::r1cs_set &set, r1cs_expanded &x) {
  const size_t ncls = set.ncls;
  x.cls.resize(ncls);
  x.lo.resize(set.ncoef);
  x.hi.resize(set.ncoef);
  x.terms.assign(set.terms, 0);
  for (size_t i = 0; i < set.ncoef; i++) { x.lo[i] = (int64_t)set.coef[i][0]; x.hi[i] = (int64_t)set.coef[i][1]; }
  uint32_t term_off = 0, row_off = 0;
  for (size_t k = 0; k < ncls; k++) {
    const b3w_r1cs_class &s = set.cls[k];
    x.cls[k] = r1cs_class_dev{s.nA, s.nB, s.nC, s.flags, s.count, s.coef_off, term_off, row_off};
    size_t pos = s.col_off;
    for (uint32_t t = 0; t < (uint32_t)(s.nA + s.nB + s.nC); t++) {
      if (pos >= set.ncols || set.cols[pos].desc0 != 0xFFFFFFFFu || term_off + s.count > set.terms) return false;
      uint32_t nruns = set.cols[pos++].count, w = 0;
      for (uint32_t r = 0; r < nruns; r++, pos++)
        for (uint32_t j = 0; j < set.cols[pos].count; j++) {
          if (w >= s.count) return false;
          x.terms[term_off + w++] = set.cols[pos].desc0 + j * (uint32_t)set.cols[pos].delta;
        }
      if (w != s.count) return false;
      term_off += s.count;
    }
    row_off += s.count;
  }
  return term_off == set.terms && row_off == set.rows;
}

// the fused (trace-space) set: class matrices as they are
static int upload_fused_set(b3w_ctx *c, const circuit_def::r1cs_set &set, b3w_ctx::r1cs_dev *out) {
  if (set.ncls == 0) return B3W_OK;
  r1cs_expanded x;
  if (!expand_r1cs_set(set, x)) return fail(B3W_ERR_INVALID, "R1CS tables of %s are corrupt", c->def->name);
  const size_t ncls = x.cls.size();
  cudaError_t e = cudaMalloc(&out->cls, ncls * sizeof(r1cs_class_dev));
  if (e == cudaSuccess) e = cudaMalloc(&out->lo, x.lo.size() * 8);
  if (e == cudaSuccess) e = cudaMalloc(&out->hi, x.hi.size() * 8);
  if (e == cudaSuccess) e = cudaMalloc(&out->terms, x.terms.size() * 4 + 16);
  if (e == cudaSuccess) e = cudaMemcpy(out->cls, x.cls.data(), ncls * sizeof(r1cs_class_dev), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(out->lo, x.lo.data(), x.lo.size() * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(out->hi, x.hi.data(), x.hi.size() * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(out->terms, x.terms.data(), x.terms.size() * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail(B3W_ERR_CUDA, "R1CS table upload: %s", cudaGetErrorString(e));
  out->ncls = (uint32_t)ncls;
  return B3W_OK;
}

template <class T>
static cudaError_t upload_vec(T **dst, const std::vector<T> &v) {
  cudaError_t e = cudaMalloc((void **)dst, v.size() * sizeof(T) + 16);
  if (e == cudaSuccess && !v.empty()) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}
static void free_fastprog(fastprog_dev *p) {
  for (void *q : {(void *)p->bool_mask, (void *)p->bool_row, (void *)p->xors, (void *)p->xor_ids, (void *)p->tiles, (void *)p->vtiles, (void *)p->items, (void *)p->row_ids, (void *)p->side_rank})
    if (q) cudaFree(q);
  memset(p, 0, sizeof *p);
}

// The slot-space system of the stand-alone check (built-in rows, or the rows of a loaded .r1cs file): compile what
// compiles (fp_compile), keep the rest as a residual class/block set for the general evaluator, upload both.
static cudaError_t upload_fastprog(fastprog_dev *P, const fastprog_host &fp) {
  memset(P, 0, sizeof *P);
  cudaError_t e = upload_vec(&P->bool_mask, fp.bool_mask);
  if (e == cudaSuccess) e = upload_vec(&P->bool_row, fp.bool_row);
  if (e == cudaSuccess) e = upload_vec(&P->xors, fp.xors);
  if (e == cudaSuccess) e = upload_vec(&P->xor_ids, fp.xor_ids);
  if (e == cudaSuccess) e = upload_vec(&P->tiles, fp.tiles);
  if (e == cudaSuccess) e = upload_vec(&P->vtiles, fp.vtiles);
  if (e == cudaSuccess) e = upload_vec(&P->items, fp.items);
  if (e == cudaSuccess) e = upload_vec(&P->row_ids, fp.row_ids);
  P->n_xors = (uint32_t)fp.xors.size();
  P->n_tiles = (uint32_t)fp.tiles.size();
  P->n_vtiles = (uint32_t)fp.vtiles.size();
  P->n_rows = fp.n_rows;
  return e;
}
// fp_compile with virtual bits where the system has them, plus the plain compilation the kernel falls back to for a
// witness whose virtual bits are not bits; both cover the same rows (fp0 stays empty when there are no virtual bits)
static void compile_programs(const std::vector<r1cs_load_detail::row> &rows, uint32_t ws, fastprog_host &fp, fastprog_host &fp0, std::vector<char> &taken,
                             const std::vector<uint8_t> *wide_hint = nullptr) {
  fp_compile(rows, ws, fp, taken, /*with_virtuals=*/true, wide_hint);
  if (fp.n_virtual == 0) return;
  std::vector<char> taken0;
  fp_compile(rows, ws, fp0, taken0, false, wide_hint);
  if (taken0 != taken) {                                      // cannot happen (fp_compile keeps the two in step); be safe
    fp = std::move(fp0);
    fp0 = fastprog_host();
    taken.swap(taken0);
  }
}
// Where the stand-alone checker keeps the non-bit slots of a witness (kernels_r1cs_fast.cuh): the side-table entries of
// 32-slot word w start at rank[w], by the CIRCUIT's slot kinds (a slot that is not DK_BIT may still hold 0 or 1: its entry
// is then simply not used).  rank has one entry per map word (slot words + virtual-bit words + 1); returns the table size.
static uint32_t side_layout(const uint32_t *desc, uint32_t ws, uint32_t n_vtiles, std::vector<uint32_t> &rank) {
  const uint32_t words = (ws + 31u) >> 5, mw = words + n_vtiles + 1u;
  rank.assign(mw, 0u);
  uint32_t total = 0;
  for (uint32_t w = 0; w < mw; w++) {
    rank[w] = total;
    for (uint32_t sl = w * 32u; w < words && sl < std::min(ws, (w + 1u) * 32u); sl++) total += (desc[sl] >> 24) != DK_BIT;
  }
  return (total + 1u) & ~1u;
}
extern "C" int b3w_debug_side_layout(uint32_t circuit, uint32_t n_vtiles, uint32_t *rank_out, uint32_t cap, uint32_t *n_words, uint32_t *total) {
  const circuit_def *d = find_def(circuit);
  if (!d) return fail(B3W_ERR_UNSUPPORTED, "unknown circuit %u", circuit);
  return guarded("b3w_debug_side_layout", [&]() {
    std::vector<uint32_t> desc, rank;
    for (size_t i = 0; i < d->n_segs; i++)
      for (uint32_t j = 0; j < d->segs[i].count; j++) desc.push_back(d->segs[i].desc0 + j * d->segs[i].delta);
    if (desc.size() != d->ws) return fail(B3W_ERR_INVALID, "slot table of %s is corrupt", d->name);
    const uint32_t t = side_layout(desc.data(), d->ws, n_vtiles, rank);
    if (n_words) *n_words = (uint32_t)rank.size();
    if (total) *total = t;
    if (rank_out) {
      if (cap < rank.size()) return fail(B3W_ERR_INVALID, "b3w_debug_side_layout: %u words, room for %u", (uint32_t)rank.size(), cap);
      memcpy(rank_out, rank.data(), rank.size() * 4);
    }
    return B3W_OK;
  });
}

static int install_slot_rows(b3w_ctx *c, std::vector<r1cs_load_detail::row> &rows, uint32_t *n_compiled) {
  fastprog_host fp, fp0;
  std::vector<char> taken;
  std::vector<uint8_t> wide(c->def->ws);                      // slots that hold field-valued quantities by their kind (IsZero's inverse)
  for (uint32_t sl = 0; sl < c->def->ws; sl++) wide[sl] = (c->h_desc[sl] >> 24) == DK_INV;
  compile_programs(rows, c->def->ws, fp, fp0, taken, &wide);
  if (fp.n_virtual && ((c->def->ws + 31u) >> 5) + fp.vtiles.size() + 1u > FP_MAPW) {     // more virtual-bit words than the kernel's maps hold:
    fp = std::move(fp0);                                                                   // the plainly compiled program alone (same rows)
    fp0 = fastprog_host();
  }
  std::vector<r1cs_load_detail::row> rest;
  for (size_t i = 0; i < rows.size(); i++)
    if (!taken[i]) rest.push_back(std::move(rows[i]));
  r1cs_host_set h;
  r1cs_group(rest, h);
  std::vector<uint32_t> blocks, nblk;
  stg_blockify(h.cls, h.terms, blocks, nblk, h.lo, h.hi);
  h.terms.swap(blocks);
  b3w_ctx::r1cs_dev d;
  memset(&d, 0, sizeof d);
  fastprog_dev P, P0;
  memset(&P0, 0, sizeof P0);
  cudaError_t e = upload_vec(&d.cls, h.cls);
  if (e == cudaSuccess) e = upload_vec(&d.nblk, nblk);
  if (e == cudaSuccess) e = upload_vec(&d.lo, h.lo);
  if (e == cudaSuccess) e = upload_vec(&d.hi, h.hi);
  if (e == cudaSuccess) e = upload_vec(&d.terms, h.terms);
  if (e == cudaSuccess) e = upload_vec(&d.coef_fr, h.coef_fr);
  if (e == cudaSuccess) e = upload_vec(&d.row_ids, h.row_ids);
  if (e == cudaSuccess) e = upload_fastprog(&P, fp);
  if (e == cudaSuccess && fp.n_virtual) e = upload_fastprog(&P0, fp0);
  if (e == cudaSuccess) {
    std::vector<uint32_t> rank;
    P.side_total = side_layout(c->h_desc, c->def->ws, P.n_vtiles, rank);
    e = upload_vec(&P.side_rank, rank);
  }
  if (e != cudaSuccess) {
    free_r1cs_dev(&d);
    free_fastprog(&P);
    free_fastprog(&P0);
    return fail(B3W_ERR_CUDA, "R1CS table upload: %s", cudaGetErrorString(e));
  }
  d.ncls = (uint32_t)h.cls.size();
  d.rows = h.rows;
  free_r1cs_dev(&c->r_slots);
  free_fastprog(&c->fp);
  free_fastprog(&c->fp0);
  c->r_slots = d;
  c->fp = P;
  c->fp0 = P0;
  c->slot_rows = (uint32_t)rows.size();
  if (n_compiled) *n_compiled = fp.n_rows;
  return B3W_OK;
}

// built-in slot-space rows (r1cs_tables.h) -> the row list install_slot_rows takes; row id = position in class order
static bool rows_from_builtin(const circuit_def::r1cs_set &set, std::vector<r1cs_load_detail::row> &rows) {
  r1cs_expanded x;
  if (!expand_r1cs_set(set, x)) return false;
  rows.clear();
  rows.reserve(set.rows);
  for (const r1cs_class_dev &k : x.cls) {
    const uint32_t n[3] = {k.nA, k.nB, k.nC};
    for (uint32_t r = 0; r < k.count; r++) {
      r1cs_load_detail::row R;
      R.id = k.row_off + r;
      R.big = false;
      uint32_t t = 0;
      for (int part = 0; part < 3; part++)
        for (uint32_t j = 0; j < n[part]; j++, t++) {
          const uint32_t ci = (k.flags & R1CS_FLAG_ROWCOEF) ? k.coef_off + t * k.count + r : k.coef_off + t;
          r1cs_load_detail::term T;
          T.wire = x.terms[k.term_off + (size_t)t * k.count + r];
          T.small = true;
          T.c = (__int128)(((unsigned __int128)(uint64_t)x.hi[ci] << 64) | (unsigned __int128)(uint64_t)x.lo[ci]);
          T.f = fr_zero();                                  // built-in coefficients are small integers: never read
          R.part[part].push_back(T);
        }
      if (R.part[0].empty() || R.part[1].empty()) { R.part[0].clear(); R.part[1].clear(); }
      rows.push_back(std::move(R));
    }
  }
  return true;
}

static int ensure_r1cs(b3w_ctx *c) {
  if (c->r1cs_ready) return B3W_OK;
  int rc = c->r_fused.ncls ? B3W_OK : upload_fused_set(c, c->def->r_fused, &c->r_fused);
  if (rc == B3W_OK && !c->r1cs_loaded && c->def->r_slots.ncls) {
    std::vector<r1cs_load_detail::row> rows;
    if (!rows_from_builtin(c->def->r_slots, rows)) return fail(B3W_ERR_INVALID, "R1CS tables of %s are corrupt", c->def->name);
    rc = install_slot_rows(c, rows, nullptr);
  }
  if (rc == B3W_OK) c->r1cs_ready = true;
  return rc;
}

// where the witnesses of a launch go decides the work split (see launch_witness)
static bool in_compressible_block(b3w_ctx *c, const void *p) {
  if (c->blocks)
    for (const vmm_block &b : *c->blocks)
      if (b.compressed && (CUdeviceptr)p >= b.va && (CUdeviceptr)p < b.va + b.size) return true;
  return false;
}

// A pair of work-item counter sets for one launch on stream s.  The pool rotates; before a pair is handed out again the
// new launch's stream is made to wait for the launch that used it last (an event per pair), so more launches in flight
// than the pool holds are serialised instead of corrupting each other's counters.
static int take_counters(b3w_ctx *c, cudaStream_t s, unsigned long long **out, uint32_t *pair_out) {
  if (!c->d_counters) CK(cudaMalloc(&c->d_counters, (size_t)N_SCHED_COUNTERS * SCHED_SET_U64 * sizeof(unsigned long long)));
  const uint32_t pair = c->next_counter % (N_SCHED_COUNTERS / 2);
  c->next_counter++;
  if (!c->ctr_ev[pair]) CK(cudaEventCreateWithFlags(&c->ctr_ev[pair], cudaEventDisableTiming));
  else CK(cudaStreamWaitEvent(s, c->ctr_ev[pair], 0));
  *out = c->d_counters + (size_t)pair * 2 * SCHED_SET_U64;
  *pair_out = pair;
  CK(cudaMemsetAsync(*out, 0, 2 * SCHED_SET_U64 * sizeof(unsigned long long), s));
  return B3W_OK;
}

struct launch_opts {
  bool check = false;
  uint32_t *d_first_bad = nullptr;
  const int8_t *d_m_ext = nullptr;       // compression: the wide-domain kernel
  unsigned long long *d_sums = nullptr;  // per-instance checksums (zeroed here)
};

template <bool CHECK, bool SUMS>
static void launch_nova(b3w_ctx *c, unsigned grid, unsigned bs, cudaStream_t s, const uint32_t *d_in, uint64_t n, uint8_t *d_out,
                        uint8_t *d_status, uint32_t *d_pub, const check_args &ck, const sched_args &sc, const sched_args &sck) {
  k_blake3_nova_witness<CHECK, SUMS><<<grid, bs, NOVA_SMEM(WARPS_PER_CTA + (CHECK ? NOVA_CHECK_WARPS : 0)), s>>>(
      d_in, n, c->d_desc, c->def->ws, c->d_field, c->d_fslots, c->n_fslots, d_out, d_status, d_pub, ck, sc, sck);
}
template <bool CHECK, bool WIDE, bool SUMS>
static void launch_comp(b3w_ctx *c, unsigned grid, unsigned bs, cudaStream_t s, const uint32_t *d_in, uint64_t n, uint8_t *d_out,
                        uint8_t *d_status, uint32_t *d_pub, const check_args &ck, const sched_args &sc, const sched_args &sck,
                        const wide_args &wd) {
  k_blake3_comp_witness<CHECK, WIDE, SUMS><<<grid, bs, 0, s>>>(d_in, n, c->d_desc, c->def->ws, d_out, d_status, d_pub, ck, sc, sck, wd);
}

static int launch_witness(b3w_ctx *c, const uint32_t *d_in, uint64_t n, uint8_t *d_out, uint8_t *d_status,
                          uint32_t *d_pub, cudaStream_t s, const launch_opts &o = launch_opts()) {
  if (n == 0) return B3W_OK;
  const bool check = o.check;
  if (o.d_m_ext && c->def->nova) return fail(B3W_ERR_UNSUPPORTED, "%s: only blake3_compression has a wide-domain kernel", c->def->name);
  // Persistent grid, work items handed out dynamically (see sched_args).  Defaults from sweeps on B200 (profiles/):
  // fastest with only 2 CTAs per SM (16 expansion warps) and 24 items per witness (32 KiB each: the GPU-wide write front
  // stays compact).  The checked kernels have the same expansion shape plus CHECK_WARPS checker warps per CTA.
  // ... unless the witnesses go to COMPRESSIBLE memory (one of this context's b3w_device_alloc blocks): HBM is then no
  // longer the limit, the SM-side store path and the per-item trace recomputation are, and 12 items per witness measured
  // best across the four kernels (profiles/r01j_compressible.jsonl: nova + fused check 7.44 -> 6.49 ms per 2^16).
  const bool compressed_out = in_compressible_block(c, d_out);
  const uint32_t parts = c->sched_parts ? c->sched_parts : compressed_out ? 12u : 24u;
  uint64_t ctas_needed = (n * parts + WARPS_PER_CTA - 1) / WARPS_PER_CTA;      // one warp per work item
  int per_sm = check ? c->ctas_per_sm_checked : c->ctas_per_sm;
  const int cap = c->ctas_limit > 0 ? c->ctas_limit : 2;
  if (cap < per_sm) per_sm = cap;
  uint64_t max_ctas = (uint64_t)c->sm_count * per_sm;
  unsigned grid = (unsigned)(ctas_needed < max_ctas ? ctas_needed : max_ctas);
  check_args ck;
  memset(&ck, 0, sizeof ck);
  ck.fault_word = B3W_NO_ROW;
  ck.sums = o.d_sums;
  if (check) {
    int rc = ensure_r1cs(c);
    if (rc) return rc;
    ck.T = r1cs_tables_dev{c->r_fused.cls, c->r_fused.lo, c->r_fused.hi, c->r_fused.terms, c->r_fused.ncls};
    ck.F = c->d_field;
    ck.first_bad = o.d_first_bad;
    ck.fault_word = c->fault_word;
    ck.fault_mask = c->fault_mask;
  }
  const unsigned bs = (WARPS_PER_CTA + (check ? (c->def->nova ? NOVA_CHECK_WARPS : CHECK_WARPS) : 0)) * 32;
  // work distribution (see sched_args): expansion items, and -- checked kernels -- one check item per instance
  sched_args sc, sck;
  uint32_t pair;
  int rc = take_counters(c, s, &sc.counter, &pair);
  if (rc) return rc;
  sc.parts = parts;
  sc.part_len = ((c->def->ws + sc.parts - 1) / sc.parts + 31) / 32 * 32;
  sck.counter = sc.counter + SCHED_SET_U64;
  sck.parts = 1;
  sck.part_len = 0;
  if (o.d_sums) CK(cudaMemsetAsync(o.d_sums, 0, n * sizeof(unsigned long long), s));
  const bool sums = o.d_sums != nullptr;
  if (c->store_mode == 1 && !c->def->nova && !check && !o.d_m_ext && !sums) {
    // experiment: shared-memory tiles + TMA bulk stores (kernels_witness.cuh, k_blake3_comp_witness_tma)
    const int smem = WARPS_PER_CTA * TMA_NBUF * TMA_TILE_BYTES;
    CK(cudaFuncSetAttribute(k_blake3_comp_witness_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_blake3_comp_witness_tma<<<grid, WARPS_PER_CTA * 32, smem, s>>>(d_in, n, c->d_desc, c->def->ws, d_out, d_status, d_pub, sc);
  } else if (c->def->nova) {
    if (check) { if (sums) launch_nova<true, true>(c, grid, bs, s, d_in, n, d_out, d_status, d_pub, ck, sc, sck); else launch_nova<true, false>(c, grid, bs, s, d_in, n, d_out, d_status, d_pub, ck, sc, sck); }
    else { if (sums) launch_nova<false, true>(c, grid, bs, s, d_in, n, d_out, d_status, d_pub, ck, sc, sck); else launch_nova<false, false>(c, grid, bs, s, d_in, n, d_out, d_status, d_pub, ck, sc, sck); }
  } else {
    const wide_args wd{o.d_m_ext, c->d_field, c->m_slot0};
#define B3W_COMP(CH, WI, SU) launch_comp<CH, WI, SU>(c, grid, bs, s, d_in, n, d_out, d_status, d_pub, ck, sc, sck, wd)
    if (o.d_m_ext) {
      if (check) { if (sums) B3W_COMP(true, true, true); else B3W_COMP(true, true, false); }
      else { if (sums) B3W_COMP(false, true, true); else B3W_COMP(false, true, false); }
    } else {
      if (check) { if (sums) B3W_COMP(true, false, true); else B3W_COMP(true, false, false); }
      else { if (sums) B3W_COMP(false, false, true); else B3W_COMP(false, false, false); }
    }
#undef B3W_COMP
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->ctr_ev[pair], s));
  return B3W_OK;
}

static int check_device_args(const char *who, b3w_ctx *c, const void *d_in, const void *d_out) {
  if (!c || !d_in || !d_out) return fail(B3W_ERR_INVALID, "%s: null argument", who);
  if (((uintptr_t)d_out & 31) != 0) return fail(B3W_ERR_INVALID, "d_out must be 32-byte aligned");
  return B3W_OK;
}

extern "C" int b3w_witness_batch_device(b3w_ctx *c, const uint32_t *d_in, uint64_t n, uint8_t *d_out,
                                        uint8_t *d_status, uint32_t *d_pub, void *stream) {
  int rc = check_device_args("b3w_witness_batch_device", c, d_in, d_out);
  if (rc) return rc;
  ON_DEVICE(c);
  LOCKED(c);
  launch_opts o;
  o.check = (c->flags & B3W_FLAG_FUSED_CHECK) != 0;
  return launch_witness(c, d_in, n, d_out, d_status, d_pub, (cudaStream_t)stream, o);
}

extern "C" int b3w_witness_batch_device_checked(b3w_ctx *c, const uint32_t *d_in, uint64_t n, uint8_t *d_out,
                                                uint8_t *d_status, uint32_t *d_pub, uint32_t *d_first_bad, void *stream) {
  int rc = check_device_args("b3w_witness_batch_device_checked", c, d_in, d_out);
  if (rc) return rc;
  ON_DEVICE(c);
  LOCKED(c);
  launch_opts o;
  o.check = true;
  o.d_first_bad = d_first_bad;
  return launch_witness(c, d_in, n, d_out, d_status, d_pub, (cudaStream_t)stream, o);
}

extern "C" int b3w_witness_batch_device_ex(b3w_ctx *c, const uint32_t *d_in, const int8_t *d_m_ext, uint64_t n, uint8_t *d_out,
                                           uint8_t *d_status, uint32_t *d_pub, uint32_t *d_first_bad, uint64_t *d_sums, int check,
                                           void *stream) {
  int rc = check_device_args("b3w_witness_batch_device_ex", c, d_in, d_out);
  if (rc) return rc;
  if (d_sums && ((uintptr_t)d_sums & 7) != 0) return fail(B3W_ERR_INVALID, "d_sums must be 8-byte aligned");
  ON_DEVICE(c);
  LOCKED(c);
  launch_opts o;
  o.check = check != 0 || d_first_bad != nullptr || (c->flags & B3W_FLAG_FUSED_CHECK) != 0;
  o.d_first_bad = d_first_bad;
  o.d_m_ext = d_m_ext;
  o.d_sums = (unsigned long long *)d_sums;
  return launch_witness(c, d_in, n, d_out, d_status, d_pub, (cudaStream_t)stream, o);
}

extern "C" int b3w_r1cs_info(uint32_t circuit, uint32_t *n_rows, uint32_t *n_terms) {
  const circuit_def *d = find_def(circuit);
  if (!d) return B3W_ERR_UNSUPPORTED;
  if (n_rows) *n_rows = d->nova ? R1CS_ROWS_NOVA_BN_O1_SLOTS : R1CS_ROWS_COMPRESSION_SLOTS;
  if (n_terms) *n_terms = d->nova ? R1CS_TERMS_NOVA_BN_O1_SLOTS : R1CS_TERMS_COMPRESSION_SLOTS;
  return B3W_OK;
}

// instances [0, n) of d_wit; or -- listed -- the instances d_list[1 .. d_list[0]] (a device-side list: the count is read by
// the kernel), skipping those whose status says "Assert Failed." (no witness was written for them)
static int r1cs_check_launch(b3w_ctx *c, const uint8_t *d_wit, const uint32_t *d_list, uint64_t n, uint8_t *d_status, uint32_t *d_first_bad,
                             cudaStream_t s, bool listed = false, bool skip_asserted = false) {
  int rc = ensure_r1cs(c);
  if (rc) return rc;
  if (c->slot_rows == 0) return fail(B3W_ERR_UNSUPPORTED, "%s: no constraint system for the stand-alone check", c->def->name);
  if (n == 0) return B3W_OK;
  const r1cs_tables_dev T{c->r_slots.cls, c->r_slots.lo, c->r_slots.hi, c->r_slots.terms, c->r_slots.ncls, c->r_slots.coef_fr, c->r_slots.row_ids, c->r_slots.nblk};
  const uint32_t mw = ((c->def->ws + 31u) >> 5) + c->fp.n_vtiles + 1u;
  // the side table holds the non-bit slots of the circuit's witness layout (install_slot_rows: side_rank / side_total); a
  // witness with non-bit slots elsewhere (not one of this circuit's) is still checked, those values are re-read from HBM
  if (mw > FP_MAPW) return fail(B3W_ERR_UNSUPPORTED, "%s: %u map words (the checker holds %u)", c->def->name, mw, FP_MAPW);
  if (c->fp.side_total > 0xFFFFu) return fail(B3W_ERR_UNSUPPORTED, "%s: %u side-table entries", c->def->name, c->fp.side_total);
  const size_t smem = (size_t)FP_SIDE_OFF + (size_t)c->fp.side_total * 8;
  CK(cudaFuncSetAttribute(k_r1cs_check_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_sm = c->ctas_limit > 0 ? c->ctas_limit : FPK_CTAS_PER_SM;
  const uint64_t cap = (uint64_t)c->sm_count * per_sm;
  unsigned long long *d_counter = nullptr;                    // instances are handed out dynamically (kernels_r1cs_fast.cuh)
  uint32_t pair = 0;
  rc = take_counters(c, s, &d_counter, &pair);
  if (rc) return rc;
  k_r1cs_check_fast<<<(unsigned)(n < cap ? n : cap), FPK_THREADS, smem, s>>>(d_wit, listed ? d_list : nullptr, n, c->def->ws, c->fp,
                                                                             c->fp.n_vtiles ? c->fp0 : c->fp, T, c->d_field,
                                                                             d_status, d_first_bad, d_counter, (listed || skip_asserted) ? 1u : 0u);
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->ctr_ev[pair], s));
  return B3W_OK;
}

extern "C" int b3w_r1cs_check_device(b3w_ctx *c, const uint8_t *d_wit, uint64_t n, uint8_t *d_status,
                                     uint32_t *d_first_bad, void *stream) {
  if (!c || !d_wit) return fail(B3W_ERR_INVALID, "b3w_r1cs_check_device: null argument");
  if (((uintptr_t)d_wit & 31) != 0) return fail(B3W_ERR_INVALID, "d_wit must be 32-byte aligned");
  ON_DEVICE(c);
  LOCKED(c);
  return r1cs_check_launch(c, d_wit, nullptr, n, d_status, d_first_bad, (cudaStream_t)stream);
}

// how the loaded / built-in system was split: rows the compiled program covers, rows left to the general evaluator
extern "C" int b3w_r1cs_program_info(b3w_ctx *c, uint32_t *n_rows, uint32_t *n_compiled, uint32_t *n_xor_runs, uint32_t *n_tiles) {
  if (!c) return fail(B3W_ERR_INVALID, "b3w_r1cs_program_info: null argument");
  ON_DEVICE(c);
  LOCKED(c);
  int rc = ensure_r1cs(c);
  if (rc) return rc;
  if (n_rows) *n_rows = c->slot_rows;
  if (n_compiled) *n_compiled = c->fp.n_rows;
  if (n_xor_runs) *n_xor_runs = c->fp.n_xors;
  if (n_tiles) *n_tiles = c->fp.n_tiles;
  return B3W_OK;
}

// host-only: what fp_compile makes of a circuit's built-in system (needs no GPU; tests/test_abi.py pins the numbers)
static int b3w_r1cs_compile_stats_impl(uint32_t circuit, uint32_t *out, uint32_t n_out) {
  const circuit_def *d = find_def(circuit);
  if (!d) return B3W_ERR_UNSUPPORTED;
  std::vector<r1cs_load_detail::row> rows;
  if (!rows_from_builtin(d->r_slots, rows)) return fail(B3W_ERR_INVALID, "R1CS tables of %s are corrupt", d->name);
  fastprog_host fp, fp0;
  std::vector<char> taken;
  compile_programs(rows, d->ws, fp, fp0, taken);
  const uint32_t v[9] = {(uint32_t)rows.size(), fp.n_rows, (uint32_t)fp.xors.size(), (uint32_t)fp.tiles.size(), (uint32_t)fp.items.size(),
                         fp.n_virtual, fp.n_fast_tiles, (uint32_t)fp0.tiles.size(), (uint32_t)fp0.items.size()};
  for (uint32_t i = 0; i < n_out && i < 9; i++) out[i] = v[i];
  return B3W_OK;
}
extern "C" int b3w_r1cs_compile_stats(uint32_t circuit, uint32_t *n_rows, uint32_t *n_compiled, uint32_t *n_xor_runs, uint32_t *n_tiles, uint32_t *n_items) {
  uint32_t v[5] = {0, 0, 0, 0, 0};
  const int rc = guarded("b3w_r1cs_compile_stats", [&]() { return b3w_r1cs_compile_stats_impl(circuit, v, 5); });
  if (n_rows) *n_rows = v[0];
  if (n_compiled) *n_compiled = v[1];
  if (n_xor_runs) *n_xor_runs = v[2];
  if (n_tiles) *n_tiles = v[3];
  if (n_items) *n_items = v[4];
  return rc;
}
extern "C" int b3w_r1cs_compile_stats_ex(uint32_t circuit, uint32_t *out, uint32_t n_out) {
  if (!out) return fail(B3W_ERR_INVALID, "b3w_r1cs_compile_stats_ex: null argument");
  return guarded("b3w_r1cs_compile_stats_ex", [&]() { return b3w_r1cs_compile_stats_impl(circuit, out, n_out); });
}

// Replace the built-in slot-space row set of this context by the constraint system of an iden3 `.r1cs` file: the
// reference's own build/*.r1cs where the user has them, or the equivalents written by tools/export_r1cs.py.
static int b3w_r1cs_load_impl(b3w_ctx *c, const uint8_t *data, size_t len, uint32_t *n_rows) {
  if (!c || !data) return fail(B3W_ERR_INVALID, "b3w_r1cs_load: null argument");
  ON_DEVICE(c);
  LOCKED(c);
  r1cs_host_set hdr;
  std::vector<r1cs_load_detail::row> rows;
  std::string err;
  int rc = r1cs_parse_rows(data, len, c->def->prime, c->def->ws, rows, hdr, err);
  if (rc) return fail(rc, "b3w_r1cs_load: %s", err.c_str());
  const uint32_t m = (uint32_t)rows.size();
  CK(cudaDeviceSynchronize());                       // no check kernel may still be reading the tables that are replaced
  rc = install_slot_rows(c, rows, nullptr);
  if (rc) return rc;
  c->r1cs_loaded = true;
  if (n_rows) *n_rows = m;
  return B3W_OK;
}
extern "C" int b3w_r1cs_load(b3w_ctx *c, const uint8_t *data, size_t len, uint32_t *n_rows) {
  return guarded("b3w_r1cs_load", [&]() { return b3w_r1cs_load_impl(c, data, len, n_rows); });
}

static int b3w_r1cs_load_file_impl(b3w_ctx *c, const char *path, uint32_t *n_rows) {
  if (!c || !path) return fail(B3W_ERR_INVALID, "b3w_r1cs_load_file: null argument");
  FILE *f = fopen(path, "rb");
  if (!f) return fail(B3W_ERR_INVALID, "b3w_r1cs_load_file: cannot open %s", path);
  std::vector<uint8_t> buf;
  uint8_t tmp[1 << 16];
  size_t k;
  while ((k = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + k);
  fclose(f);
  return b3w_r1cs_load(c, buf.data(), buf.size(), n_rows);
}
extern "C" int b3w_r1cs_load_file(b3w_ctx *c, const char *path, uint32_t *n_rows) {
  return guarded("b3w_r1cs_load_file", [&]() { return b3w_r1cs_load_file_impl(c, path, n_rows); });
}

// host-only (no GPU, no context): compile the rows of an `.r1cs` file the way b3w_r1cs_load does and copy one table of the
// program out -- what tests/test_r1cs_program.py evaluates with Python integers against the file's own rows.
// section: 0 bool_mask, 1 bool_row, 2 xors, 3 xor_ids, 4 tiles, 5 virtual-bit groups, 6 items, 7 row_ids, 8 taken[] (u8 per row)
static int b3w_debug_r1cs_program_impl(const uint8_t *r1cs, size_t len, const uint8_t prime[32], uint32_t n_wires, int plain, uint32_t section,
                                       void *out, size_t cap, size_t *n_bytes) {
  if (!r1cs || !prime || !n_bytes) return fail(B3W_ERR_INVALID, "b3w_debug_r1cs_program: null argument");
  r1cs_host_set hdr;
  std::vector<r1cs_load_detail::row> rows;
  std::string err;
  int rc = r1cs_parse_rows(r1cs, len, prime, n_wires, rows, hdr, err);
  if (rc) return fail(rc, "b3w_debug_r1cs_program: %s", err.c_str());
  fastprog_host fp, fp0;
  std::vector<char> taken;
  compile_programs(rows, n_wires, fp, fp0, taken);
  const fastprog_host &P = (plain && fp.n_virtual) ? fp0 : fp;
  const void *src = nullptr;
  size_t nb = 0;
  switch (section) {
    case 0: src = P.bool_mask.data(); nb = P.bool_mask.size() * 4; break;
    case 1: src = P.bool_row.data(); nb = P.bool_row.size() * 4; break;
    case 2: src = P.xors.data(); nb = P.xors.size() * sizeof(fp_xor); break;
    case 3: src = P.xor_ids.data(); nb = P.xor_ids.size() * 4; break;
    case 4: src = P.tiles.data(); nb = P.tiles.size() * sizeof(fp_tile); break;
    case 5: src = P.vtiles.data(); nb = P.vtiles.size() * sizeof(fp_tile); break;
    case 6: src = P.items.data(); nb = P.items.size() * sizeof(fp_item); break;
    case 7: src = P.row_ids.data(); nb = P.row_ids.size() * 4; break;
    case 8: src = taken.data(); nb = taken.size(); break;
    default: return fail(B3W_ERR_INVALID, "b3w_debug_r1cs_program: section %u", section);
  }
  *n_bytes = nb;
  if (out && nb <= cap && nb) memcpy(out, src, nb);
  return B3W_OK;
}
extern "C" int b3w_debug_r1cs_program(const uint8_t *r1cs, size_t len, const uint8_t prime[32], uint32_t n_wires, int plain, uint32_t section,
                                      void *out, size_t cap, size_t *n_bytes) {
  return guarded("b3w_debug_r1cs_program", [&]() { return b3w_debug_r1cs_program_impl(r1cs, len, prime, n_wires, plain, section, out, cap, n_bytes); });
}

extern "C" int b3w_debug_inject_fault(b3w_ctx *c, uint32_t trace_word, uint32_t xor_mask) {
  if (!c) return fail(B3W_ERR_INVALID, "b3w_debug_inject_fault: null argument");
  if (trace_word != B3W_NO_ROW && trace_word >= (c->def->nova ? (uint32_t)NOVA_TRACE_WORDS : (uint32_t)TR_NOVA))
    return fail(B3W_ERR_INVALID, "trace word %u out of range", trace_word);
  c->fault_word = trace_word;
  c->fault_mask = xor_mask;
  return B3W_OK;
}

extern "C" int b3w_debug_set_launch(b3w_ctx *c, int ctas_per_sm, uint32_t parts) {
  if (!c || ctas_per_sm < 0 || parts > 65536) return fail(B3W_ERR_INVALID, "b3w_debug_set_launch: bad argument");
  c->ctas_limit = ctas_per_sm;
  c->sched_parts = parts;
  return B3W_OK;
}

extern "C" int b3w_debug_set_store_mode(b3w_ctx *c, int mode) {
  if (!c || mode < 0 || mode > 1) return fail(B3W_ERR_INVALID, "b3w_debug_set_store_mode: bad argument");
  LOCKED(c);
  c->store_mode = mode;
  return B3W_OK;
}

static int alloc_ring(b3w_ctx *c, uint32_t cap) {
  const circuit_def *d = c->def;
  c->ring_cap = cap;
  for (int k = 0; k < 2; k++) {
    CK(cudaStreamCreateWithFlags(&c->st[k], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev[k], cudaEventDisableTiming));
    CK(cudaEventCreate(&c->ev_k0[k]));
    CK(cudaEventCreate(&c->ev_k1[k]));
    // compressible unless the caller asked for ordinary memory; a driver without the virtual-memory entry points (or a
    // device that does not grant compression) silently gets ordinary memory -- one default in C, Python and the N-API addon
    if ((c->flags & B3W_FLAG_PLAIN_RING) ||
        device_alloc(c, (size_t)cap * d->ws * 32, B3W_MEM_COMPRESSIBLE, (void **)&c->d_ring[k]) != B3W_OK)
      CK(cudaMalloc(&c->d_ring[k], (size_t)cap * d->ws * 32));
    CK(cudaMalloc(&c->d_in[k], (size_t)cap * d->n_inputs * 4));
    CK(cudaMalloc(&c->d_status[k], (size_t)cap));
    CK(cudaMalloc(&c->d_pub[k], (size_t)cap * d->n_public * 4));
    CK(cudaMalloc(&c->d_sums[k], (size_t)cap * 8));
    CK(cudaMalloc(&c->d_fbad[k], (size_t)cap * 4));
    if (!d->nova) CK(cudaMalloc(&c->d_ext[k], (size_t)cap * 16));
  }
  return B3W_OK;
}
// the two ring slots exist completely or not at all (a half-built ring is released before the error is returned)
// Sized for the call at hand: min(chunk, n) instances per slot, rounded up to a power of two (at least 64), so that a
// context used for single witnesses (b3w_witness_one, the CLI) never maps gigabytes; a larger batch re-creates the ring
// (the slots are idle here: every host-buffer call drains its streams before it returns, under the context's lock).
static int ensure_ring(b3w_ctx *c, uint64_t n) {
  uint32_t need = 64;
  while (need < c->chunk && need < n) need <<= 1;
  if (need > c->chunk) need = c->chunk;
  if (c->ring_ready && c->ring_cap >= need) return B3W_OK;
  free_ring(c);
  const int rc = alloc_ring(c, need);
  if (rc) { free_ring(c); return rc; }
  c->ring_ready = true;
  return B3W_OK;
}
// staging of Fr256 input rows (b3w_witness_batch_fr): allocated on first use
static int ensure_fr_staging(b3w_ctx *c) {
  if (c->fr_ready) return B3W_OK;
  for (int k = 0; k < 2; k++) {
    CK(cudaMalloc(&c->d_fr[k], (size_t)c->ring_cap * c->def->n_inputs * 32));          // (free_ring releases these with the ring)
    if (c->def->nova) CK(cudaMalloc(&c->d_wlist[k], ((size_t)c->ring_cap + 1) * 4));
  }
  c->fr_ready = true;
  return B3W_OK;
}

// ---- per-call timing (b3w_last_timing) ---------------------------------------------------------------------------------
static void timing_begin(b3w_ctx *c) {
  memset(&c->timing, 0, sizeof c->timing);
  c->timing.total_ms = -now_ms();
}
// the kernel events of slot k have been recorded and their stream has drained: add the launch to the call's total
static void timing_collect(b3w_ctx *c, int k) {
  if (!c->ev_pending[k]) return;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, c->ev_k0[k], c->ev_k1[k]) == cudaSuccess) c->timing.kernel_ms += ms;
  c->ev_pending[k] = false;
}
static void timing_end(b3w_ctx *c) {
  for (int k = 0; k < 2; k++) timing_collect(c, k);
  c->timing.total_ms += now_ms();
}
extern "C" int b3w_last_timing(b3w_ctx *c, b3w_timing *out) {
  if (!c || !out) return fail(B3W_ERR_INVALID, "b3w_last_timing: null argument");
  LOCKED(c);
  *out = c->timing;
  return B3W_OK;
}

// ---- nova step circuits on field-element inputs (kernels_nova_wide.cuh) ---------------------------------------------
// trace words whose value is a function of a possibly field-valued input: the slots that read them through a W32 / W64 /
// S64 / INV descriptor are rewritten by the wide kernel's override pass
static bool nova_field_word(uint32_t t) {
  return (t >= NV_IN && t < NV_IN + 32) || t == NV_LDM1 || t == NV_DP1 || t == NV_DEPTH_OUT || (t >= NV_NEG_DEPTH && t < NV_BC_OUT + 2) ||
         (t >= NV_TMP_DOWN && t < NV_EQ_D + 128) || (t >= TR_IN + 8 && t < TR_IN + 24);
}
static int ensure_nova_wide(b3w_ctx *c) {
  if (c->nw_ready) return B3W_OK;
  const circuit_def *d = c->def;
  std::vector<uint2> lanes[32];
  for (uint32_t sl = 0; sl < d->ws; sl++) {
    const uint32_t dsc = c->h_desc[sl], kind = dsc >> 24, t = dsc & 0xFFFFu;
    if (kind >= DK_W32 && nova_field_word(t)) lanes[sl & 31].push_back(make_uint2(sl, dsc));
    else if (kind >= DK_S64) return fail(B3W_ERR_INVALID, "slot table of %s: field slot %u is not covered by the wide kernel", d->name, sl);
    // Num2Bits(65).out[64]: constant 0 for u32 inputs; present in the O1 build only, right after out[63]
    if (dsc == ((DK_BIT << 24) | (31u << 16) | (NV_IN + 11)) && sl + 1 < d->ws && c->h_desc[sl + 1] == ((DK_W32 << 24) | TR_ZERO))
      lanes[(sl + 1) & 31].push_back(make_uint2(sl + 1, DK_WIDE_BIT64 << 24));
  }
  std::vector<uint2> all;
  uint32_t off[33];
  for (int l = 0; l < 32; l++) {
    off[l] = (uint32_t)all.size();
    all.insert(all.end(), lanes[l].begin(), lanes[l].end());
  }
  off[32] = (uint32_t)all.size();
  for (void **q : {(void **)&c->d_wslots, (void **)&c->d_lane_off})       // a failed earlier attempt
    if (*q) { cudaFree(*q); *q = nullptr; }
  CK(cudaMalloc(&c->d_wslots, all.size() * sizeof(uint2) + 16));
  CK(cudaMemcpy(c->d_wslots, all.data(), all.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&c->d_lane_off, sizeof off));
  CK(cudaMemcpy(c->d_lane_off, off, sizeof off, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(k_blake3_nova_witness_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, NW_WARPS * NW_STRIDE * 4));
  c->nw_ready = true;
  return B3W_OK;
}

// What one host-buffer batch call is made of.  Exactly one of `in` (u32 rows) / `in_fr` (Fr256 rows) is set.
struct batch_job {
  const uint32_t *in = nullptr;
  const uint8_t *in_fr = nullptr;
  const int8_t *m_ext = nullptr;      // with `in`, compression only: the wide-domain kernel
  uint64_t n = 0;
  uint8_t *out = nullptr;
  uint8_t *status = nullptr;
  uint32_t *pub = nullptr;
  b3w_batch_extras ex = {};
};

// Host-buffer batches: chunks of c->chunk instances through the two ring slots; chunk j runs on stream j & 1 and its D2H
// overlaps the next chunk's kernel.
static int batch_chunks(b3w_ctx *c, const batch_job &J) {
  const circuit_def *d = c->def;
  const size_t wbytes = (size_t)d->ws * 32;
  const bool check = (c->flags & B3W_FLAG_FUSED_CHECK) != 0, byte_check = (c->flags & B3W_FLAG_BYTE_CHECK) != 0;
  // the sample instances in index order: every chunk copies the ones it holds out of its ring slot
  std::vector<std::pair<uint64_t, uint32_t>> samples;
  for (uint32_t j = 0; j < J.ex.n_samples; j++) samples.push_back({J.ex.sample_idx[j], j});
  std::sort(samples.begin(), samples.end());
  size_t sp = 0;
  uint64_t done = 0;
  int k = 0;
  while (done < J.n) {
    const uint64_t m = J.n - done < c->ring_cap ? J.n - done : c->ring_cap;
    cudaStream_t s = c->st[k];
    launch_opts o;
    o.check = check;
    o.d_first_bad = ((check || byte_check) && J.ex.first_bad) ? c->d_fbad[k] : nullptr;
    o.d_sums = J.ex.sums ? (unsigned long long *)c->d_sums[k] : nullptr;
    {
      nvtx_range r("b3w:h2d");
      if (J.in_fr) {
        const size_t fb = (size_t)m * d->n_inputs * 32;
        CK(cudaMemcpyAsync(c->d_fr[k], J.in_fr + done * d->n_inputs * 32, fb, cudaMemcpyHostToDevice, s));
        c->timing.h2d_bytes += fb;
        const unsigned grid = (unsigned)std::min<uint64_t>((m + 7) / 8, (uint64_t)c->sm_count * 8);
        if (d->nova) {
          CK(cudaMemsetAsync(c->d_wlist[k], 0, 4, s));
          k_fr_to_rows_nova<<<grid, 256, 0, s>>>(c->d_fr[k], m, c->d_field, c->d_in[k], c->d_wlist[k]);
        } else {
          k_fr_to_rows_compression<<<grid, 256, 0, s>>>(c->d_fr[k], m, c->d_field, c->d_in[k], c->d_ext[k]);
          o.d_m_ext = c->d_ext[k];
        }
        CK(cudaGetLastError());
      } else {
        CK(cudaMemcpyAsync(c->d_in[k], J.in + done * d->n_inputs, (size_t)m * d->n_inputs * 4, cudaMemcpyHostToDevice, s));
        c->timing.h2d_bytes += (size_t)m * d->n_inputs * 4;
        if (J.m_ext) {
          CK(cudaMemcpyAsync(c->d_ext[k], J.m_ext + done * 16, (size_t)m * 16, cudaMemcpyHostToDevice, s));
          c->timing.h2d_bytes += (size_t)m * 16;
          o.d_m_ext = c->d_ext[k];
        }
      }
    }
    {
      nvtx_range r("b3w:kernel");
      CK(cudaEventRecord(c->ev_k0[k], s));
      int rc = launch_witness(c, c->d_in[k], m, c->d_ring[k], c->d_status[k], c->d_pub[k], s, o);
      if (rc) return rc;
      if (J.in_fr && d->nova) {
        // the instances of the wide list (field-valued inputs): the general kernel fills their places; with the fused-check
        // flag their witnesses are then checked where they lie (the stand-alone evaluator, on the listed instances only)
        const nova_wide_args wa{c->d_fr[k], c->d_wslots, c->d_lane_off, c->d_field};
        k_blake3_nova_witness_wide<<<c->sm_count * 4, NW_WARPS * 32, NW_WARPS * NW_STRIDE * 4, s>>>(
            wa, c->d_wlist[k], m, c->d_desc, d->ws, c->d_ring[k], c->d_status[k], c->d_pub[k], o.d_sums);
        CK(cudaGetLastError());
        if (check) {
          rc = r1cs_check_launch(c, c->d_ring[k], c->d_wlist[k], m, c->d_status[k], o.d_first_bad, s, /*listed=*/true);
          if (rc) return rc;
        }
      }
      if (byte_check) {
        // every witness of the chunk is read back from the ring slot and ALL rows of the constraint system are evaluated on
        // its bytes (what a consumer does with the vector it is handed, rust_fold/src/utils.rs:78-85): a wrong slot
        // descriptor, a wrong split or a lost store cannot pass, which the fused check -- it sees the trace -- cannot say.
        // (Making chunk k + 1's generator run beside this checker, with fewer checker CTAs per SM so that both are resident,
        // was measured: 5.65 -> 5.85 M witnesses/s at best, slower for 4 096-instance slots; profiles/r02z_byte_check_overlap_sweep.jsonl)
        rc = r1cs_check_launch(c, c->d_ring[k], nullptr, m, c->d_status[k], o.d_first_bad, s, false, /*skip_asserted=*/true);
        if (rc) return rc;
        c->timing.launches++;
      }
      CK(cudaEventRecord(c->ev_k1[k], s));
      c->ev_pending[k] = true;
      c->timing.launches++;
    }
    {
      nvtx_range r("b3w:d2h");
      size_t bytes = 0;
      if (J.out) { CK(cudaMemcpyAsync(J.out + done * wbytes, c->d_ring[k], (size_t)m * wbytes, cudaMemcpyDeviceToHost, s)); bytes += (size_t)m * wbytes; }
      if (J.status) { CK(cudaMemcpyAsync(J.status + done, c->d_status[k], (size_t)m, cudaMemcpyDeviceToHost, s)); bytes += m; }
      if (J.pub) { CK(cudaMemcpyAsync(J.pub + done * d->n_public, c->d_pub[k], (size_t)m * d->n_public * 4, cudaMemcpyDeviceToHost, s)); bytes += (size_t)m * d->n_public * 4; }
      if (J.ex.sums) { CK(cudaMemcpyAsync(J.ex.sums + done, c->d_sums[k], (size_t)m * 8, cudaMemcpyDeviceToHost, s)); bytes += (size_t)m * 8; }
      if (o.d_first_bad) { CK(cudaMemcpyAsync(J.ex.first_bad + done, c->d_fbad[k], (size_t)m * 4, cudaMemcpyDeviceToHost, s)); bytes += (size_t)m * 4; }
      for (; sp < samples.size() && samples[sp].first < done + m; sp++) {
        CK(cudaMemcpyAsync(J.ex.sample_out + (size_t)samples[sp].second * wbytes, c->d_ring[k] + (samples[sp].first - done) * wbytes, wbytes,
                           cudaMemcpyDeviceToHost, s));
        bytes += wbytes;
      }
      c->timing.d2h_bytes += bytes;
    }
    done += m;
    k ^= 1;
    // before reusing slot k (two chunks ago) its stream must have drained
    if (done < J.n) {
      CK(cudaStreamSynchronize(c->st[k]));
      timing_collect(c, k);
    }
  }
  return B3W_OK;
}

static int batch_host(b3w_ctx *c, const batch_job &J, const char *who) {
  const b3w_batch_extras &ex = J.ex;
  if (ex.n_samples > B3W_MAX_SAMPLES) return fail(B3W_ERR_INVALID, "%s: %u samples (at most %u)", who, ex.n_samples, B3W_MAX_SAMPLES);
  if (ex.n_samples && (!ex.sample_idx || !ex.sample_out)) return fail(B3W_ERR_INVALID, "%s: samples without sample_idx / sample_out", who);
  for (uint32_t j = 0; j < ex.n_samples; j++)
    if (ex.sample_idx[j] >= J.n) return fail(B3W_ERR_INVALID, "%s: sample %u = instance %llu of %llu", who, j, (unsigned long long)ex.sample_idx[j], (unsigned long long)J.n);
  ON_DEVICE(c);
  LOCKED(c);
  nvtx_range r(who);
  timing_begin(c);
  c->timing.instances = J.n;
  int rc = ensure_ring(c, J.n);
  if (rc == B3W_OK && J.in_fr) rc = ensure_fr_staging(c);
  if (rc == B3W_OK && J.in_fr && c->def->nova) rc = ensure_nova_wide(c);
  if (rc) return rc;
  if (ex.first_bad && !(c->flags & (B3W_FLAG_FUSED_CHECK | B3W_FLAG_BYTE_CHECK)))
    for (uint64_t i = 0; i < J.n; i++) ex.first_bad[i] = B3W_NO_ROW;          // no check ran
  rc = guarded(who, [&]() { return batch_chunks(c, J); });
  // drain both slots on every path: after an error no copy into the caller's buffers may still be in flight
  const cudaError_t e0 = cudaStreamSynchronize(c->st[0]), e1 = cudaStreamSynchronize(c->st[1]);
  timing_end(c);
  if (rc) return rc;
  if (e0 != cudaSuccess || e1 != cudaSuccess)
    return fail(B3W_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
  return B3W_OK;
}

extern "C" int b3w_witness_batch(b3w_ctx *c, const uint32_t *in, uint64_t n, uint8_t *out, uint8_t *status,
                                 uint32_t *pub) {
  if (!c || (!in && n)) return fail(B3W_ERR_INVALID, "b3w_witness_batch: null argument");
  batch_job J;
  J.in = in; J.n = n; J.out = out; J.status = status; J.pub = pub;
  return batch_host(c, J, "b3w_witness_batch");
}

extern "C" int b3w_witness_batch_ex(b3w_ctx *c, const uint32_t *in, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub,
                                    const b3w_batch_extras *extras) {
  if (!c || (!in && n)) return fail(B3W_ERR_INVALID, "b3w_witness_batch_ex: null argument");
  batch_job J;
  J.in = in; J.n = n; J.out = out; J.status = status; J.pub = pub;
  if (extras) J.ex = *extras;
  return batch_host(c, J, "b3w_witness_batch_ex");
}

// Inputs as field elements (canonical or not): what `normalize` (witness_calculator.js:319-323) leaves is value mod p;
// the u32 row entry points cover the circuits' honest domain [0, 2^32), anything else is refused with B3W_ERR_DOMAIN.
static bool ge256(const uint32_t a[8], const uint32_t b[8]) {
  for (int i = 7; i >= 0; i--)
    if (a[i] != b[i]) return a[i] > b[i];
  return true;
}
extern "C" int b3w_inputs_from_fr(uint32_t circuit, const uint8_t *in_fr, uint64_t n, uint32_t *rows) {
  const circuit_def *d = find_def(circuit);
  if (!d) return B3W_ERR_UNSUPPORTED;
  if ((!in_fr || !rows) && n) return fail(B3W_ERR_INVALID, "b3w_inputs_from_fr: null argument");
  uint32_t p[8];
  memcpy(p, d->prime, 32);
  for (uint64_t i = 0; i < n; i++)
    for (uint32_t k = 0; k < d->n_inputs; k++) {
      uint32_t v[8];
      memcpy(v, in_fr + (i * d->n_inputs + k) * 32, 32);
      while (ge256(v, p)) {                                  // value mod p (p > 2^253: at most a few rounds)
        uint64_t br = 0;
        for (int j = 0; j < 8; j++) {
          uint64_t t = (uint64_t)v[j] - p[j] - br;
          v[j] = (uint32_t)t;
          br = (t >> 32) & 1;
        }
      }
      if (v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7]) {
        const char *nm = "?";
        uint32_t idx = k;
        for (int sg = 0; sg < d->n_sig; sg++)
          if (k >= d->sig[sg].off && k < d->sig[sg].off + d->sig[sg].size) { nm = d->sig[sg].name; idx = k - d->sig[sg].off; }
        return fail(B3W_ERR_DOMAIN, "instance %llu: input %s[%u] is outside the supported u32 domain", (unsigned long long)i, nm, idx);
      }
      rows[i * d->n_inputs + k] = v[0];
    }
  return B3W_OK;
}

static fr_t prime_of(const circuit_def *d) {
  fr_t p;
  memcpy(p.l, d->prime, 32);
  return p;
}

// The full input domain of blake3_compression (wide_domain.h): Fr256 inputs -> u32 rows + the signed high parts of the
// message words.  Nothing is refused: an instance that cannot satisfy the circuit is marked (m_ext[i][0] = 127) and comes
// back from the kernels with status 4, like the reference's "Assert Failed.".  (Host form of k_fr_to_rows_compression, for
// callers that keep u32 rows; b3w_witness_batch_fr converts on the device.)
extern "C" int b3w_inputs_from_fr_wide(uint32_t circuit, const uint8_t *in_fr, uint64_t n, uint32_t *rows, int8_t *m_ext,
                                       uint64_t *n_wide) {
  const circuit_def *d = find_def(circuit);
  if (!d) return B3W_ERR_UNSUPPORTED;
  if (d->nova) return fail(B3W_ERR_UNSUPPORTED, "%s: only blake3_compression has the u32 + m_ext row form (use b3w_witness_batch_fr)", d->name);
  if ((!in_fr || !rows || !m_ext) && n) return fail(B3W_ERR_INVALID, "b3w_inputs_from_fr_wide: null argument");
  const fr_t p = prime_of(d);
  uint64_t wide = 0;
  for (uint64_t i = 0; i < n; i++) {
    fr_t v[28];
    for (int k = 0; k < 28; k++) v[k] = wd_load_reduced(in_fr + (i * 28 + k) * 32, p);
    if (wd_convert_compression(v, p, rows + i * 28, m_ext + i * 16)) wide++;
  }
  if (n_wide) *n_wide = wide;
  return B3W_OK;
}

extern "C" int b3w_witness_batch_wide(b3w_ctx *c, const uint32_t *in, const int8_t *m_ext, uint64_t n, uint8_t *out, uint8_t *status,
                                      uint32_t *pub) {
  if (!c || ((!in || !m_ext) && n)) return fail(B3W_ERR_INVALID, "b3w_witness_batch_wide: null argument");
  if (c->def->nova) return fail(B3W_ERR_UNSUPPORTED, "%s: only blake3_compression has a wide-domain kernel", c->def->name);
  batch_job J;
  J.in = in; J.m_ext = m_ext; J.n = n; J.out = out; J.status = status; J.pub = pub;
  return batch_host(c, J, "b3w_witness_batch_wide");
}

extern "C" int b3w_witness_batch_device_wide(b3w_ctx *c, const uint32_t *d_in, const int8_t *d_m_ext, uint64_t n, uint8_t *d_out,
                                             uint8_t *d_status, uint32_t *d_pub, uint32_t *d_first_bad, void *stream) {
  if (!d_m_ext) return fail(B3W_ERR_INVALID, "b3w_witness_batch_device_wide: null argument");
  return b3w_witness_batch_device_ex(c, d_in, d_m_ext, n, d_out, d_status, d_pub, d_first_bad, nullptr, 0, stream);
}

// Fr256 rows, any field elements, HOST buffers: the rows are copied to the device as they are and converted there
// (kernels_fr_input.cuh); u32 instances run on the hot kernels, the others on the wide paths, into the same outputs.
extern "C" int b3w_witness_batch_fr_ex(b3w_ctx *c, const uint8_t *in_fr, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub,
                                       const b3w_batch_extras *extras) {
  if (!c || (!in_fr && n)) return fail(B3W_ERR_INVALID, "b3w_witness_batch_fr: null argument");
  batch_job J;
  J.in_fr = in_fr; J.n = n; J.out = out; J.status = status; J.pub = pub;
  if (extras) J.ex = *extras;
  return batch_host(c, J, "b3w_witness_batch_fr");
}
extern "C" int b3w_witness_batch_fr(b3w_ctx *c, const uint8_t *in_fr, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub) {
  return b3w_witness_batch_fr_ex(c, in_fr, n, out, status, pub, nullptr);
}

// b3w_assert_trace for field-element inputs, any values: the circuits' range constraints are replayed in the wasm's
// execution order (wide_domain.h).
extern "C" int b3w_assert_trace_fr(uint32_t circuit, const uint8_t *in_fr, char *buf, size_t cap) {
  const circuit_def *d = find_def(circuit);
  if (!d) return B3W_ERR_UNSUPPORTED;
  if (!in_fr || (!buf && cap)) return fail(B3W_ERR_INVALID, "b3w_assert_trace_fr: null argument");
  if (cap) buf[0] = 0;
  const fr_t p = prime_of(d);
  if (d->nova) {
    fr_t v[32];
    for (int k = 0; k < 32; k++) v[k] = wd_load_reduced(in_fr + k * 32, p);
    return wd_assert_text_nova(v, p, buf, cap) ? B3W_CIRCOM_ASSERT : B3W_OK;
  }
  fr_t v[28];
  for (int k = 0; k < 28; k++) v[k] = wd_load_reduced(in_fr + k * 32, p);
  const wd_fail f = wd_first_assert_compression(v, p);
  if (f.kind == 0) return B3W_OK;
  wd_assert_text_compression(f, buf, cap);
  return B3W_CIRCOM_ASSERT;
}

extern "C" int b3w_witness_one(b3w_ctx *c, const uint32_t *in, uint8_t *out) {
  if (!c || !in || !out) return fail(B3W_ERR_INVALID, "b3w_witness_one: null argument");
  uint8_t status = 0;
  int rc = b3w_witness_batch(c, in, 1, out, &status, nullptr);
  if (rc) return rc;
  if (status) return fail((int)status, "Assert Failed.");
  return B3W_OK;
}

extern "C" int b3w_checksum_device(b3w_ctx *c, const uint8_t *d_wit, uint64_t n, uint64_t *d_sums, void *stream) {
  if (!c || !d_wit || !d_sums) return fail(B3W_ERR_INVALID, "b3w_checksum_device: null argument");
  ON_DEVICE(c);
  LOCKED(c);
  if (n == 0) return B3W_OK;
  uint64_t ctas = (n + 7) / 8, cap = (uint64_t)c->sm_count * 8;
  k_checksum<<<(unsigned)(ctas < cap ? ctas : cap), 256, 0, (cudaStream_t)stream>>>((const uint64_t *)d_wit, n, c->def->ws, d_sums);
  CK(cudaGetLastError());
  return B3W_OK;
}

extern "C" int b3w_calib_fill_items(b3w_ctx *c, uint8_t *d_buf, uint64_t bytes, void *stream) {
  if (!c || !d_buf) return fail(B3W_ERR_INVALID, "b3w_calib_fill_items: null argument");
  if (((uintptr_t)d_buf & 31) != 0) return fail(B3W_ERR_INVALID, "d_buf must be 32-byte aligned");
  ON_DEVICE(c);
  LOCKED(c);
  const uint32_t item_slots = c->sched_parts >= 32 ? c->sched_parts / 32 * 32 : 1024;   // default 32 KiB = the witness kernels' work item (tuning hook: set_launch parts >= 32 = slots per item)
  const uint64_t n_items = bytes / (item_slots * 32ull);
  if (n_items == 0) return B3W_OK;
  sched_args sc;
  uint32_t pair;
  int rc = take_counters(c, (cudaStream_t)stream, &sc.counter, &pair);
  if (rc) return rc;
  sc.parts = 1;
  sc.part_len = item_slots;
  const int per_sm = c->ctas_limit > 0 ? c->ctas_limit : 2;
  k_fill_items<<<c->sm_count * per_sm, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream>>>(d_buf, n_items, item_slots, sc);
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->ctr_ev[pair], (cudaStream_t)stream));
  return B3W_OK;
}

extern "C" int b3w_calib_fill_bulk(b3w_ctx *c, uint8_t *d_buf, uint64_t bytes, void *stream) {
  if (!c || !d_buf) return fail(B3W_ERR_INVALID, "b3w_calib_fill_bulk: null argument");
  if (((uintptr_t)d_buf & 127) != 0) return fail(B3W_ERR_INVALID, "d_buf must be 128-byte aligned");
  ON_DEVICE(c);
  LOCKED(c);
  const uint32_t item_bytes = (c->sched_parts >= 32 ? c->sched_parts / 32 * 32 : 1024) * 32;     // tuning hook as in b3w_calib_fill_items
  const uint64_t n_items = bytes / item_bytes;
  if (n_items == 0) return B3W_OK;
  sched_args sc;
  uint32_t pair;
  int rc = take_counters(c, (cudaStream_t)stream, &sc.counter, &pair);
  if (rc) return rc;
  sc.parts = 1;
  sc.part_len = item_bytes / 32;
  CK(cudaFuncSetAttribute(k_fill_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)item_bytes));
  const int per_sm = c->ctas_limit > 0 ? c->ctas_limit : 2;
  k_fill_bulk<<<c->sm_count * per_sm, 128, item_bytes, (cudaStream_t)stream>>>(d_buf, n_items, item_bytes, sc);
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->ctr_ev[pair], (cudaStream_t)stream));
  return B3W_OK;
}

extern "C" int b3w_calib_fill(b3w_ctx *c, uint8_t *d_buf, uint64_t bytes, void *stream) {
  if (!c || !d_buf) return fail(B3W_ERR_INVALID, "b3w_calib_fill: null argument");
  ON_DEVICE(c);
  LOCKED(c);
  k_fill<<<c->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(d_buf, bytes / 32);
  CK(cudaGetLastError());
  return B3W_OK;
}


// ------------------------------------------------------------------------------------------------
// chained-chunk driver, host side
// ------------------------------------------------------------------------------------------------
static uint64_t chunk_count_of(uint64_t len) { return len == 0 ? 1 : (len + 1023) / 1024; }
static uint32_t blocks_of_chunk(uint64_t len, uint64_t c) {
  uint64_t cb = len - c * 1024 < 1024 ? len - c * 1024 : 1024;
  uint32_t nb = (uint32_t)((cb + 63) / 64);
  return nb ? nb : 1;
}

struct tree_plan {
  std::vector<uint32_t> nodes;        // parents in creation (post-) order: {left ref, right ref}; ref < n_chunks = chunk
  std::vector<uint32_t> height;       // per parent
  std::vector<uint32_t> depth_of;     // per chunk: number of parents above it
  std::vector<uint32_t> path;         // per chunk: sibling refs, root first, max_depth entries
  uint32_t max_depth = 1;
  uint32_t root = 0;
};
// BLAKE3 tree shape: the left subtree takes the largest power of two of chunks that leaves >= 1 on the right
static uint32_t build_tree(tree_plan &t, uint64_t first, uint64_t n, uint64_t n_chunks, uint32_t &h_out) {
  if (n == 1) { h_out = 0; return (uint32_t)first; }
  uint64_t left = 1;
  while (left * 2 < n) left *= 2;
  uint32_t hl, hr;
  uint32_t l = build_tree(t, first, left, n_chunks, hl), r = build_tree(t, first + left, n - left, n_chunks, hr);
  t.nodes.push_back(l);
  t.nodes.push_back(r);
  h_out = (hl > hr ? hl : hr) + 1;
  t.height.push_back(h_out);
  return (uint32_t)(n_chunks + t.height.size() - 1);
}
// Sibling of every parent on the way down to each chunk.  stack = the {left, right} children of those parents, root first.
//   true siblings (default): the child the path does NOT descend into -- what a bao slice proves against the BLAKE3 root;
//   reference_siblings: the reference's rule (rust_fold/src/blake3_hash.rs:60-78): for parent i of par_len the direction is
//     read off the chunk index, `leaf & (1 << (par_len - i - 1)) == 0` = Left, and the sibling is the parent's OTHER half by
//     that direction.  The two agree on every perfect (2^k-chunk) tree and wherever the bits of the chunk index spell the real
//     path; for e.g. chunk 4 of 5 the reference picks the chunk's own subtree as its "sibling".
static void fill_paths(tree_plan &t, uint32_t ref, uint64_t n_chunks, std::vector<std::pair<uint32_t, uint32_t>> &stack, std::vector<uint8_t> &went_left,
                       bool reference_siblings) {
  if (ref < n_chunks) {
    const size_t par_len = stack.size();
    t.depth_of[ref] = (uint32_t)par_len;
    for (size_t i = 0; i < par_len; i++) {
      bool left = went_left[i] != 0;
      if (reference_siblings) left = par_len - i - 1 >= 64 ? true : (((uint64_t)ref >> (par_len - i - 1)) & 1) == 0;
      t.path[(size_t)ref * t.max_depth + i] = left ? stack[i].second : stack[i].first;
    }
    return;
  }
  uint32_t j = ref - (uint32_t)n_chunks, l = t.nodes[2 * j], r = t.nodes[2 * j + 1];
  stack.push_back({l, r});
  went_left.push_back(1); fill_paths(t, l, n_chunks, stack, went_left, reference_siblings); went_left.pop_back();
  went_left.push_back(0); fill_paths(t, r, n_chunks, stack, went_left, reference_siblings); went_left.pop_back();
  stack.pop_back();
}

// Everything the host decides about a file before the device starts: the BLAKE3 tree (parents ordered by height so that
// one launch hashes one level), every chunk's sibling path and its first step.
struct chain_plan {
  uint64_t nc = 0, total = 0;
  size_t n_par = 0;
  uint32_t max_depth = 1;
  std::vector<uint32_t> nodes;                 // level-ordered parents: {left ref, right ref}
  std::vector<std::pair<uint32_t, uint32_t>> levels;   // {first parent, count} per tree level, bottom up
  std::vector<uint32_t> path, depth_of;
  std::vector<uint64_t> step_off;              // nc + 1
};
static int make_chain_plan(uint64_t len, chain_plan &p, bool reference_siblings = false) {
  const uint64_t nc = chunk_count_of(len);
  if (nc > 0x7FFFFFFFull) return fail(B3W_ERR_INVALID, "input too large for the chain driver (%llu chunks)", (unsigned long long)nc);
  tree_plan t;
  uint32_t hgt;
  t.root = build_tree(t, 0, nc, nc, hgt);
  t.max_depth = hgt ? hgt : 1;
  t.depth_of.assign(nc, 0);
  t.path.assign((size_t)nc * t.max_depth, 0);
  {
    std::vector<std::pair<uint32_t, uint32_t>> st;
    std::vector<uint8_t> wl;
    fill_paths(t, t.root, nc, st, wl, reference_siblings);
  }
  p.nc = nc;
  p.max_depth = t.max_depth;
  p.n_par = t.height.size();
  p.step_off.assign(nc + 1, 0);
  for (uint64_t k = 0; k < nc; k++) p.step_off[k + 1] = p.step_off[k] + blocks_of_chunk(len, k) + t.depth_of[k];
  p.total = p.step_off[nc];
  std::vector<uint32_t> order(p.n_par), remap(p.n_par);
  for (size_t i = 0; i < p.n_par; i++) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return t.height[a] < t.height[b]; });
  for (size_t i = 0; i < p.n_par; i++) remap[order[i]] = (uint32_t)i;
  auto fix = [&](uint32_t ref) { return ref < nc ? ref : (uint32_t)(nc + remap[ref - nc]); };
  p.nodes.assign(2 * p.n_par + 2, 0);
  for (size_t i = 0; i < p.n_par; i++) { p.nodes[2 * i] = fix(t.nodes[2 * order[i]]); p.nodes[2 * i + 1] = fix(t.nodes[2 * order[i] + 1]); }
  for (auto &r : t.path) r = fix(r);
  for (size_t i = 0; i < p.n_par;) {
    size_t j = i;
    while (j < p.n_par && t.height[order[j]] == t.height[order[i]]) j++;
    p.levels.push_back({(uint32_t)i, (uint32_t)(j - i)});
    i = j;
  }
  p.path.swap(t.path);
  p.depth_of.swap(t.depth_of);
  return B3W_OK;
}

static int b3w_nova_chain_size_impl(uint64_t len, uint64_t *n_chunks, uint64_t *total_steps) {
  chain_plan p;
  int rc = make_chain_plan(len, p);
  if (rc) return rc;
  if (n_chunks) *n_chunks = p.nc;
  if (total_steps) *total_steps = p.total;
  return B3W_OK;
}
extern "C" int b3w_nova_chain_size(uint64_t len, uint64_t *n_chunks, uint64_t *total_steps) {
  return guarded("b3w_nova_chain_size", [&]() { return b3w_nova_chain_size_impl(len, n_chunks, total_steps); });
}

// grow-only device scratch of the chain driver, kept in the context (cudaMalloc / cudaFree per call cost more than the
// whole 1 MiB job)
enum { CS_DATA, CS_CV, CS_NODES, CS_PATH, CS_DEPTH, CS_OFF, CS_ROWS, CS_ROOT, CS_SLOTS };
static int chain_scratch(b3w_ctx *c, int slot, size_t need, void **out) {
  if (c->cs_cap[slot] < need) {
    if (c->cs_ptr[slot]) cudaFree(c->cs_ptr[slot]);
    c->cs_ptr[slot] = nullptr;
    c->cs_cap[slot] = 0;
    size_t cap = need + need / 4 + 256;
    CK(cudaMalloc(&c->cs_ptr[slot], cap));
    c->cs_cap[slot] = cap;
  }
  *out = c->cs_ptr[slot];
  return B3W_OK;
}

// Where the step witnesses of a chain go.  Host form: through the ring into the caller's host arrays (any may be NULL).
// Device form: straight into caller-supplied DEVICE buffers (d_out etc.), nothing crosses PCIe but the file itself.
struct chain_sink {
  bool device = false;
  uint8_t *out = nullptr;
  uint8_t *status = nullptr;
  uint32_t *pub = nullptr;
  uint32_t *rows = nullptr;
  uint8_t *root = nullptr;          // host, 32 bytes, in both forms
};

// Steps of chunks [lo, hi) of the file: this device hashes the whole tree (cheap), builds the rows of its chunks and
// generates their step witnesses; outputs are the caller's FULL arrays, written at this range's offsets.
static int nova_chain_steps(b3w_ctx *c, const chain_plan &p, const uint8_t *data, uint64_t len, uint64_t lo, uint64_t hi, const chain_sink &K) {
  const uint64_t nc = p.nc, first = p.step_off[lo], total = p.step_off[hi] - first;
  int rc = ensure_ring(c, total);
  if (rc) return rc;
  uint8_t *d_data; uint32_t *d_cv, *d_nodes, *d_path, *d_depth, *d_rows, *d_root; uint64_t *d_off;
  const size_t padded = (size_t)nc * 1024;
  if ((rc = chain_scratch(c, CS_DATA, padded, (void **)&d_data))) return rc;
  if ((rc = chain_scratch(c, CS_CV, (nc + p.n_par) * 32, (void **)&d_cv))) return rc;
  if ((rc = chain_scratch(c, CS_NODES, p.nodes.size() * 4, (void **)&d_nodes))) return rc;
  if ((rc = chain_scratch(c, CS_PATH, p.path.size() * 4, (void **)&d_path))) return rc;
  if ((rc = chain_scratch(c, CS_DEPTH, nc * 4, (void **)&d_depth))) return rc;
  if ((rc = chain_scratch(c, CS_OFF, (nc + 1) * 8, (void **)&d_off))) return rc;
  if ((rc = chain_scratch(c, CS_ROOT, 32, (void **)&d_root))) return rc;
  if (K.device && K.rows) d_rows = K.rows + first * 32;                          // the caller's device array IS the row buffer
  else if ((rc = chain_scratch(c, CS_ROWS, (size_t)total * 128, (void **)&d_rows))) return rc;
  cudaStream_t s0 = c->st[0];
  {
    nvtx_range r("b3w:h2d");
    // the tail of the last chunk is zero padding (load_block reads whole 64-byte blocks)
    const size_t tail = len & ~(size_t)1023;
    CK(cudaMemsetAsync(d_data + tail, 0, padded - tail, s0));
    if (len) CK(cudaMemcpyAsync(d_data, data, len, cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(d_nodes, p.nodes.data(), p.nodes.size() * 4, cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(d_path, p.path.data(), p.path.size() * 4, cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(d_depth, p.depth_of.data(), nc * 4, cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(d_off, p.step_off.data(), (nc + 1) * 8, cudaMemcpyHostToDevice, s0));
    c->timing.h2d_bytes += len + p.nodes.size() * 4 + p.path.size() * 4 + nc * 4 + (nc + 1) * 8;
  }
  {
    nvtx_range r("b3w:tree");
    k_chunk_cvs<<<(unsigned)((nc + 127) / 128), 128, 0, s0>>>(d_data, len, nc, d_cv);
    for (const auto &lv : p.levels)                                   // one launch per tree level
      k_parent_cvs<<<(lv.second + 127) / 128, 128, 0, s0>>>(d_nodes, lv.first, lv.second, nc, d_cv);
    k_chain_rows<<<(unsigned)((hi - lo + 63) / 64), 64, 0, s0>>>(d_data, len, lo, hi, d_cv, d_path, d_depth, p.max_depth, d_off, d_rows, d_root);
    CK(cudaGetLastError());
  }
  if (K.rows && !K.device) CK(cudaMemcpyAsync(K.rows + first * 32, d_rows, (size_t)total * 128, cudaMemcpyDeviceToHost, s0));
  if (K.root && lo == 0) CK(cudaMemcpyAsync(K.root, d_root, 32, cudaMemcpyDeviceToHost, s0));
  const bool check = (c->flags & B3W_FLAG_FUSED_CHECK) != 0;
  const size_t wbytes = (size_t)c->def->ws * 32;
  if (K.device) {
    // every step witness of the range in ONE launch, straight into the caller's device buffers
    nvtx_range r("b3w:kernel");
    launch_opts o;
    o.check = check;
    CK(cudaEventRecord(c->ev_k0[0], s0));
    rc = launch_witness(c, d_rows, total, K.out + first * wbytes, K.status ? K.status + first : nullptr, K.pub ? K.pub + first * 15 : nullptr, s0, o);
    if (rc) return rc;
    if ((c->flags & B3W_FLAG_BYTE_CHECK) && K.status) {                 // every row on the bytes just written (see batch_chunks)
      rc = r1cs_check_launch(c, K.out + first * wbytes, nullptr, total, K.status + first, nullptr, s0, false, /*skip_asserted=*/true);
      if (rc) return rc;
      c->timing.launches++;
    }
    CK(cudaEventRecord(c->ev_k1[0], s0));
    c->ev_pending[0] = true;
    c->timing.launches++;
    return B3W_OK;
  }
  CK(cudaEventRecord(c->ev[0], s0));
  CK(cudaStreamWaitEvent(c->st[1], c->ev[0], 0));                   // the rows exist before slot 1's first launch
  // all step witnesses of the range, ring chunk by ring chunk
  uint64_t done = 0;
  int k = 0;
  while (done < total) {
    const uint64_t m = total - done < c->ring_cap ? total - done : c->ring_cap;
    cudaStream_t s = c->st[k];
    launch_opts o;
    o.check = check;
    {
      nvtx_range r("b3w:kernel");
      CK(cudaEventRecord(c->ev_k0[k], s));
      rc = launch_witness(c, d_rows + done * 32, m, c->d_ring[k], c->d_status[k], c->d_pub[k], s, o);
      if (rc) return rc;
      if (c->flags & B3W_FLAG_BYTE_CHECK) {                            // every row on the bytes in the ring slot (see batch_chunks)
        rc = r1cs_check_launch(c, c->d_ring[k], nullptr, m, c->d_status[k], nullptr, s, false, /*skip_asserted=*/true);
        if (rc) return rc;
        c->timing.launches++;
      }
      CK(cudaEventRecord(c->ev_k1[k], s));
      c->ev_pending[k] = true;
      c->timing.launches++;
    }
    const uint64_t g = first + done;
    {
      nvtx_range r("b3w:d2h");
      if (K.out) { CK(cudaMemcpyAsync(K.out + g * wbytes, c->d_ring[k], (size_t)m * wbytes, cudaMemcpyDeviceToHost, s)); c->timing.d2h_bytes += (size_t)m * wbytes; }
      if (K.status) CK(cudaMemcpyAsync(K.status + g, c->d_status[k], (size_t)m, cudaMemcpyDeviceToHost, s));
      if (K.pub) CK(cudaMemcpyAsync(K.pub + g * 15, c->d_pub[k], (size_t)m * 60, cudaMemcpyDeviceToHost, s));
      c->timing.d2h_bytes += (size_t)m * 61;
    }
    done += m;
    k ^= 1;
    if (done < total) {                                              // before reusing slot k its stream must have drained
      CK(cudaStreamSynchronize(c->st[k]));
      timing_collect(c, k);
    }
  }
  return B3W_OK;
}
static int nova_chain_range(b3w_ctx *c, const chain_plan &p, const uint8_t *data, uint64_t len, uint64_t lo, uint64_t hi, const chain_sink &K) {
  ON_DEVICE(c);
  LOCKED(c);
  nvtx_range r("b3w_nova_chain");
  timing_begin(c);
  c->timing.instances = p.step_off[hi] - p.step_off[lo];
  const int rc = guarded("b3w_nova_chain", [&]() { return nova_chain_steps(c, p, data, len, lo, hi, K); });
  // drain on every path: after an error no copy into the caller's buffers may still be in flight
  cudaError_t e0 = cudaSuccess, e1 = cudaSuccess;
  if (c->ring_ready) { e0 = cudaStreamSynchronize(c->st[0]); e1 = cudaStreamSynchronize(c->st[1]); }
  timing_end(c);
  if (rc) return rc;
  if (e0 != cudaSuccess || e1 != cudaSuccess) return fail(B3W_ERR_CUDA, "b3w_nova_chain: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
  return B3W_OK;
}

static int b3w_nova_chain_impl(b3w_ctx *c, const uint8_t *data, uint64_t len, const chain_sink &K, uint64_t *step_off_out) {
  if (!c || (!data && len)) return fail(B3W_ERR_INVALID, "b3w_nova_chain: null argument");
  if (!c->def->nova) return fail(B3W_ERR_INVALID, "b3w_nova_chain needs a nova circuit context");
  if (K.device && (!K.out || ((uintptr_t)K.out & 31) != 0)) return fail(B3W_ERR_INVALID, "b3w_nova_chain_device: d_out must be a 32-byte aligned device buffer");
  chain_plan p;
  int rc = make_chain_plan(len, p, (c->flags & B3W_FLAG_REFERENCE_SIBLINGS) != 0);
  if (rc) return rc;
  if (step_off_out) memcpy(step_off_out, p.step_off.data(), (p.nc + 1) * 8);
  return nova_chain_range(c, p, data, len, 0, p.nc, K);
}
extern "C" int b3w_nova_chain(b3w_ctx *c, const uint8_t *data, uint64_t len, uint8_t *out, uint8_t *status, uint32_t *pub,
                              uint32_t *rows_out, uint64_t *step_off_out, uint8_t root_out[32]) {
  chain_sink K;
  K.out = out; K.status = status; K.pub = pub; K.rows = rows_out; K.root = root_out;
  return guarded("b3w_nova_chain", [&]() { return b3w_nova_chain_impl(c, data, len, K, step_off_out); });
}
extern "C" int b3w_nova_chain_device(b3w_ctx *c, const uint8_t *data, uint64_t len, uint8_t *d_out, uint8_t *d_status, uint32_t *d_pub,
                                     uint32_t *d_rows, uint64_t *step_off_out, uint8_t root_out[32]) {
  chain_sink K;
  K.device = true;
  K.out = d_out; K.status = d_status; K.pub = d_pub; K.rows = d_rows; K.root = root_out;
  return guarded("b3w_nova_chain_device", [&]() { return b3w_nova_chain_impl(c, data, len, K, step_off_out); });
}

// (the multi-GPU form, b3w_multi_nova_chain, is with the other b3w_multi_* entry points below: the unit of sharding is a
// chunk -- its steps chain into each other, chunks do not -- and every device hashes the whole tree.)

// ------------------------------------------------------------------------------------------------
// compact witnesses, host side
// ------------------------------------------------------------------------------------------------
static uint32_t packed_words_of(const circuit_def *d) { return (d->trace_words + 3u) & ~3u; }
static size_t trace_smem_of(const circuit_def *d) { return (size_t)WARPS_PER_CTA * (d->nova ? NOVA_TRACE_STRIDE : TRACE_STRIDE) * 4; }

extern "C" int b3w_packed_words(uint32_t circuit, uint32_t *words) {
  const circuit_def *d = find_def(circuit);
  if (!d) return B3W_ERR_UNSUPPORTED;
  if (!words) return fail(B3W_ERR_INVALID, "b3w_packed_words: null argument");
  *words = packed_words_of(d);
  return B3W_OK;
}

static int packed_launch(b3w_ctx *c, const uint32_t *d_in, uint64_t n, uint32_t *d_packed, uint8_t *d_status, uint32_t *d_pub, cudaStream_t s) {
  if (n == 0) return B3W_OK;
  const uint64_t ctas = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA, cap = (uint64_t)c->sm_count * 6;
  const unsigned grid = (unsigned)(ctas < cap ? ctas : cap);
  const uint32_t sw = packed_words_of(c->def);
  if (c->def->nova) k_witness_packed<true><<<grid, WARPS_PER_CTA * 32, trace_smem_of(c->def), s>>>(d_in, n, sw, d_packed, d_status, d_pub);
  else k_witness_packed<false><<<grid, WARPS_PER_CTA * 32, trace_smem_of(c->def), s>>>(d_in, n, sw, d_packed, d_status, d_pub);
  CK(cudaGetLastError());
  return B3W_OK;
}
extern "C" int b3w_witness_batch_packed_device(b3w_ctx *c, const uint32_t *d_in, uint64_t n, uint32_t *d_packed, uint8_t *d_status,
                                               uint32_t *d_pub, void *stream) {
  if (!c || !d_in || !d_packed) return fail(B3W_ERR_INVALID, "b3w_witness_batch_packed_device: null argument");
  if (((uintptr_t)d_packed & 15) != 0) return fail(B3W_ERR_INVALID, "d_packed must be 16-byte aligned");
  ON_DEVICE(c);
  LOCKED(c);
  return packed_launch(c, d_in, n, d_packed, d_status, d_pub, (cudaStream_t)stream);
}

extern "C" int b3w_unpack_device(b3w_ctx *c, const uint32_t *d_packed, uint64_t n, uint8_t *d_out, void *stream) {
  if (!c || !d_packed || !d_out) return fail(B3W_ERR_INVALID, "b3w_unpack_device: null argument");
  if (((uintptr_t)d_packed & 15) != 0 || ((uintptr_t)d_out & 31) != 0) return fail(B3W_ERR_INVALID, "d_packed must be 16-byte, d_out 32-byte aligned");
  ON_DEVICE(c);
  LOCKED(c);
  if (n == 0) return B3W_OK;
  const uint32_t parts = 8, part_len = ((c->def->ws + parts - 1) / parts + 31) / 32 * 32;
  const uint64_t ctas = (n * parts + WARPS_PER_CTA - 1) / WARPS_PER_CTA, cap = (uint64_t)c->sm_count * 2;
  const unsigned grid = (unsigned)(ctas < cap ? ctas : cap);
  const uint32_t sw = packed_words_of(c->def);
  if (c->def->nova) {
    k_unpack<true><<<grid, WARPS_PER_CTA * 32, trace_smem_of(c->def), (cudaStream_t)stream>>>(d_packed, n, sw, c->d_desc, c->def->ws, c->d_field, c->d_fslots, c->n_fslots, d_out, parts, part_len);
  } else {
    k_unpack<false><<<grid, WARPS_PER_CTA * 32, trace_smem_of(c->def), (cudaStream_t)stream>>>(d_packed, n, sw, c->d_desc, c->def->ws, nullptr, nullptr, 0, d_out, parts, part_len);
  }
  CK(cudaGetLastError());
  return B3W_OK;
}

// host buffers: chunks of PACKED_CHUNK instances through two device slots (allocated on first use, kept in the context), so
// that the D2H of chunk j overlaps the kernel and H2D of chunk j+1
#define PACKED_CHUNK (1u << 15)
static int ensure_packed_ring(b3w_ctx *c) {
  if (c->pk_ready) return B3W_OK;
  const circuit_def *d = c->def;
  for (int k = 0; k < 2; k++) {
    CK(cudaStreamCreateWithFlags(&c->pk_st[k], cudaStreamNonBlocking));
    CK(cudaMalloc(&c->pk_in[k], (size_t)PACKED_CHUNK * d->n_inputs * 4));
    CK(cudaMalloc(&c->pk_buf[k], (size_t)PACKED_CHUNK * packed_words_of(d) * 4));
    CK(cudaMalloc(&c->pk_status[k], (size_t)PACKED_CHUNK));
    CK(cudaMalloc(&c->pk_pub[k], (size_t)PACKED_CHUNK * d->n_public * 4));
  }
  c->pk_ready = true;
  return B3W_OK;
}
static void free_packed_ring(b3w_ctx *c) {
  for (int k = 0; k < 2; k++) {
    if (c->pk_in[k]) cudaFree(c->pk_in[k]);
    if (c->pk_buf[k]) cudaFree(c->pk_buf[k]);
    if (c->pk_status[k]) cudaFree(c->pk_status[k]);
    if (c->pk_pub[k]) cudaFree(c->pk_pub[k]);
    if (c->pk_st[k]) cudaStreamDestroy(c->pk_st[k]);
    if (c->hy_host[k]) cudaFreeHost(c->hy_host[k]);
    c->pk_in[k] = nullptr; c->pk_buf[k] = nullptr; c->pk_status[k] = nullptr; c->pk_pub[k] = nullptr; c->pk_st[k] = nullptr; c->hy_host[k] = nullptr;
  }
  c->pk_ready = false;
  c->hy_ready = false;
}

// ---- host-side expansion of packed witnesses -----------------------------------------------------------------------------
// The .wtns body of an instance is a pure function of its packed record (the trace) and the static slot table, so the
// expansion can also run on the HOST: the format conversion the reference's writer does slot by slot
// (witness_calculator.js:263-269), here for records the GPU computed.  One thread per contiguous range of instances;
// every slot is written with two non-temporal 16-byte stores (the witness is written once and not read back by this code).
// This is what b3w_witness_batch_hybrid uses to keep the 770 KB-per-witness expansion off PCIe.
static void unpack_host_range(const circuit_def *d, const uint32_t *desc, const field_consts *F, const uint32_t *packed, uint32_t sw,
                              uint64_t first, uint64_t count, uint8_t *out) {
  const uint32_t ws = d->ws;
  const bool aligned = ((uintptr_t)out & 15) == 0;
  const __m128i zero = _mm_setzero_si128();
  for (uint64_t i = first; i < first + count; i++) {
    const uint32_t *trace = packed + i * sw;
    uint8_t *dst = out + i * (size_t)ws * 32;
    if (aligned) {
      // 97 % of the slots are bits and nearly all others 32-bit words: one movd (upper lanes zero) and two streaming stores
      // per slot.  (Measured against a version that assembled every slot lane by lane: the same 190 GB/s with 16 threads --
      // the host's memory write rate bounds the expansion, not the cores.)
      for (uint32_t s = 0; s < ws; s++, dst += 32) {
        const uint32_t dsc = desc[s], t = dsc & 0xFFFFu, kind = dsc >> 24;
        __m128i lo;
        __m128i hi = zero;
        if (kind == DK_BIT) lo = _mm_cvtsi32_si128((int)((trace[t] >> ((dsc >> 16) & 31u)) & 1u));
        else if (kind == DK_W32) lo = _mm_cvtsi32_si128((int)trace[t]);
        else if (kind == DK_W64) lo = _mm_cvtsi64_si128((long long)(((uint64_t)trace[t + 1] << 32) | trace[t]));
        else {
          const int64_t x = (int64_t)(((uint64_t)trace[t + 1] << 32) | trace[t]);
          const fr_t v = kind == DK_S64 ? fr_from_s64(x, F->p) : fr_inv_s64(x, *F);
          lo = _mm_loadu_si128((const __m128i *)v.l);
          hi = _mm_loadu_si128((const __m128i *)(v.l + 4));
        }
        _mm_stream_si128((__m128i *)dst, lo);
        _mm_stream_si128((__m128i *)(dst + 16), hi);
      }
      continue;
    }
    for (uint32_t s = 0; s < ws; s++) {
      const uint32_t dsc = desc[s], t = dsc & 0xFFFFu, k = (dsc >> 16) & 31u, kind = dsc >> 24;
      uint32_t l[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (kind == DK_BIT) l[0] = (trace[t] >> k) & 1u;
      else if (kind == DK_W32) l[0] = trace[t];
      else if (kind == DK_W64) { l[0] = trace[t]; l[1] = trace[t + 1]; }
      else {
        const int64_t x = (int64_t)(((uint64_t)trace[t + 1] << 32) | trace[t]);
        const fr_t v = kind == DK_S64 ? fr_from_s64(x, F->p) : fr_inv_s64(x, *F);
        memcpy(l, v.l, 32);
      }
      memcpy(dst + (size_t)s * 32, l, 32);
    }
  }
  _mm_sfence();
}
static void unpack_host_mt(const circuit_def *d, const uint32_t *desc, const field_consts *F, const uint32_t *packed, uint32_t sw, uint64_t n,
                           uint8_t *out, uint32_t threads) {
  if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
  if (threads > n) threads = (uint32_t)(n ? n : 1);
  thread_group tg;
  for (uint32_t g = 0; g < threads; g++) {
    const uint64_t base = n / threads, extra = n % threads;
    const uint64_t first = g * base + (g < extra ? g : extra), count = base + (g < extra ? 1 : 0);
    if (count == 0) continue;
    if (threads == 1) unpack_host_range(d, desc, F, packed, sw, first, count, out);
    else tg.th.emplace_back([=]() { unpack_host_range(d, desc, F, packed, sw, first, count, out); });
  }
  tg.join();
}
extern "C" int b3w_unpack_host(b3w_ctx *c, const uint32_t *packed, uint64_t n, uint8_t *out, uint32_t threads) {
  if (!c || ((!packed || !out) && n)) return fail(B3W_ERR_INVALID, "b3w_unpack_host: null argument");
  return guarded("b3w_unpack_host", [&]() {
    unpack_host_mt(c->def, c->h_desc, c->h_field, packed, packed_words_of(c->def), n, out, threads);
    return B3W_OK;
  });
}

// packed batch from host rows; with `expand_out` the chunks are expanded on the host into .wtns bodies as they arrive
// (b3w_witness_batch_hybrid), `packed` may then be NULL
static int packed_chunks(b3w_ctx *c, const uint32_t *in, uint64_t n, uint32_t *packed, uint8_t *status, uint32_t *pub, uint8_t *expand_out,
                         uint32_t threads) {
  const circuit_def *d = c->def;
  const uint32_t sw = packed_words_of(d);
  uint64_t done = 0, pend_first[2] = {0, 0}, pend_count[2] = {0, 0};
  int k = 0;
  auto retire = [&](int q) -> int {            // chunk in slot q has left the device: expand it on the host
    CK(cudaStreamSynchronize(c->pk_st[q]));
    if (expand_out && pend_count[q]) {
      nvtx_range r("b3w:host_unpack");
      const double t0 = now_ms();
      const uint32_t *src = packed ? packed + pend_first[q] * sw : c->hy_host[q];
      unpack_host_mt(d, c->h_desc, c->h_field, src, sw, pend_count[q], expand_out + pend_first[q] * (size_t)d->ws * 32, threads);
      c->timing.host_ms += now_ms() - t0;
    }
    pend_count[q] = 0;
    return B3W_OK;
  };
  while (done < n) {
    const uint64_t m = n - done < PACKED_CHUNK ? n - done : PACKED_CHUNK;
    cudaStream_t s = c->pk_st[k];
    int rc = retire(k);                                    // slot k's previous chunk has left the device (and is expanded)
    if (rc) return rc;
    CK(cudaMemcpyAsync(c->pk_in[k], in + done * d->n_inputs, m * d->n_inputs * 4, cudaMemcpyHostToDevice, s));
    rc = packed_launch(c, c->pk_in[k], m, c->pk_buf[k], c->pk_status[k], c->pk_pub[k], s);
    if (rc) return rc;
    c->timing.launches++;
    uint32_t *dst = packed ? packed + done * sw : c->hy_host[k];
    CK(cudaMemcpyAsync(dst, c->pk_buf[k], m * sw * 4, cudaMemcpyDeviceToHost, s));
    if (status) CK(cudaMemcpyAsync(status + done, c->pk_status[k], m, cudaMemcpyDeviceToHost, s));
    if (pub) CK(cudaMemcpyAsync(pub + done * d->n_public, c->pk_pub[k], m * d->n_public * 4, cudaMemcpyDeviceToHost, s));
    c->timing.h2d_bytes += m * d->n_inputs * 4;
    c->timing.d2h_bytes += m * (sw * 4 + 1 + d->n_public * 4);
    pend_first[k] = done;
    pend_count[k] = m;
    done += m;
    k ^= 1;
  }
  int rc = retire(k);
  if (rc) return rc;
  return retire(k ^ 1);
}
static int packed_host(b3w_ctx *c, const uint32_t *in, uint64_t n, uint32_t *packed, uint8_t *status, uint32_t *pub, uint8_t *expand_out,
                       uint32_t threads, const char *who) {
  ON_DEVICE(c);
  LOCKED(c);
  nvtx_range r(who);
  timing_begin(c);
  c->timing.instances = n;
  if (n == 0) { timing_end(c); return B3W_OK; }
  int rc = ensure_packed_ring(c);
  if (rc) { free_packed_ring(c); return rc; }
  if (expand_out && !packed && !c->hy_ready) {
    for (int k = 0; k < 2; k++)
      if (cudaHostAlloc((void **)&c->hy_host[k], (size_t)PACKED_CHUNK * packed_words_of(c->def) * 4, cudaHostAllocPortable) != cudaSuccess) {
        free_packed_ring(c);
        return fail(B3W_ERR_NOMEM, "%s: pinned staging allocation failed", who);
      }
    c->hy_ready = true;
  }
  rc = guarded(who, [&]() { return packed_chunks(c, in, n, packed, status, pub, expand_out, threads); });
  // drain on every path: after an error no copy into the caller's buffers may still be in flight
  const cudaError_t e0 = cudaStreamSynchronize(c->pk_st[0]), e1 = cudaStreamSynchronize(c->pk_st[1]);
  timing_end(c);
  if (rc) return rc;
  if (e0 != cudaSuccess || e1 != cudaSuccess) return fail(B3W_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
  return B3W_OK;
}

extern "C" int b3w_witness_batch_packed(b3w_ctx *c, const uint32_t *in, uint64_t n, uint32_t *packed, uint8_t *status, uint32_t *pub) {
  if (!c || (!in && n) || (!packed && n)) return fail(B3W_ERR_INVALID, "b3w_witness_batch_packed: null argument");
  return packed_host(c, in, n, packed, status, pub, nullptr, 0, "b3w_witness_batch_packed");
}

extern "C" int b3w_witness_batch_hybrid(b3w_ctx *c, const uint32_t *in, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub,
                                        uint32_t threads) {
  if (!c || (!in && n) || (!out && n)) return fail(B3W_ERR_INVALID, "b3w_witness_batch_hybrid: null argument");
  return packed_host(c, in, n, nullptr, status, pub, out, threads, "b3w_witness_batch_hybrid");
}

// ------------------------------------------------------------------------------------------------
// multi-GPU: one context + one host thread per device, contiguous index ranges, no collective
// ------------------------------------------------------------------------------------------------
struct b3w_multi {
  std::vector<b3w_ctx *> ctx;
};

static int b3w_multi_create_impl(const b3w_config *cfg, const int32_t *devices, uint32_t n_devices, b3w_multi **out) {
  if (!cfg || !out || (n_devices && !devices)) return fail(B3W_ERR_INVALID, "b3w_multi_create: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(B3W_ERR_CUDA, "no CUDA device: %s (libblake3wit has no CPU path)", cudaGetErrorString(e));
  b3w_multi *m = new (std::nothrow) b3w_multi();
  if (!m) return fail(B3W_ERR_NOMEM, "out of host memory");
  const uint32_t n = n_devices ? n_devices : (uint32_t)ndev;          // 0 devices listed = every visible device
  for (uint32_t i = 0; i < n; i++) {
    b3w_config c = *cfg;
    c.device = n_devices ? devices[i] : (int32_t)i;
    b3w_ctx *x = nullptr;
    int rc = b3w_create(&c, &x);
    if (rc) {
      for (b3w_ctx *y : m->ctx) b3w_destroy(y);
      delete m;
      return rc;
    }
    m->ctx.push_back(x);
  }
  *out = m;
  return B3W_OK;
}
extern "C" int b3w_multi_create(const b3w_config *cfg, const int32_t *devices, uint32_t n_devices, b3w_multi **out) {
  return guarded("b3w_multi_create", [&]() { return b3w_multi_create_impl(cfg, devices, n_devices, out); });
}

extern "C" void b3w_multi_destroy(b3w_multi *m) {
  if (!m) return;
  for (b3w_ctx *x : m->ctx) b3w_destroy(x);
  delete m;
}

extern "C" uint32_t b3w_multi_size(const b3w_multi *m) { return m ? (uint32_t)m->ctx.size() : 0u; }

// shard g of G over [0, n): contiguous, balanced (the first n % G shards get one extra instance)
static void shard_of(uint64_t n, uint32_t g, uint32_t G, uint64_t *first, uint64_t *count) {
  const uint64_t base = n / G, extra = n % G;
  *first = g * base + (g < extra ? g : extra);
  *count = base + (g < extra ? 1 : 0);
}

extern "C" int b3w_shard_range(uint64_t n, uint32_t g, uint32_t n_shards, uint64_t *first, uint64_t *count) {
  if (!first || !count || n_shards == 0 || g >= n_shards) return fail(B3W_ERR_INVALID, "b3w_shard_range: bad argument");
  shard_of(n, g, n_shards, first, count);
  return B3W_OK;
}

// `ex` (may be NULL): sums / first_bad are sliced like status; every sample is fetched by the device that owns its instance
static int b3w_multi_witness_batch_impl(b3w_multi *m, const uint32_t *in, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub,
                                        const b3w_batch_extras *ex) {
  if (!m || m->ctx.empty() || (!in && n)) return fail(B3W_ERR_INVALID, "b3w_multi_witness_batch: null argument");
  const uint32_t G = (uint32_t)m->ctx.size();
  const circuit_def *d = m->ctx[0]->def;
  if (ex && ex->n_samples > B3W_MAX_SAMPLES) return fail(B3W_ERR_INVALID, "b3w_multi_witness_batch_ex: %u samples (at most %u)", ex->n_samples, B3W_MAX_SAMPLES);
  if (ex && ex->n_samples && (!ex->sample_idx || !ex->sample_out)) return fail(B3W_ERR_INVALID, "b3w_multi_witness_batch_ex: samples without sample_idx / sample_out");
  if (ex)
    for (uint32_t j = 0; j < ex->n_samples; j++)
      if (ex->sample_idx[j] >= n) return fail(B3W_ERR_INVALID, "b3w_multi_witness_batch_ex: sample %u = instance %llu of %llu", j, (unsigned long long)ex->sample_idx[j], (unsigned long long)n);
  const size_t wbytes = (size_t)d->ws * 32;
  std::vector<int> rc(G, B3W_OK);
  std::vector<std::string> err(G);
  thread_group tg;
  for (uint32_t g = 0; g < G; g++) {
    tg.th.emplace_back([&, g]() {
      uint64_t first, count;
      shard_of(n, g, G, &first, &count);
      if (count == 0) return;
      b3w_batch_extras e = {};
      std::vector<uint64_t> idx;                       // this shard's samples, relative to its range ...
      std::vector<uint32_t> pos;                       // ... and where each goes in the caller's sample_out
      std::vector<uint8_t> tmp;
      if (ex) {
        if (ex->sums) e.sums = ex->sums + first;
        if (ex->first_bad) e.first_bad = ex->first_bad + first;
        for (uint32_t j = 0; j < ex->n_samples; j++)
          if (ex->sample_idx[j] >= first && ex->sample_idx[j] < first + count) { idx.push_back(ex->sample_idx[j] - first); pos.push_back(j); }
      }
      // the shard's samples arrive in a contiguous scratch and are then put in the caller's order
      if (!idx.empty()) {
        try { tmp.resize(idx.size() * wbytes); } catch (...) { rc[g] = B3W_ERR_NOMEM; err[g] = "out of host memory"; return; }
        e.sample_idx = idx.data(); e.n_samples = (uint32_t)idx.size(); e.sample_out = tmp.data();
      }
      rc[g] = b3w_witness_batch_ex(m->ctx[g], in + first * d->n_inputs, count, out ? out + first * wbytes : nullptr,
                                   status ? status + first : nullptr, pub ? pub + first * d->n_public : nullptr, &e);
      if (rc[g]) { err[g] = g_err; return; }            // g_err is thread-local: carry the text over to the caller
      for (size_t k = 0; k < idx.size(); k++) memcpy(ex->sample_out + (size_t)pos[k] * wbytes, tmp.data() + k * wbytes, wbytes);
    });
  }
  tg.join();
  for (uint32_t g = 0; g < G; g++)
    if (rc[g]) return fail(rc[g], "device %d (shard %u of %u): %s", m->ctx[g]->device, g, G, err[g].c_str());
  return B3W_OK;
}
extern "C" int b3w_multi_witness_batch(b3w_multi *m, const uint32_t *in, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub) {
  return guarded("b3w_multi_witness_batch", [&]() { return b3w_multi_witness_batch_impl(m, in, n, out, status, pub, nullptr); });
}
extern "C" int b3w_multi_witness_batch_ex(b3w_multi *m, const uint32_t *in, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub,
                                          const b3w_batch_extras *extras) {
  return guarded("b3w_multi_witness_batch_ex", [&]() { return b3w_multi_witness_batch_impl(m, in, n, out, status, pub, extras); });
}

static int b3w_multi_nova_chain_impl(b3w_multi *m, const uint8_t *data, uint64_t len, uint8_t *out, uint8_t *status, uint32_t *pub,
                                    uint32_t *rows_out, uint64_t *step_off_out, uint8_t root_out[32]) {
  if (!m || m->ctx.empty() || (!data && len)) return fail(B3W_ERR_INVALID, "b3w_multi_nova_chain: null argument");
  if (!m->ctx[0]->def->nova) return fail(B3W_ERR_INVALID, "b3w_multi_nova_chain needs a nova circuit context");
  chain_plan p;
  int rc0 = make_chain_plan(len, p, (m->ctx[0]->flags & B3W_FLAG_REFERENCE_SIBLINGS) != 0);
  if (rc0) return rc0;
  if (step_off_out) memcpy(step_off_out, p.step_off.data(), (p.nc + 1) * 8);
  const uint32_t G = (uint32_t)m->ctx.size();
  // chunk boundaries that split the STEPS evenly: device g takes chunks [cut[g], cut[g+1])
  std::vector<uint64_t> cut(G + 1, p.nc);
  cut[0] = 0;
  for (uint32_t g = 1; g < G; g++) {
    const uint64_t target = p.total / G * g;
    cut[g] = (uint64_t)(std::lower_bound(p.step_off.begin(), p.step_off.end(), target) - p.step_off.begin());
    if (cut[g] > p.nc) cut[g] = p.nc;
    if (cut[g] < cut[g - 1]) cut[g] = cut[g - 1];
  }
  chain_sink K;
  K.out = out; K.status = status; K.pub = pub; K.rows = rows_out; K.root = root_out;
  std::vector<int> rc(G, B3W_OK);
  std::vector<std::string> err(G);
  thread_group tg;
  for (uint32_t g = 0; g < G; g++) {
    if (cut[g] == cut[g + 1]) continue;
    tg.th.emplace_back([&, g]() {
      rc[g] = nova_chain_range(m->ctx[g], p, data, len, cut[g], cut[g + 1], K);
      if (rc[g]) err[g] = g_err;
    });
  }
  tg.join();
  for (uint32_t g = 0; g < G; g++)
    if (rc[g]) return fail(rc[g], "device %d (chunks %llu..%llu): %s", m->ctx[g]->device, (unsigned long long)cut[g],
                           (unsigned long long)cut[g + 1], err[g].c_str());
  return B3W_OK;
}
extern "C" int b3w_multi_nova_chain(b3w_multi *m, const uint8_t *data, uint64_t len, uint8_t *out, uint8_t *status, uint32_t *pub,
                                    uint32_t *rows_out, uint64_t *step_off_out, uint8_t root_out[32]) {
  return guarded("b3w_multi_nova_chain", [&]() { return b3w_multi_nova_chain_impl(m, data, len, out, status, pub, rows_out, step_off_out, root_out); });
}

// NUMA node of a CUDA device (from sysfs; -1 when unknown)
static int device_numa_node(int device) {
  char pci[32] = {0}, path[128];
  if (cudaDeviceGetPCIBusId(pci, sizeof pci, device) != cudaSuccess) return -1;
  for (char *p = pci; *p; p++) *p = (char)tolower(*p);
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", pci);
  FILE *f = fopen(path, "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}

// Pinned host memory on the NUMA node the device hangs off: with 8 GPUs on two sockets, D2H streams that cross the
// socket interconnect cap the whole box (measured: 92 GB/s total at N = 8 against 174 GB/s at N = 4 with default placement).
// The pages are faulted in inside cudaHostAlloc, so a preferred-node policy around the call is enough.
extern "C" void *b3w_host_alloc_near(size_t bytes, int device) {
  int dev = device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) dev = -1;
  const int node = dev >= 0 ? device_numa_node(dev) : -1;
  bool policy_set = false;
#ifdef __linux__
  if (node >= 0 && node < 1024) {
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] = 1ul << (node % (8 * sizeof(unsigned long)));
    policy_set = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, (unsigned long)(8 * sizeof mask)) == 0;
  }
#endif
  void *p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
#ifdef __linux__
  if (policy_set) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
#endif
  if (e != cudaSuccess) {
    fail(B3W_ERR_NOMEM, "cudaHostAlloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}

extern "C" void *b3w_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    fail(B3W_ERR_NOMEM, "cudaHostAlloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
extern "C" void b3w_host_free(void *p) {
  if (p) cudaFreeHost(p);
}
