// kernels_aux.cuh -- compact witnesses (pack / unpack), checksums and the pure-store calibration kernels.  Included by blake3wit.cu only, after kernels_witness.cuh.
#pragma once
// ---- compact ("packed") witnesses: SURVEY.md 8(f) rank 3 --------------------------------------------------------
// Every slot of a witness is a pure function of the instance's trace (<= 1 324 u32) and the static slot table, so the
// trace IS the witness in compact form: 3 776 B (compression) / 5 296 B (nova) instead of 770 976 / 745 312 B, ~200x less
// to keep in HBM, move over PCIe or hand to a prover on the same GPU.  k_witness_packed writes traces, k_unpack expands
// traces that are resident in device memory into the .wtns body layout (the expansion phase of the main kernels).
template <bool NOVA>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_witness_packed(const uint32_t *__restrict__ in, uint64_t n, uint32_t stride_words, uint32_t *__restrict__ packed,
                 uint8_t *__restrict__ status, uint32_t *__restrict__ pub) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t *trace = s_dyn + wib * (NOVA ? NOVA_TRACE_STRIDE : TRACE_STRIDE);
  const lane_sched ls = load_lane_sched(lane);
  const uint64_t warp = (uint64_t)blockIdx.x * WARPS_PER_CTA + wib, nwarps = (uint64_t)gridDim.x * WARPS_PER_CTA;
  constexpr int N_IN = NOVA ? 32 : 28, N_PUB = NOVA ? 15 : 16;
  for (uint64_t i = warp; i < n; i += nwarps) {
    __syncwarp();
    for (uint32_t w = lane; w < stride_words; w += 32) trace[w] = 0u;     // words no template writes stay 0
    __syncwarp();
    if (lane == 0) trace[TR_ONE] = 1u;
    if (lane < N_IN) trace[(NOVA ? NV_IN : TR_IN) + lane] = __ldg(in + i * N_IN + lane);
    __syncwarp();
    bool ok = true;
    if (NOVA) ok = nova_trace(trace, lane);
    __syncwarp();
    if (ok) compression_trace(trace, lane, ls);
    __syncwarp();
    if (!ok && lane == 0) trace[TR_ONE] = 0u;                             // marks "no witness exists" (Assert Failed.)
    if (status && lane == 0) status[i] = ok ? 0 : B3W_CIRCOM_ASSERT;
    if (pub && lane < N_PUB) pub[i * N_PUB + lane] = !ok ? 0u : NOVA ? nova_public_output(trace, lane) : trace[TR_OUT + lane];
    __syncwarp();
    uint4 *dst = reinterpret_cast<uint4 *>(packed + i * stride_words);
    const uint4 *src = reinterpret_cast<const uint4 *>(trace);
    for (uint32_t q = lane; q < stride_words / 4; q += 32) dst[q] = src[q];
  }
}

template <bool HAS_FIELD>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_unpack(const uint32_t *__restrict__ packed, uint64_t n, uint32_t stride_words, const uint32_t *__restrict__ desc, uint32_t ws,
         const field_consts *__restrict__ F, const uint2 *__restrict__ fslots, uint32_t n_fslots, uint8_t *__restrict__ out,
         uint32_t parts, uint32_t part_len) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t *trace = s_dyn + wib * (HAS_FIELD ? NOVA_TRACE_STRIDE : TRACE_STRIDE);
  const uint64_t warp = (uint64_t)blockIdx.x * WARPS_PER_CTA + wib, nwarps = (uint64_t)gridDim.x * WARPS_PER_CTA;
  const uint64_t total = n * parts;
  for (uint64_t item = warp; item < total; item += nwarps) {
    const uint64_t i = item / parts;
    const uint32_t part = (uint32_t)(item % parts);
    const uint32_t a = part * part_len, b = a + part_len < ws ? a + part_len : ws;
    __syncwarp();
    const uint4 *src = reinterpret_cast<const uint4 *>(packed + i * stride_words);
    uint4 *dst = reinterpret_cast<uint4 *>(trace);
    for (uint32_t q = lane; q < stride_words / 4; q += 32) dst[q] = __ldg(src + q);
    __syncwarp();
    expand_slots<HAS_FIELD>(trace, desc, a, b, out + i * (uint64_t)ws * 32, lane, F, fslots, n_fslots);
  }
}

// ---- checksum of resident witnesses (verification helper; reads HBM) ----
__device__ __forceinline__ uint64_t mix64(uint64_t x) { return (x + 1) * 0x9E3779B97F4A7C15ull; }
__global__ void __launch_bounds__(256) k_checksum(const uint64_t *__restrict__ wit, uint64_t n, uint32_t ws,
                                                  uint64_t *__restrict__ sums) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t i = warp; i < n; i += nwarps) {
    // 16 bytes per lane, four loads in flight: wide requests are what compressible memory (device_mem.h) rewards --
    // 8-byte lane loads read such a buffer at 2 TB/s, these at the full rate
    const ulonglong2 *w = reinterpret_cast<const ulonglong2 *>(wit + i * (uint64_t)ws * 4);
    const uint32_t nq = ws * 2;
    uint64_t acc = 0;
    for (uint32_t q0 = lane; q0 < nq; q0 += 128) {
      ulonglong2 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) v[u] = __ldg(w + min(q0 + 32u * u, nq - 1));
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t q = q0 + 32u * u;
        if (q < nq) acc += (v[u].x + 1) * mix64(2 * q) + (v[u].y + 1) * mix64(2 * q + 1);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sums[i] = acc;
  }
}

// ---- pure-store calibration ----
__global__ void __launch_bounds__(256) k_fill(uint8_t *buf, uint64_t nslots) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += stride)
    st_slot(buf + s * 32, (uint32_t)s & 1u, 0u, 0u, 0u, 0u, 0u, 0u, 0u);
}

// The same store stream as the witness kernels without any of their work: warps take 32 KiB items from the dynamic
// counters and write them with 1 KiB warp stores.  What this reaches is the ceiling of the store path for this access
// pattern; the witness kernel is judged against it (and against the driver's copy benchmark).
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_fill_items(uint8_t *buf, uint64_t n_items, uint32_t item_slots, const sched_args sc) {
  const int lane = threadIdx.x & 31;
  uint32_t sub = (uint32_t)((blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5)) % SCHED_LANES), tries = 0;
  while (tries < SCHED_LANES) {
    unsigned long long id = lane == 0 ? atomicAdd(sc.counter + sub * SCHED_STRIDE, 1ull) * SCHED_LANES + sub : 0ull;
    id = __shfl_sync(0xffffffffu, id, 0);
    if (id >= n_items) { tries++; sub = (sub + 1) % SCHED_LANES; continue; }
    uint8_t *dst = buf + id * (uint64_t)item_slots * 32;
#pragma unroll 4
    for (uint32_t sl = lane; sl < item_slots; sl += 32) st_slot(dst + (size_t)sl * 32, sl & 1u, 0u, 0u, 0u, 0u, 0u, 0u, 0u);
  }
}


// The same item stream written by the TMA engine instead of the LSU: one elected thread per CTA issues bulk copies
// (cp.async.bulk.global.shared::cta, SASS UBLKCP) of a constant shared-memory tile.  Measures what staging expanded
// witness tiles in shared memory + bulk stores could reach at best ("considered and not built").
__global__ void __launch_bounds__(128) k_fill_bulk(uint8_t *buf, uint64_t n_items, uint32_t item_bytes, const sched_args sc) {
  extern __shared__ __align__(128) uint8_t s_tile[];
  for (uint32_t i = threadIdx.x * 16; i < item_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4 *>(s_tile + i) = make_uint4((i >> 5) & 1u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  const uint32_t src = (uint32_t)__cvta_generic_to_shared(s_tile);
  uint32_t sub = blockIdx.x % SCHED_LANES, tries = 0;
  while (tries < SCHED_LANES) {
    const unsigned long long id = atomicAdd(sc.counter + sub * SCHED_STRIDE, 1ull) * SCHED_LANES + sub;
    if (id >= n_items) { tries++; sub = (sub + 1) % SCHED_LANES; continue; }
    uint8_t *dst = buf + id * (uint64_t)item_bytes;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(item_bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");      // the tile is constant: only bound the queue depth
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
