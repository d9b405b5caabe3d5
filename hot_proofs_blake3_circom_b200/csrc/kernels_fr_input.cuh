// kernels_fr_input.cuh -- field-element inputs on the device.  Included by blake3wit.cu only, after kernels_nova_wide.cuh.
//
// The reference takes ANY field element for every input (`normalize`, witness_calculator.js:319-323) and rust_fold holds its
// step inputs as `Vec<(String, Vec<F>)>` (rust_fold/src/blake3_circuit.rs:197-289): b3w_witness_batch_fr is the entry point
// for that form.  Its Fr256 rows (32 bytes per input) are copied to the device as they are and converted HERE, one warp per
// instance, lane = input: reduce mod p, classify, and emit what the witness kernels take --
//   blake3_compression: the u32 row + the signed high parts m_ext of the message words (wide_domain.h's rules: h, t, b, d must
//     fit 32 bits, a message word must be ext * 2^32 + lo with ext in [-2, 3], else the instance is marked B3W_EXT_ASSERT);
//   nova step circuits: the u32 row when every input fits 32 bits (the hot kernel's domain).  An instance that holds
//     anything else gets a row on which the hot kernel asserts at once (leaf_depth = depth = 0: no expansion, no stores), its
//     index is appended to the chunk's WIDE LIST and its 32 inputs are written back in canonical form: the general kernel
//     (kernels_nova_wide.cuh) then generates exactly the listed instances into the same output buffers.
// So a batch with 1 % field-valued instances costs the hot kernel's time plus 1 % of the general kernel's, instead of
// sending the whole batch to the general kernel from a single-threaded host loop (round 1).
#pragma once

__device__ __forceinline__ fr_t fr_load_reduced(const uint8_t *p32, const fr_t &p) {
  fr_t v;
  ld_slot(p32, v.l);
  while (fr_gte(v, p)) fr_raw_sub(v, v, p);            // p > 2^253: a 256-bit value needs at most a few rounds
  return v;
}

// compression: 28 inputs per instance
__global__ void __launch_bounds__(256) k_fr_to_rows_compression(const uint8_t *__restrict__ in_fr, uint64_t n, const field_consts *__restrict__ F,
                                                                uint32_t *__restrict__ rows, int8_t *__restrict__ m_ext) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  fr_t p;
#pragma unroll
  for (int j = 0; j < 8; j++) p.l[j] = F->p.l[j];
  for (uint64_t i = warp; i < n; i += nwarps) {
    uint32_t word = 0;
    int ext = 0;
    bool dead = false;
    if (lane < 28) {
      const fr_t v = fr_load_reduced(in_fr + (i * 28 + lane) * 32, p);
      word = v.l[0];
      if (lane < 8 || lane >= 24) {
        dead = !nw_fits(v, 32);                          // h, t, b, d: ToBits(32) cannot hold it
      } else if (nw_fits(v, 34)) {
        ext = (int)v.l[1];
      } else {
        fr_t k;                                          // v = p - k: the integer -k, if k <= 2^33
        fr_raw_sub(k, p, v);
        const uint64_t k64 = ((uint64_t)k.l[1] << 32) | k.l[0];
        if (!nw_fits(k, 34) || k64 > (1ull << 33)) dead = true;
        const uint64_t x = 0ull - k64;
        word = (uint32_t)x;
        ext = (int)(int32_t)(uint32_t)(x >> 32);         // -1 or -2
      }
    }
    dead = __any_sync(0xffffffffu, dead);
    if (lane < 28) rows[i * 28 + lane] = word;
    if (lane >= 8 && lane < 24) m_ext[i * 16 + (lane - 8)] = dead ? (lane == 8 ? (int8_t)B3W_EXT_ASSERT : (int8_t)0) : (int8_t)ext;
  }
}

// nova: 32 inputs per instance.  wlist[0] = number of listed instances (zeroed before the launch), wlist[1 + k] = their indices.
__global__ void __launch_bounds__(256) k_fr_to_rows_nova(uint8_t *__restrict__ in_fr /* rewritten in canonical form */, uint64_t n,
                                                         const field_consts *__restrict__ F, uint32_t *__restrict__ rows,
                                                         uint32_t *__restrict__ wlist) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  fr_t p;
#pragma unroll
  for (int j = 0; j < 8; j++) p.l[j] = F->p.l[j];
  for (uint64_t i = warp; i < n; i += nwarps) {
    uint8_t *src = in_fr + (i * 32 + lane) * 32;
    const fr_t v = fr_load_reduced(src, p);
    const bool wide = __any_sync(0xffffffffu, !nw_fits(v, 32));
    uint32_t word = v.l[0];
    if (wide) {
      *reinterpret_cast<uint4 *>(src) = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
      *reinterpret_cast<uint4 *>(src + 16) = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
      if (lane == 12 || lane == 14) word = 0u;         // leaf_depth = depth = 0: the hot kernel's CheckDepth asserts at once
      if (lane == 0) wlist[1 + atomicAdd(wlist, 1u)] = (uint32_t)i;
    }
    rows[i * 32 + lane] = word;
  }
}
