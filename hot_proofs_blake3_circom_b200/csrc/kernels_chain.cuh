// kernels_chain.cuh -- device side of the chained-chunk driver (BASELINE config 3): BLAKE3 tree hashing and the step
// rows of every chunk.  Included by blake3wit.cu only, after kernels_witness.cuh.
#pragma once
// ------------------------------------------------------------------------------------------------
// Chained-chunk driver (BASELINE config 3): the step schedule of the reference's Nova driver
// (rust_fold/src/main.rs:71-94,130-142,166-171 and rust_fold/src/blake3_circuit.rs:160-290), batched.
// Step i+1 consumes step i's outputs (h, block_count, depth), but those are plain BLAKE3 chaining values, so the
// whole chain of every chunk is pre-computed with native u32 compressions and all step witnesses are then
// generated independently by k_blake3_nova_witness.  The sibling chaining values that the reference gets from
// bao slice extraction (rust_fold/src/blake3_hash.rs:17-93) come from a BLAKE3 tree hashed on the device.
// ------------------------------------------------------------------------------------------------
#define B3_CHUNK_START 1u
#define B3_CHUNK_END 2u
#define B3_PARENT 4u
#define B3_ROOT 8u

__device__ __forceinline__ void b3_g(uint32_t *v, int a, int b, int c, int d, uint32_t x, uint32_t y) {
  v[a] = v[a] + v[b] + x; v[d] = rotr32(v[d] ^ v[a], 16);
  v[c] = v[c] + v[d];     v[b] = rotr32(v[b] ^ v[c], 12);
  v[a] = v[a] + v[b] + y; v[d] = rotr32(v[d] ^ v[a], 8);
  v[c] = v[c] + v[d];     v[b] = rotr32(v[b] ^ v[c], 7);
}
// plain BLAKE3 compression, first 8 output words (the chaining value)
__device__ void b3_compress_cv(const uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t t1, uint32_t blen,
                               uint32_t flags, uint32_t out[8]) {
  uint32_t v[16] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7],
                    0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, t0, t1, blen, flags};
#pragma unroll 1
  for (int r = 0; r < 7; r++) {
    const uint8_t *s = MSG_SCHED[r];
    b3_g(v, 0, 4, 8, 12, m[s[0]], m[s[1]]);   b3_g(v, 1, 5, 9, 13, m[s[2]], m[s[3]]);
    b3_g(v, 2, 6, 10, 14, m[s[4]], m[s[5]]);  b3_g(v, 3, 7, 11, 15, m[s[6]], m[s[7]]);
    b3_g(v, 0, 5, 10, 15, m[s[8]], m[s[9]]);  b3_g(v, 1, 6, 11, 12, m[s[10]], m[s[11]]);
    b3_g(v, 2, 7, 8, 13, m[s[12]], m[s[13]]); b3_g(v, 3, 4, 9, 14, m[s[14]], m[s[15]]);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = v[i] ^ v[i + 8];
}
__constant__ uint32_t B3_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};

// bytes [off, off+64) of the (zero padded) input as 16 little-endian words + the count of real bytes
__device__ __forceinline__ uint32_t load_block(const uint8_t *data, uint64_t len, uint64_t off, uint32_t m[16]) {
  const uint32_t *w = reinterpret_cast<const uint32_t *>(data + off);     // the device copy is padded to 64 B
#pragma unroll
  for (int i = 0; i < 16; i++) m[i] = w[i];
  return off >= len ? 0u : (uint32_t)(len - off < 64 ? len - off : 64);
}
__device__ __forceinline__ uint32_t chunk_blocks(uint64_t len, uint64_t c) {
  const uint64_t cb = len - c * 1024 < 1024 ? len - c * 1024 : 1024;   // bytes in chunk c
  const uint32_t nb = (uint32_t)((cb + 63) / 64);                          // utils.rs:112-114
  return nb ? nb : 1;                                                      // the empty input is one empty block
}

// chunk chaining values: cv[c] for c < n_chunks (one thread per chunk)
__global__ void k_chunk_cvs(const uint8_t *__restrict__ data, uint64_t len, uint64_t n_chunks, uint32_t *__restrict__ cv) {
  const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  uint32_t h[8], m[16];
#pragma unroll
  for (int i = 0; i < 8; i++) h[i] = B3_IV[i];
  const uint32_t nb = chunk_blocks(len, c);
  for (uint32_t k = 0; k < nb; k++) {
    const uint32_t bl = load_block(data, len, c * 1024 + 64ull * k, m);
    b3_compress_cv(h, m, (uint32_t)c, (uint32_t)(c >> 32), bl, (k == 0 ? B3_CHUNK_START : 0u) | (k == nb - 1 ? B3_CHUNK_END : 0u), h);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) cv[c * 8 + i] = h[i];
}
// one tree level: parent j = compress(IV, cv[left] || cv[right], PARENT)
__global__ void k_parent_cvs(const uint32_t *__restrict__ nodes /* [first..first+count) x {left, right} */, uint32_t first,
                             uint32_t count, uint64_t n_chunks, uint32_t *__restrict__ cv) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  const uint32_t l = nodes[2 * (first + j)], r = nodes[2 * (first + j) + 1];
  uint32_t m[16], h[8], o[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { m[i] = cv[(uint64_t)l * 8 + i]; m[8 + i] = cv[(uint64_t)r * 8 + i]; h[i] = B3_IV[i]; }
  b3_compress_cv(h, m, 0, 0, 64, B3_PARENT, o);
#pragma unroll
  for (int i = 0; i < 8; i++) cv[(n_chunks + first + j) * 8 + i] = o[i];
}
// Step rows of every chunk (one thread per chunk): blake3_circuit.rs format_input() (:197-289) applied along
// update_for_step() (:185-195), with z0 from main.rs:130-142 and z_{i+1} = the circuit's outputs.
__global__ void k_chain_rows(const uint8_t *__restrict__ data, uint64_t len, uint64_t chunk_lo, uint64_t chunk_hi,
                             const uint32_t *__restrict__ cv, const uint32_t *__restrict__ path /* [chunk][max_depth] sibling refs */,
                             const uint32_t *__restrict__ depth_of /* parents above chunk c */, uint32_t max_depth,
                             const uint64_t *__restrict__ step_off, uint32_t *__restrict__ rows /* of chunks [lo, hi) */,
                             uint32_t *__restrict__ root /* h_out of chunk 0's last step */) {
  const uint64_t c = chunk_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= chunk_hi) return;
  const uint32_t n_par = depth_of[c];                 // parent_path.len()
  const uint32_t total_depth = n_par + 1;             // = leaf_depth (blake3_circuit.rs:169, main.rs:71)
  const uint32_t nb = chunk_blocks(len, c);
  uint32_t h[8], m[16];
#pragma unroll
  for (int i = 0; i < 8; i++) h[i] = B3_IV[i];
  uint32_t block_count = 0, depth = total_depth - 1;
  uint32_t *row = rows + (step_off[c] - step_off[chunk_lo]) * 32;
  const uint32_t steps = nb + total_depth - 1;        // main.rs:94
  for (uint32_t st = 0; st < steps; st++, row += 32) {
    const bool leaf = st < nb;
    uint32_t bl;
    if (leaf) {
      bl = load_block(data, len, c * 1024 + 64ull * st, m);                 // :207-224
    } else {
      const uint32_t sib = path[c * max_depth + depth];                    // parent_path[current_depth] (:234)
#pragma unroll
      for (int i = 0; i < 8; i++) { m[i] = cv[(uint64_t)sib * 8 + i]; m[8 + i] = 0; }
      bl = 64;                                                             // :229
    }
    row[0] = nb; row[1] = block_count;
#pragma unroll
    for (int i = 0; i < 8; i++) row[2 + i] = h[i];
    row[10] = (uint32_t)c; row[11] = (uint32_t)(c >> 32);
    row[12] = total_depth; row[13] = total_depth; row[14] = depth;
#pragma unroll
    for (int i = 0; i < 16; i++) row[15 + i] = m[i];
    row[31] = bl;
    // what the circuit will output (circuits/blake3_nova.circom:122-167, 229-266), natively
    const bool is_parent = depth + 1 < total_depth, is_root = depth == 0;
    const bool last = block_count + 1 == nb;
    uint32_t mm[16], hh[8];
    uint32_t flags;
    if (is_parent) {
      const bool left = ((c >> (total_depth - 2 - depth)) & 1) == 0;       // Blake3GetDownLeftPath (:47-84)
#pragma unroll
      for (int i = 0; i < 8; i++) { mm[i] = left ? h[i] : m[i]; mm[8 + i] = left ? m[i] : h[i]; hh[i] = B3_IV[i]; }
      flags = B3_PARENT | (is_root ? B3_ROOT : 0u);
      b3_compress_cv(hh, mm, 0, 0, bl, flags, h);
    } else {
      flags = (block_count == 0 ? B3_CHUNK_START : 0u) | (last ? B3_CHUNK_END : 0u) | (last && is_root ? B3_ROOT : 0u);
      b3_compress_cv(h, m, (uint32_t)c, (uint32_t)(c >> 32), bl, flags, h);
      block_count += 1;
    }
    if ((is_parent || last) && !is_root) depth -= 1;
  }
  if (c == 0) {
#pragma unroll
    for (int i = 0; i < 8; i++) root[i] = h[i];
  }
}

