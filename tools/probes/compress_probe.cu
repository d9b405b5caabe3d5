// compress_probe.cu -- does this GPU / driver grant COMPRESSIBLE device memory (compute data compression: L2 compresses
// lines on their way to HBM), and what does it do to a witness-shaped store stream?  Scratch probe, built and run by hand:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o compress_probe compress_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CKD(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_; cuGetErrorString(r_, &s_); printf("%s -> %s\n", #x, s_); return 1; } } while (0)
#define CKR(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

// witness-shaped data: 32-byte slots {bit, 0, 0, 0, 0, 0, 0, 0}; mode 1: every 32nd slot holds a random 32-bit word
__global__ void k_store(uint8_t *out, uint64_t slots, int mode) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += stride) {
    uint32_t h = (uint32_t)(s * 2654435761u) ^ (uint32_t)(s >> 7);
    uint32_t lo = mode == 2 ? h : (mode == 1 && (s & 31) == 0) ? h : (h >> 13) & 1u;
    asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.b32 [%0], {%1,%2,%2,%2,%2,%2,%2,%2};" ::"l"(out + s * 32), "r"(lo), "r"(0u) : "memory");
  }
}
__global__ void k_sum(const uint4 *in, uint64_t n16, unsigned long long *acc) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long a = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) { uint4 v = in[i]; a += v.x + v.y + v.z + v.w; }
  atomicAdd(acc, a);
}

static float time_store(uint8_t *p, uint64_t bytes, int mode) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; i++) k_store<<<148 * 8, 256>>>(p, bytes / 32, mode);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; i++) k_store<<<148 * 8, 256>>>(p, bytes / 32, mode);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
static float time_read(uint8_t *p, uint64_t bytes, unsigned long long *d_acc) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_sum<<<148 * 8, 256>>>((const uint4 *)p, bytes / 16, d_acc);
  cudaEventRecord(e0);
  for (int i = 0; i < 3; i++) k_sum<<<148 * 8, 256>>>((const uint4 *)p, bytes / 16, d_acc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 3;
}

int main() {
  CKR(cudaSetDevice(0)); CKR(cudaFree(0));
  CUdevice dev; CKD(cuDeviceGet(&dev, 0));
  int sup = -1; CKD(cuDeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev));
  printf("GENERIC_COMPRESSION_SUPPORTED = %d\n", sup);
  const uint64_t want = 16ull << 30;
  uint8_t *plain; CKR(cudaMalloc(&plain, want));
  unsigned long long *d_acc; CKR(cudaMalloc(&d_acc, 8)); cudaMemset(d_acc, 0, 8);
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
  prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
  size_t gran = 0; CKD(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
  const size_t size = (want + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h; CKD(cuMemCreate(&h, size, &prop, 0));
  CUmemAllocationProp got = {}; CKD(cuMemGetAllocationPropertiesFromHandle(&got, h));
  printf("granularity %zu, compressionType granted = %d (1 = generic)\n", gran, (int)got.allocFlags.compressionType);
  CUdeviceptr va; CKD(cuMemAddressReserve(&va, size, 0, 0, 0)); CKD(cuMemMap(va, size, 0, h, 0));
  CUmemAccessDesc acc = {}; acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE; CKD(cuMemSetAccess(va, size, &acc, 1));
  uint8_t *comp = (uint8_t *)va;
  for (int mode = 0; mode < 3; mode++) {
    const float a = time_store(plain, want, mode), b = time_store(comp, want, mode);
    const float ra = time_read(plain, want, d_acc), rb = time_read(comp, want, d_acc);
    printf("mode %d (%s): store plain %.3f ms = %.0f GB/s | compressible %.3f ms = %.0f GB/s || read plain %.0f GB/s | compressible %.0f GB/s\n", mode,
           mode == 0 ? "bits only" : mode == 1 ? "bits + a word per 32 slots" : "random low word in every slot", a, want / a / 1e6, b, want / b / 1e6, want / ra / 1e6, want / rb / 1e6);
  }
  // same data in both?
  unsigned long long s1 = 0, s2 = 0;
  cudaMemset(d_acc, 0, 8); k_sum<<<148 * 8, 256>>>((const uint4 *)plain, want / 16, d_acc); cudaMemcpy(&s1, d_acc, 8, cudaMemcpyDeviceToHost);
  cudaMemset(d_acc, 0, 8); k_sum<<<148 * 8, 256>>>((const uint4 *)comp, want / 16, d_acc); cudaMemcpy(&s2, d_acc, 8, cudaMemcpyDeviceToHost);
  printf("checksums %llu %llu %s\n", s1, s2, s1 == s2 ? "equal" : "DIFFERENT");
  return 0;
}
