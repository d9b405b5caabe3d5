// device_mem.h -- COMPRESSIBLE device memory for witness buffers.  Included by blake3wit.cu only.
//
// Blackwell's L2 can compress lines on their way to HBM ("compute data compression"); memory has to be allocated as
// compressible for that, which only the driver's virtual-memory API offers (cuMemCreate with
// CU_MEM_ALLOCATION_COMP_GENERIC; cudaMalloc memory is never compressed).  A witness is the ideal payload: every 32-byte
// slot is a bit or a word followed by 24+ zero bytes.  Measured on B200 (profiles/r01j_compressible.jsonl): the
// blake3_compression witness kernel writes 2^16 witnesses in 6.18 ms instead of 6.99 ms (10.6 M witnesses/s, 8.18 TB/s of
// witness bytes: more than the HBM interface moves uncompressed), wide coalesced reads of such a buffer run at 8.9 TB/s
// instead of 6.6; narrow (8 bytes per lane) reads are slower than on ordinary memory.
// The driver entry points come from cudaGetDriverEntryPoint, so the library still does not link libcuda and still loads
// on a machine without a driver.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>

struct vmm_api {
  PFN_cuMemCreate create;
  PFN_cuMemRelease release;
  PFN_cuMemAddressReserve reserve;
  PFN_cuMemAddressFree addr_free;
  PFN_cuMemMap map;
  PFN_cuMemUnmap unmap;
  PFN_cuMemSetAccess set_access;
  PFN_cuMemGetAllocationGranularity granularity;
  PFN_cuMemGetAllocationPropertiesFromHandle props;
  PFN_cuDeviceGetAttribute dev_attr;
  bool ok;
};

static bool vmm_load(vmm_api &v) {
  memset(&v, 0, sizeof v);
  struct { const char *name; void **fn; } want[] = {
      {"cuMemCreate", (void **)&v.create}, {"cuMemRelease", (void **)&v.release}, {"cuMemAddressReserve", (void **)&v.reserve},
      {"cuMemAddressFree", (void **)&v.addr_free}, {"cuMemMap", (void **)&v.map}, {"cuMemUnmap", (void **)&v.unmap},
      {"cuMemSetAccess", (void **)&v.set_access}, {"cuMemGetAllocationGranularity", (void **)&v.granularity},
      {"cuMemGetAllocationPropertiesFromHandle", (void **)&v.props}, {"cuDeviceGetAttribute", (void **)&v.dev_attr}};
  for (auto &w : want) {
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(w.name, w.fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !*w.fn) return false;
  }
  v.ok = true;
  return true;
}

struct vmm_block { CUdeviceptr va; size_t size; CUmemGenericAllocationHandle handle; bool compressed; };

// Maps `bytes` of device memory of `device`, compressible if the device grants it.  Returns nullptr on failure (err set).
static void *vmm_alloc(const vmm_api &v, int device, size_t bytes, bool want_compression, vmm_block &blk, const char **err) {
  *err = nullptr;
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof prop);
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  int sup = 0;
  if (want_compression && v.dev_attr(&sup, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, device) == CUDA_SUCCESS && sup)
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
  size_t gran = 0;
  if (v.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) { *err = "cuMemGetAllocationGranularity failed"; return nullptr; }
  blk.size = (bytes + gran - 1) / gran * gran;
  if (v.create(&blk.handle, blk.size, &prop, 0) != CUDA_SUCCESS) { *err = "cuMemCreate failed (out of device memory?)"; return nullptr; }
  CUmemAllocationProp got;
  memset(&got, 0, sizeof got);
  blk.compressed = v.props(&got, blk.handle) == CUDA_SUCCESS && got.allocFlags.compressionType == CU_MEM_ALLOCATION_COMP_GENERIC;
  if (v.reserve(&blk.va, blk.size, 0, 0, 0) != CUDA_SUCCESS) { v.release(blk.handle); *err = "cuMemAddressReserve failed"; return nullptr; }
  if (v.map(blk.va, blk.size, 0, blk.handle, 0) != CUDA_SUCCESS) { v.addr_free(blk.va, blk.size); v.release(blk.handle); *err = "cuMemMap failed"; return nullptr; }
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof acc);
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (v.set_access(blk.va, blk.size, &acc, 1) != CUDA_SUCCESS) {
    v.unmap(blk.va, blk.size); v.addr_free(blk.va, blk.size); v.release(blk.handle);
    *err = "cuMemSetAccess failed";
    return nullptr;
  }
  return (void *)blk.va;
}
static void vmm_free(const vmm_api &v, const vmm_block &blk) {
  v.unmap(blk.va, blk.size);
  v.addr_free(blk.va, blk.size);
  v.release(blk.handle);
}
