"""compress_witness.py -- the witness kernels writing into COMPRESSIBLE device memory (cuMemCreate with
CU_MEM_ALLOCATION_COMP_GENERIC, through cuda-python) against ordinary cudaMalloc memory.  Scratch probe."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cuda import cuda
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200.inputs import lcg_compression_inputs, splitmix_nova_inputs


def ck(r):
    if isinstance(r, tuple):
        err, *rest = r
    else:
        err, rest = r, []
    assert err == cuda.CUresult.CUDA_SUCCESS, err
    return rest[0] if len(rest) == 1 else rest


def alloc_compressible(nbytes):
    prop = cuda.CUmemAllocationProp()
    prop.type = cuda.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = cuda.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = 0
    prop.allocFlags.compressionType = 1          # CU_MEM_ALLOCATION_COMP_GENERIC
    gran = ck(cuda.cuMemGetAllocationGranularity(prop, cuda.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
    size = (nbytes + gran - 1) // gran * gran
    h = ck(cuda.cuMemCreate(size, prop, 0))
    va = ck(cuda.cuMemAddressReserve(size, 0, 0, 0))
    ck(cuda.cuMemMap(va, size, 0, h, 0))
    acc = cuda.CUmemAccessDesc()
    acc.location = prop.location
    acc.flags = cuda.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    ck(cuda.cuMemSetAccess(va, size, [acc], 1))
    return int(va)


def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


torch.cuda.init(); torch.zeros(1, device="cuda")
n = 1 << 16
for name, gen in (("blake3_compression", lcg_compression_inputs), ("blake3_nova_pasta", splitmix_nova_inputs)):
    wc = pkg.builder(name, device=0)
    nbytes = n * wc.witnessSize * 32
    d_in = torch.from_numpy(gen(n).view(np.int32)).cuda()
    d_plain = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    p_comp = alloc_compressible(nbytes)
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_sum = torch.empty((2, n), dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    out = {"circuit": name, "n": n}
    for tag, ptr in (("plain", d_plain.data_ptr()), ("compressible", p_comp)):
        t = timeit(lambda: wc.witness_batch_device(d_in.data_ptr(), n, ptr, d_st.data_ptr(), 0, s))
        tc = timeit(lambda: wc.witness_batch_device_checked(d_in.data_ptr(), n, ptr, d_st.data_ptr(), 0, 0, s))
        tr = timeit(lambda: wc.checksum_device(ptr, n, d_sum[0 if tag == "plain" else 1].data_ptr(), s), reps=3)
        out.update({tag + "_ms": round(t, 3), tag + "_wit_per_s": round(n / t * 1e3), tag + "_write_gbs": round(nbytes / t / 1e6),
                    tag + "_checked_ms": round(tc, 3), tag + "_checksum_read_gbs": round(nbytes / tr / 1e6)})
    assert torch.equal(d_sum[0], d_sum[1]), "different witnesses"
    print(json.dumps(out), flush=True)
    wc.close()
    del d_plain
    torch.cuda.empty_cache()
