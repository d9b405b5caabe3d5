"""Device and end-to-end timings of the compact (packed) witness path (scratch tool; CUDA events / wall clock)."""
import ctypes as C
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from hot_proofs_blake3_circom_b200 import inputs as gen

L = pkg.lib()


def ev_time(f, reps=5):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, rows_fn in (("blake3_compression", gen.lcg_compression_inputs), ("blake3_nova_pasta", gen.splitmix_nova_inputs)):
    wc = pkg.builder(name, device=0)
    s = torch.cuda.current_stream().cuda_stream
    n = 1 << 20
    rows = rows_fn(n)
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_pk = torch.empty((n, wc.packedWords), dtype=torch.int32, device="cuda")
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_pub = torch.empty(n * wc.nPublic, dtype=torch.int32, device="cuda")
    ms = ev_time(lambda: wc.witness_batch_packed_device(d_in.data_ptr(), n, d_pk.data_ptr(), d_st.data_ptr(), d_pub.data_ptr(), s))
    out = {"circuit": name, "n": n, "packed_bytes": wc.packedWords * 4, "pack_ms": round(ms, 3), "pack_wit_per_s": round(n / ms * 1e3),
           "pack_GBps": round(n * wc.packedWords * 4 / ms / 1e6, 1)}
    m = 1 << 15
    d_out = torch.empty(m * wc.witnessSize * 32, dtype=torch.uint8, device="cuda")
    ms = ev_time(lambda: wc.unpack_device(d_pk.data_ptr(), m, d_out.data_ptr(), s))
    out.update(unpack_n=m, unpack_ms=round(ms, 3), unpack_wit_per_s=round(m / ms * 1e3), unpack_GBps=round(m * wc.witnessSize * 32 / ms / 1e6, 1))
    del d_out
    # end to end through the host-buffer C ABI call: pinned inputs H2D, traces D2H
    h_in = L.b3w_host_alloc(rows.nbytes)
    C.memmove(h_in, rows.ctypes.data, rows.nbytes)
    h_pk = L.b3w_host_alloc(n * wc.packedWords * 4)
    h_st = L.b3w_host_alloc(n)
    h_pub = L.b3w_host_alloc(n * wc.nPublic * 4)
    f = lambda: _lib.check(L.b3w_witness_batch_packed(wc._h, h_in, n, h_pk, h_st, h_pub))
    f()
    t = time.perf_counter()
    for _ in range(3):
        f()
    dt = (time.perf_counter() - t) / 3
    out.update(e2e_packed_wit_per_s=round(n / dt), e2e_d2h_GBps=round(n * wc.packedWords * 4 / dt / 1e9, 1))
    for p in (h_in, h_pk, h_st, h_pub):
        L.b3w_host_free(p)
    print(json.dumps(out), flush=True)
    wc.close()
