"""compress_sweep.py -- launch shape of the compression witness kernel when it writes COMPRESSIBLE memory (HBM is no longer
the limit there: does the best CTAs-per-SM / items-per-witness change?).  Scratch probe."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200.inputs import lcg_compression_inputs, splitmix_nova_inputs

name = sys.argv[1] if len(sys.argv) > 1 else "blake3_compression"
n = 1 << 16
wc = pkg.builder(name, device=0)
d_in = torch.from_numpy((lcg_compression_inputs if name == "blake3_compression" else splitmix_nova_inputs)(n).view(np.int32)).cuda()
ptr, granted = wc.device_alloc(n * wc.witnessSize * 32, compressible=True)
assert granted
d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream


def timeit(f, reps=6):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for checked in (False, True):
    for cap in (2,):
        for parts in (4, 6, 8, 12, 16, 24):
            wc.set_launch(cap, parts)
            f = (lambda: wc.witness_batch_device_checked(d_in.data_ptr(), n, ptr, d_st.data_ptr(), 0, 0, s)) if checked else \
                (lambda: wc.witness_batch_device(d_in.data_ptr(), n, ptr, d_st.data_ptr(), 0, s))
            t = timeit(f)
            print(json.dumps({"circuit": name, "checked": checked, "ctas_per_sm": cap, "parts": parts, "ms": round(t, 3),
                              "wit_per_s": round(n / t * 1e3)}), flush=True)
