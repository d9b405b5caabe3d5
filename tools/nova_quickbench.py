"""Quick device timings of the nova kernels (scratch tool)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200.inputs import splitmix_nova_inputs
for name in ("blake3_nova", "blake3_nova_pasta", "blake3_nova_o1"):
    wc = pkg.builder(name, device=0)
    n = 1 << 16
    rows = splitmix_nova_inputs(n)
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_out = torch.empty(n * wc.witnessSize * 32, dtype=torch.uint8, device="cuda")
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_pub = torch.empty(n * 15, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    f = lambda: wc.witness_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), d_pub.data_ptr(), s)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"circuit": name, "n": n, "ms": round(ms, 3), "GBps": round(n * wc.witnessSize * 32 / ms / 1e6, 1), "wit_per_s": round(n / ms * 1e3)}), flush=True)
    del d_out; wc.close()
