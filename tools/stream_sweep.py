"""Sweep of the launch shape of the witness kernels (scratch tool; device timings with CUDA events).
usage: python tools/stream_sweep.py [circuit] [log2_n]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from hot_proofs_blake3_circom_b200 import _lib
if os.environ.get("B3W_EXP_LIB"):          # experiment builds of the library (e.g. other store cache hints)
    _lib.lib_path = lambda: os.environ["B3W_EXP_LIB"]
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import inputs as gen

circuit = sys.argv[1] if len(sys.argv) > 1 else "blake3_compression"
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 16
checked = len(sys.argv) > 3 and sys.argv[3] == "checked"
n = 1 << logn
wc = pkg.builder(circuit, device=0)
rows = gen.lcg_compression_inputs(n) if circuit == "blake3_compression" else gen.splitmix_nova_inputs(n)
d_in = torch.from_numpy(rows.view(np.int32)).cuda()
ws = wc.witnessSize
wb = ws * 32
d_out = torch.empty(n * wb, dtype=torch.uint8, device="cuda")
d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
d_pub = torch.empty(n * wc.nPublic, dtype=torch.int32, device="cuda")
d_sum = torch.empty(n, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream


def timed(f, reps=5):
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


ref_sum = None
if os.environ.get("B3W_EXP_LIB"):
    configs = [(4, 8), (2, 8), (3, 8)]
else:
    configs = [(c, p) for c in (4, 2, 1) for p in (8, 16, 24, 32, 40, 48, 64, 96)]
if checked:
    configs = [(c, p) for c in (2, 3, 1) for p in (16, 24, 32)]
for rnd in range(2):
    for ctas, parts in configs:
        wc.set_launch(ctas, parts)
        f = lambda: wc.witness_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), d_pub.data_ptr(), s)
        if checked:
            f = lambda: wc.witness_batch_device_checked(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), d_pub.data_ptr(), 0, s)
        if rnd == 0:
            d_out.zero_()
        ms = timed(f)
        if rnd == 0:
            wc.checksum_device(d_out.data_ptr(), n, d_sum.data_ptr(), s)
            torch.cuda.synchronize()
            cs = int(d_sum.sum().item())
            if ref_sum is None:
                ref_sum = cs
            assert cs == ref_sum, "variant changes the witness bytes"
        print(json.dumps({"lib": os.path.basename(_lib.lib_path()), "circuit": circuit + ("+check" if checked else ""), "ctas_per_sm": ctas, "parts": parts, "ms": round(ms, 3),
                          "GBps": round(n * wb / ms / 1e6, 1), "wit_per_s": round(n / ms * 1e3)}), flush=True)
    fill = timed(lambda: wc.calib_fill(d_out.data_ptr(), n * wb, s))
    print(json.dumps({"k_fill_GBps": round(n * wb / fill / 1e6, 1), "bytes": n * wb}), flush=True)
