"""compress_fill.py -- the pure-store calibration streams (LSU st.global.v8 items, TMA bulk stores) into COMPRESSIBLE
memory: with HBM out of the way, which store path has the higher ceiling?  Scratch probe."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import hot_proofs_blake3_circom_b200 as pkg
wc = pkg.builder("blake3_compression", device=0)
nbytes = 32 << 30
ptr, granted = wc.device_alloc(nbytes, compressible=True)
assert granted
plain = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream


def timed(f, reps=4):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for mem, p in (("compressible", ptr), ("plain", plain.data_ptr())):
    for ctas, slots in ((2, 1024), (2, 2048), (4, 1024), (4, 512), (8, 256)):
        wc.set_launch(ctas, slots)
        ms = timed(lambda: wc.calib_fill(p, nbytes, s, items=True))
        print(json.dumps({"memory": mem, "store": "st.global.v8 (LSU)", "ctas_per_sm": ctas, "item_KiB": slots * 32 // 1024, "GBps": round(nbytes / ms / 1e6, 1)}), flush=True)
    for ctas, slots in ((1, 4096), (2, 2048), (2, 1024), (4, 1024), (4, 512), (8, 512), (8, 256)):
        wc.set_launch(ctas, slots)
        ms = timed(lambda: wc.calib_fill(p, nbytes, s, items="bulk"))
        print(json.dumps({"memory": mem, "store": "cp.async.bulk (TMA)", "ctas_per_sm": ctas, "item_KiB": slots * 32 // 1024, "GBps": round(nbytes / ms / 1e6, 1)}), flush=True)
