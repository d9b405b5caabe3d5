"""Sweep of the pure-store calibration kernel's stream shape (scratch tool): item size x CTAs per SM."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hot_proofs_blake3_circom_b200 as pkg
wc = pkg.builder("blake3_compression", device=0)
nbytes = 48 << 30
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def timed(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
mode = sys.argv[1] if len(sys.argv) > 1 else "lsu"
if mode == "lsu":
    for ctas in (1, 2, 3, 4, 6, 8):
        for slots in (128, 256, 512, 1024, 2048, 4096, 16384):
            wc.set_launch(ctas, slots)
            ms = timed(lambda: wc.calib_fill(d.data_ptr(), nbytes, s, items=True))
            print(json.dumps({"store": "st.global.v8 (LSU)", "ctas_per_sm": ctas, "item_KiB": slots * 32 // 1024, "GBps": round(nbytes / ms / 1e6, 1)}), flush=True)
else:
    for ctas in (1, 2, 4, 6, 8, 16):
        for slots in (64, 128, 256, 512, 1024, 2048, 4096):
            if ctas * slots * 32 > 200 * 1024:
                continue
            wc.set_launch(ctas, slots)
            ms = timed(lambda: wc.calib_fill(d.data_ptr(), nbytes, s, items="bulk"))
            print(json.dumps({"store": "cp.async.bulk (TMA)", "ctas_per_sm": ctas, "item_KiB": slots * 32 // 1024, "GBps": round(nbytes / ms / 1e6, 1)}), flush=True)
