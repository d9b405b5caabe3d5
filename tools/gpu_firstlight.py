"""First-light GPU script (scratch tool): smoke + quick timings of the witness kernel and the pure-store
calibration kernel at several batch sizes."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as g
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200.inputs import lcg_compression_inputs

print("nproc", os.cpu_count(), flush=True)
os.system("free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv")
t = time.time(); g.smoke(); print("smoke took", time.time() - t, flush=True)

wc = pkg.builder("blake3_compression", device=0)
WS = wc.witnessSize
stream = torch.cuda.current_stream().cuda_stream
for logn in (10, 12, 14, 16):
    n = 1 << logn
    rows = torch.from_numpy(lcg_compression_inputs(n).view(np.int32)).cuda()
    out = torch.empty(n * WS * 32, dtype=torch.uint8, device="cuda")
    pub = torch.empty(n * 16, dtype=torch.int32, device="cuda")
    st = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, fn in (("witness", lambda: wc.witness_batch_device(rows.data_ptr(), n, out.data_ptr(), st.data_ptr(), pub.data_ptr(), stream)),
                     ("fill", lambda: wc.calib_fill(out.data_ptr(), n * WS * 32, stream))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = n * WS * 32 / ms / 1e6
        print(json.dumps({"kernel": name, "n": n, "ms": round(ms, 4), "GBps": round(gbs, 1), "wit_per_s": round(n / ms * 1e3)}), flush=True)
    # verify: re-run witness last, checksum on device vs numpy checksum of host copy of a few instances
    wc.witness_batch_device(rows.data_ptr(), n, out.data_ptr(), st.data_ptr(), pub.data_ptr(), stream)
    sums = torch.empty(n, dtype=torch.int64, device="cuda")
    wc.checksum_device(out.data_ptr(), n, sums.data_ptr(), stream)
    torch.cuda.synchronize()
    sel = [0, 1, n // 2, n - 1]
    w = out.view(n, WS * 32)[sel].cpu().numpy().view(np.uint64)
    e = np.arange(WS * 4, dtype=np.uint64)
    with np.errstate(over="ignore"):
        mix = (e + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        ref = ((w + np.uint64(1)) * mix[None, :]).sum(axis=1, dtype=np.uint64)
    assert (ref == sums[sel].cpu().numpy().view(np.uint64)).all()
    del out
print("firstlight ok")
