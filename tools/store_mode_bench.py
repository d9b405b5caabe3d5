"""store_mode_bench.py -- the staged-tile TMA store experiment (b3w_debug_set_store_mode): the plain blake3_compression
kernel with direct 256-bit stores (mode 0) vs shared-memory tiles + cp.async.bulk (mode 1), 2^16 instances, into ordinary
and compressible memory, next to the two pure-store ceilings.  One JSON line (profiles/r02_store_mode.jsonl)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200.inputs import lcg_compression_inputs

n = 1 << 16
wc = pkg.builder("blake3_compression", device=0)
rows = lcg_compression_inputs(n)
d_in = torch.from_numpy(rows.view(np.int32)).cuda()
nbytes = n * wc.witnessSize * 32
d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
ptr_c, granted = wc.device_alloc(nbytes, compressible=True)
d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
d_sum = [torch.empty(n, dtype=torch.int64, device="cuda") for _ in range(2)]
s = torch.cuda.current_stream().cuda_stream


def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"n": n, "witness_bytes": wc.witnessSize * 32, "compressible_granted": granted}
for mem, ptr in (("plain", d_out.data_ptr()), ("compressible", ptr_c)):
    for mode in (0, 1):
        for parts in ((24, 12) if mode == 0 else (24, 12, 8, 6)):
            wc.set_store_mode(mode)
            wc.set_launch(0, parts)
            t = timeit(lambda: wc.witness_batch_device(d_in.data_ptr(), n, ptr, d_st.data_ptr(), 0, s))
            wc.checksum_device(ptr, n, d_sum[mode].data_ptr(), s)
            torch.cuda.synchronize()
            out["%s_mode%d_parts%d_ms" % (mem, mode, parts)] = round(t, 4)
            out["%s_mode%d_parts%d_gbs" % (mem, mode, parts)] = round(nbytes / t / 1e6)
    assert torch.equal(d_sum[0], d_sum[1]), "store modes disagree"
    wc.set_store_mode(0)
    wc.set_launch(0, 0)
    for items, nm in ((True, "fill_items"), ("bulk", "fill_bulk")):
        t = timeit(lambda: wc.calib_fill(ptr, nbytes, s, items=items), reps=5)
        out["%s_%s_gbs" % (mem, nm)] = round(nbytes // 32768 * 32768 / t / 1e6)
print(json.dumps(out), flush=True)
