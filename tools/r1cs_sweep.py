"""r1cs_sweep.py -- the stand-alone checker across experiment builds of the library (build_exp/*.so: FPK_INFLIGHT / FPK_CTAS_PER_SM)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
from hot_proofs_blake3_circom_b200 import _lib
if os.environ.get("B3W_EXP_LIB"): _lib.lib_path = lambda: os.environ["B3W_EXP_LIB"]
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200.inputs import lcg_compression_inputs, splitmix_nova_inputs
out = {"lib": os.path.basename(os.environ.get("B3W_EXP_LIB", "default"))}
for name, gen in (("blake3_compression", lcg_compression_inputs), ("blake3_nova_pasta", splitmix_nova_inputs), ("blake3_nova_o1", splitmix_nova_inputs)):
    wc = pkg.builder(name, device=0)
    n = 1 << 15
    d_in = torch.from_numpy(gen(n).view(np.int32)).cuda()
    nbytes = n * wc.witnessSize * 32
    d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_bad = torch.empty(n, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    wc.witness_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), 0, s)
    def t(ptr):
        for _ in range(2): wc.r1cs_check_device(ptr, n, d_st.data_ptr(), d_bad.data_ptr(), s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): wc.r1cs_check_device(ptr, n, d_st.data_ptr(), d_bad.data_ptr(), s)
        e1.record(); torch.cuda.synchronize()
        assert int(d_st.max()) == 0
        return e0.elapsed_time(e1) / 5
    ms = t(d_out.data_ptr())
    ptr, granted = wc.device_alloc(nbytes, compressible=True)
    wc.witness_batch_device(d_in.data_ptr(), n, ptr, d_st.data_ptr(), 0, s)
    ms_c = t(ptr)
    out[name] = {"ms": round(ms, 3), "M_per_s": round(n / ms / 1e3, 2), "read_gbs": round(nbytes / ms / 1e6), "compressible_ms": round(ms_c, 3), "compressible_M_per_s": round(n / ms_c / 1e3, 2)}
    wc.device_free(ptr); del d_out; wc.close()
print(json.dumps(out), flush=True)
''' % ROOT
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "build_exp", "libb3w_*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["B3W_EXP_LIB"] = lib
    subprocess.run([sys.executable, "-c", code], env=env)
