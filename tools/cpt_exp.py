"""cpt_exp.py -- times b3w_r1cs_check_device (compact evaluator) with the product library or an experiment build
(B3W_EXP_LIB=<libblake3wit built with -DCPT_EXP=1|2>); scratch tool behind profiles/r01i_r1cs_check.jsonl."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hot_proofs_blake3_circom_b200 import _lib
if os.environ.get("B3W_EXP_LIB"):
    _lib.lib_path = lambda: os.environ["B3W_EXP_LIB"]
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200.inputs import lcg_compression_inputs
os.environ["B3W_STANDALONE_CHECK"] = "compact"
wc = pkg.builder("blake3_compression", device=0)
n = 1 << 15
d_in = torch.from_numpy(lcg_compression_inputs(n).view(np.int32)).cuda()
d_out = torch.empty(n * wc.witnessSize * 32, dtype=torch.uint8, device="cuda")
d_st = torch.empty(n, dtype=torch.uint8, device="cuda"); d_bad = torch.empty(n, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
wc.witness_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), 0, s)
def timeit(f, reps=3):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t = timeit(lambda: wc.r1cs_check_device(d_out.data_ptr(), n, d_st.data_ptr(), d_bad.data_ptr(), s))
print(json.dumps({"lib": os.environ.get("B3W_EXP_LIB", "product"), "ms": round(t, 3), "wit_per_s": round(n / t * 1e3), "read_gbs": round(n * wc.witnessSize * 32 / t / 1e6)}))
