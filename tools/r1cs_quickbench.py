"""Quick device timings of the fused-check kernels and the stand-alone HBM checker (scratch tool)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from hot_proofs_blake3_circom_b200 import _lib
if os.environ.get("B3W_EXP_LIB"):          # experiment builds of the library
    _lib.lib_path = lambda: os.environ["B3W_EXP_LIB"]
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200.inputs import lcg_compression_inputs, splitmix_nova_inputs

def timeit(f, reps=3):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for name, gen in (("blake3_compression", lcg_compression_inputs), ("blake3_nova_o1", splitmix_nova_inputs), ("blake3_nova_pasta", splitmix_nova_inputs)):
    wc = pkg.builder(name, device=0)
    n = 1 << 15
    rows = gen(n)
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_out = torch.empty(n * wc.witnessSize * 32, dtype=torch.uint8, device="cuda")
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_bad = torch.empty(n, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    t_plain = timeit(lambda: wc.witness_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), 0, s))
    t_fused = timeit(lambda: wc.witness_batch_device_checked(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), 0, d_bad.data_ptr(), s))
    out = {"circuit": name, "n": n, "plain_ms": round(t_plain, 3), "fused_ms": round(t_fused, 3),
           "plain_wit_per_s": round(n / t_plain * 1e3), "fused_wit_per_s": round(n / t_fused * 1e3)}
    if "pasta" not in name:
        # stand-alone check of the witnesses now resident in HBM: the default evaluator of this circuit's built-in rows, then
        # each evaluator explicitly
        t_hbm = timeit(lambda: wc.r1cs_check_device(d_out.data_ptr(), n, d_st.data_ptr(), d_bad.data_ptr(), s))
        assert int(d_st.max()) == 0
        out.update(hbm_check_ms=round(t_hbm, 3), hbm_check_wit_per_s=round(n / t_hbm * 1e3),
                   hbm_check_read_gbs=round(n * wc.witnessSize * 32 / t_hbm / 1e6))
        for mode in ("warp", "compact", "staged"):
            os.environ["B3W_STANDALONE_CHECK"] = mode
            wc2 = pkg.builder(name, device=0)
            t_m = timeit(lambda: wc2.r1cs_check_device(d_out.data_ptr(), n, d_st.data_ptr(), d_bad.data_ptr(), s))
            assert int(d_st.max()) == 0
            del os.environ["B3W_STANDALONE_CHECK"]
            out.update({"hbm_check_%s_ms" % mode: round(t_m, 3), "hbm_check_%s_wit_per_s" % mode: round(n / t_m * 1e3),
                        "hbm_check_%s_read_gbs" % mode: round(n * wc.witnessSize * 32 / t_m / 1e6)})
            wc2.close()
    print(json.dumps(out), flush=True)
    del d_out; wc.close()
