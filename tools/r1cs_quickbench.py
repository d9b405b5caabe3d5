"""Device timings of the fused-check kernels and the stand-alone HBM checker, all four circuits, ordinary and compressible
witness buffers.  One JSON line per circuit (profiles/r02_r1cs_check.jsonl)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from hot_proofs_blake3_circom_b200 import _lib
if os.environ.get("B3W_EXP_LIB"):          # experiment builds of the library
    _lib.lib_path = lambda: os.environ["B3W_EXP_LIB"]
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200.inputs import lcg_compression_inputs, splitmix_nova_inputs


def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
only = sys.argv[1:]
for name, gen in (("blake3_compression", lcg_compression_inputs), ("blake3_nova_o1", splitmix_nova_inputs), ("blake3_nova_pasta", splitmix_nova_inputs),
                  ("blake3_nova", splitmix_nova_inputs)):
    if only and name not in only:
        continue
    wc = pkg.builder(name, device=0)
    n = 1 << 15
    rows = gen(n)
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_out = torch.empty(n * wc.witnessSize * 32, dtype=torch.uint8, device="cuda")
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_bad = torch.empty(n, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    nbytes = n * wc.witnessSize * 32
    t_plain = timeit(lambda: wc.witness_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), 0, s))
    t_fused = timeit(lambda: wc.witness_batch_device_checked(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), 0, d_bad.data_ptr(), s))
    out = {"circuit": name, "n": n, "plain_ms": round(t_plain, 3), "fused_ms": round(t_fused, 3),
           "plain_wit_per_s": round(n / t_plain * 1e3), "fused_wit_per_s": round(n / t_fused * 1e3), "program": wc.r1cs_program_info()}
    wc.witness_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), 0, s)
    t_hbm = timeit(lambda: wc.r1cs_check_device(d_out.data_ptr(), n, d_st.data_ptr(), d_bad.data_ptr(), s))
    assert int(d_st.max()) == 0
    out.update(hbm_check_ms=round(t_hbm, 3), hbm_check_wit_per_s=round(n / t_hbm * 1e3), hbm_check_read_gbs=round(nbytes / t_hbm / 1e6),
               hbm_check_frac_of_measured_peak=round(nbytes / t_hbm / 1e6 / peak, 3))
    for per_sm in (2, 3, 5, 6, 8):
        wc.set_launch(per_sm, 0)
        t = timeit(lambda: wc.r1cs_check_device(d_out.data_ptr(), n, d_st.data_ptr(), d_bad.data_ptr(), s))
        out["hbm_check_ms_%dcta" % per_sm] = round(t, 3)
    wc.set_launch(0, 0)
    # the same witnesses in a COMPRESSIBLE buffer: wide coalesced reads of such memory run above the HBM rate
    ptr, granted = wc.device_alloc(nbytes, compressible=True)
    if granted:
        wc.witness_batch_device(d_in.data_ptr(), n, ptr, d_st.data_ptr(), 0, s)
        t_c = timeit(lambda: wc.r1cs_check_device(ptr, n, d_st.data_ptr(), d_bad.data_ptr(), s))
        assert int(d_st.max()) == 0
        out.update(hbm_check_compressible_ms=round(t_c, 3), hbm_check_compressible_wit_per_s=round(n / t_c * 1e3),
                   hbm_check_compressible_read_gbs=round(nbytes / t_c / 1e6))
    wc.device_free(ptr)
    print(json.dumps(out), flush=True)
    del d_out; wc.close()
