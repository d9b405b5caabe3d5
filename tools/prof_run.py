"""prof_run.py -- minimal driver for ncu captures of the witness kernel (no timing claims).
usage: python tools/prof_run.py [log2_n] [launches] [circuit] [checked|plain|r1cs] [compressible]
`r1cs`: one witness launch, then `launches` launches of the stand-alone checker k_r1cs_check_fast on those witnesses."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200.inputs import lcg_compression_inputs

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 5
circuit = sys.argv[3] if len(sys.argv) > 3 else "blake3_compression"
checked = len(sys.argv) > 4 and sys.argv[4] == "checked"
n = 1 << logn
wc = pkg.builder(circuit, device=0)
if circuit == "blake3_compression":
    rows = lcg_compression_inputs(n)
else:
    from hot_proofs_blake3_circom_b200.inputs import splitmix_nova_inputs
    rows = splitmix_nova_inputs(n)
d_in = torch.from_numpy(rows.view(np.int32)).cuda()
if len(sys.argv) > 5 and sys.argv[5] == "compressible":
    out_ptr, granted = wc.device_alloc(n * wc.witnessSize * 32, compressible=True)
    assert granted
else:
    d_out = torch.empty(n * wc.witnessSize * 32, dtype=torch.uint8, device="cuda")
    out_ptr = d_out.data_ptr()
d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
d_pub = torch.empty(n * wc.nPublic, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
if len(sys.argv) > 4 and sys.argv[4] == "r1cs":
    wc.witness_batch_device(d_in.data_ptr(), n, out_ptr, d_st.data_ptr(), d_pub.data_ptr(), s)
    d_bad = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(launches):
        wc.r1cs_check_device(out_ptr, n, d_st.data_ptr(), d_bad.data_ptr(), s)
    torch.cuda.synchronize()
    assert int(d_st.max()) == 0
    launches = 0
for _ in range(launches):
    if checked:
        wc.witness_batch_device_checked(d_in.data_ptr(), n, out_ptr, d_st.data_ptr(), d_pub.data_ptr(), 0, s)
    else:
        wc.witness_batch_device(d_in.data_ptr(), n, out_ptr, d_st.data_ptr(), d_pub.data_ptr(), s)
torch.cuda.synchronize()
print("prof_run done", n, launches)
