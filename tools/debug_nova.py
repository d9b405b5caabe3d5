import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hot_proofs_blake3_circom_b200 as pkg
for variant, name in (("nova_bn_o2", "blake3_nova"), ("nova_pasta_o2", "blake3_nova_pasta"), ("nova_bn_o1", "blake3_nova_o1")):
    fx = np.load(os.path.join(ROOT, "tests/golden/%s_cases.npz" % variant))
    wc = pkg.builder(name, device=0)
    res = wc.calculateWitnessBatch(fx["rows"])
    ws = wc.witnessSize
    print(variant, "status", list(res["status"]), list(fx["status"]))
    for i in range(len(fx["rows"])):
        if fx["status"][i]: continue
        a = res["witness"][i].reshape(ws, 32); b = fx["witness"][i].reshape(ws, 32)
        bad = np.nonzero((a != b).any(axis=1))[0]
        if len(bad):
            print(" row", i, "nbad", len(bad), "first", bad[:6], "got", a[bad[0]].view(np.uint32), "want", b[bad[0]].view(np.uint32), "inputs", fx["rows"][i][[0,1,12,13,14]])
    wc.close()
