"""wide_fuzz.py -- long differential run of the field-element input paths against Oracle B (scratch tool, GPU box).
usage: python tools/wide_fuzz.py [instances per circuit] [seed]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib.util
import numpy as np
import hot_proofs_blake3_circom_b200 as pkg
from oracle import port


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
tw, tn = load("test_gpu_wide"), load("test_gpu_nova_wide")
jobs = [("blake3_compression", "compression", lambda p: tw.random_wide_rows(n, seed, 0.15))] + \
       [(nm, v, lambda p: tn.random_rows(n, p, seed)) for nm, v in tn.NOVA]
for name, variant, make in jobs:
    wc = pkg.builder(name, device=0, chunk=1024)
    vals = make(wc.prime)
    t = time.time()
    res = wc.calculateWitnessBatchFr(vals)
    t_gpu = time.time() - t
    bad = 0
    n_ok = 0
    for i, v in enumerate(vals):
        rc, w = port.witness_fr(variant, [x % wc.prime for x in v])
        if rc != res["status"][i] or (rc == 0 and not np.array_equal(res["witness"][i], w)):
            bad += 1
            if bad <= 3:
                print("MISMATCH", name, i, rc, int(res["status"][i]), v, flush=True)
        n_ok += rc == 0
    print(json.dumps({"circuit": name, "instances": n, "seed": seed, "valid": int(n_ok), "asserting": n - int(n_ok), "mismatches": bad,
                      "gpu_seconds_incl_conversion": round(t_gpu, 2)}), flush=True)
    wc.close()
