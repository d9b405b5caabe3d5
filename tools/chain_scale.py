"""chain_scale.py -- b3w_nova_chain on large files (scale check: 64-bit offsets, ring streaming, host result arrays).
usage: python tools/chain_scale.py [MiB ...]   default 128 512.  Checks root == BLAKE3(file), every chunk folds to it, no status."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import blake3
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import inputs as gen

sizes = [int(x) for x in sys.argv[1:]] or [128, 512]
wc = pkg.builder("blake3_nova", device=0, chunk=16384)
for mib in sizes:
    n_words = mib << 18                                  # u32 words
    data = np.concatenate([gen.splitmix_words(0xB3B30003, np.arange(s, min(s + (1 << 24), n_words), dtype=np.uint64), 1)[:, 0]
                           for s in range(0, n_words, 1 << 24)]).tobytes()
    t0 = time.perf_counter()
    res = wc.novaChain(data)
    dt = time.perf_counter() - t0
    digest = blake3.blake3(data, max_threads=blake3.blake3.AUTO).digest()
    ns, off = int(res["total_steps"]), res["step_off"].astype(np.int64)
    last = off[1:] - 1
    folds = bool((res["pub"][last, 2:10].view(np.uint8).reshape(len(last), 32) == np.frombuffer(digest, np.uint8)).all())
    print(json.dumps({"MiB": len(data) >> 20, "chunks": int(res["n_chunks"]), "step_witnesses": ns, "GB_generated": round(ns * wc.witnessSize * 32 / 1e9, 1),
                      "seconds_incl_python_alloc": round(dt, 3), "root_is_blake3": res["root"] == digest, "every_chunk_folds_to_root": folds,
                      "status_any": bool(res["status"].any()), "timing": wc.lastTiming()}), flush=True)
    assert res["root"] == digest and folds and not res["status"].any()
    del res, data
