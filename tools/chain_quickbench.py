"""b3w_nova_chain wall-clock vs input size (scratch tool): fixed overhead vs per-step rate."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import inputs as gen
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
wc = pkg.builder("blake3_nova", device=0, chunk=chunk)
for mib in (1, 4, 16, 64):
    data = gen.splitmix_words(0xB3B30003, np.arange(mib << 18, dtype=np.uint64), 1)[:, 0].tobytes()
    res = wc.novaChain(data)
    # time the C ABI call itself with pinned result buffers allocated once (what a long-running host would do)
    import ctypes as C
    from hot_proofs_blake3_circom_b200 import _lib
    from hot_proofs_blake3_circom_b200.witness_calculator import pinned_array
    L = pkg.lib()
    ns = int(res["total_steps"])
    rows, status, pub = pinned_array((ns, 32), np.uint32), pinned_array((ns,), np.uint8), pinned_array((ns, 15), np.uint32)
    h_data = pinned_array((len(data),), np.uint8)
    h_data[:] = np.frombuffer(data, np.uint8)
    root = np.zeros(32, np.uint8)
    call = lambda: _lib.check(L.b3w_nova_chain(wc._h, h_data.ctypes.data, len(data), None, status.ctypes.data, pub.ctypes.data,
                                               rows.ctypes.data, None, root.ctypes.data))
    call()
    torch.cuda.synchronize()
    t = time.perf_counter()
    reps = 5
    for _ in range(reps):
        call()
    dt = (time.perf_counter() - t) / reps
    assert root.tobytes() == res["root"] and not status.any()
    print(json.dumps({"MiB": mib, "ring_chunk": chunk, "steps": int(res["total_steps"]), "ms": round(dt * 1e3, 2),
                      "step_witnesses_per_s": round(res["total_steps"] / dt), "hbm_GBps": round(res["total_steps"] * wc.witnessSize * 32 / dt / 1e9, 1)}), flush=True)
