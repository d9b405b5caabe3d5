"""chunk_sweep.py -- throughput of the host-buffer batch call against b3w_config.chunk (instances per ring slot).
usage: python tools/chunk_sweep.py [log2_n]      (one JSON line per chunk size; out = NULL, status + pub come back)"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import inputs as gen

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << logn
L = pkg.lib()
rows = gen.parallel_rows(gen.splitmix_compression_inputs, n, threads=os.cpu_count() or 4)
L.b3w_host_alloc.restype = C.c_void_p
hin = L.b3w_host_alloc(rows.nbytes)
C.memmove(hin, rows.ctypes.data, rows.nbytes)
hst, hpub = L.b3w_host_alloc(n), L.b3w_host_alloc(n * 64)
for chunk in (0, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
    for byte_check in (False, True):
        wc = pkg.builder("blake3_compression", device=0, chunk=chunk, byte_check=byte_check)
        call = lambda: L.b3w_witness_batch(wc._h, C.c_void_p(hin), C.c_uint64(n), None, C.c_void_p(hst), C.c_void_p(hpub))
        assert call() == 0
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            assert call() == 0
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"chunk": chunk, "byte_check": byte_check, "instances": n, "M_per_s": round(n / dt / 1e6, 3),
                          "ring_GB": round(2 * (chunk or 1024) * wc.witnessSize * 32 / 1e9, 2)}), flush=True)
        wc.close()
