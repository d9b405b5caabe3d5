"""bench_configs.py -- measurements for the BASELINE configs that bench.py's single JSON line does not carry
(configs[0], configs[2..4]); one JSON line each, written for profiles/.  Not the driver's benchmark (that is bench.py)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from hot_proofs_blake3_circom_b200 import inputs as gen

L = pkg.lib()


def pinned(arr):
    p = L.b3w_host_alloc(arr.nbytes)
    C.memmove(p, arr.ctypes.data, arr.nbytes)
    return p


def timed(f, reps=3, warm=1):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


def config1():
    """BASELINE configs[0]: ONE witness from inputs/blake3_compression (the LCG(6429) golden input) -- what
    `node generate_witness.js` does in the reference (0.25-0.35 s in its wasm).  Latency of the C ABI call and of the
    calculateWTNSBin wrapper, plus the CPU reference (Oracle A) on the same input, one thread."""
    wc = pkg.builder("blake3_compression", device=0)
    row = gen.lcg_compression_inputs(1)[0]
    inp = {"h": [int(x) for x in row[0:8]], "m": [int(x) for x in row[8:24]], "t": [0, 0], "b": 64, "d": 0}
    golden = np.load(os.path.join(ROOT, "tests", "golden", "compression_golden.npz"))
    assert wc.calculateWTNSBin(inp, 0).tobytes() == golden["wtns"].tobytes()
    out = np.empty(wc.witnessSize * 32, np.uint8)
    from hot_proofs_blake3_circom_b200.witness_calculator import pinned_array
    out_p = pinned_array((wc.witnessSize * 32,), np.uint8)

    def lat(f, reps=300):
        for _ in range(20):
            f()
        ts = []
        for _ in range(reps):
            t = time.perf_counter()
            f()
            ts.append(time.perf_counter() - t)
        return float(np.median(ts)), float(np.percentile(ts, 99))
    abi = lat(lambda: _lib.check(L.b3w_witness_one(wc._h, row.ctypes.data, out.ctypes.data)))
    abi_p = lat(lambda: _lib.check(L.b3w_witness_one(wc._h, row.ctypes.data, out_p.ctypes.data)))
    api = lat(lambda: wc.calculateWTNSBin(inp, 0), reps=100)
    line = {"config": "configs[0]: blake3_compression single witness (golden input), 1 B200", "witness_bytes": wc.witnessSize * 32,
            "b3w_witness_one_median_us": abi[0] * 1e6, "b3w_witness_one_p99_us": abi[1] * 1e6,
            "b3w_witness_one_pinned_out_median_us": abi_p[0] * 1e6,
            "calculateWTNSBin_python_median_us": api[0] * 1e6,
            "note": "host row in, 770 976-byte witness out (pageable / pinned buffer); byte-identical to the reference's witness.wtns"}
    from oracle import ref_wasm
    if ref_wasm.available("compression"):
        ref = ref_wasm.RefWasm("compression")
        _, st, secs = ref.batch_u32(gen.lcg_compression_inputs(4), nthreads=1, want_out=True)
        line["reference_wasm_one_thread_ms"] = secs / 4 * 1e3
    print(json.dumps(line), flush=True)
    wc.close()


def config3():
    """BASELINE configs[2]: all Nova step witnesses of a synthetic 1 MiB file through b3w_nova_chain; the C ABI call is
    timed with pinned result buffers allocated once (what a long-running host does)."""
    from hot_proofs_blake3_circom_b200.witness_calculator import pinned_array
    wc = pkg.builder("blake3_nova", device=0, chunk=8192)
    data = gen.splitmix_words(0xB3B30003, np.arange(1 << 18, dtype=np.uint64), 1)[:, 0].tobytes()
    res = wc.novaChain(data)
    import blake3
    assert res["root"] == blake3.blake3(data).digest() and (res["status"] == 0).all()
    ns = int(res["total_steps"])
    rows, status, pub = pinned_array((ns, 32), np.uint32), pinned_array((ns,), np.uint8), pinned_array((ns, 15), np.uint32)
    h_data = pinned_array((len(data),), np.uint8)
    h_data[:] = np.frombuffer(data, np.uint8)
    root = np.zeros(32, np.uint8)
    f = lambda: _lib.check(L.b3w_nova_chain(wc._h, h_data.ctypes.data, len(data), None, status.ctypes.data, pub.ctypes.data,
                                            rows.ctypes.data, None, root.ctypes.data))
    dt = timed(f, reps=5)
    assert root.tobytes() == res["root"] and np.array_equal(rows, res["rows"]) and np.array_equal(pub, res["pub"])
    print(json.dumps({"config": "configs[2]: blake3_nova (BN254, O2) chained-chunk witnesses, synthetic 1 MiB input",
                      "chunks": res["n_chunks"], "step_witnesses": ns, "witness_bytes": wc.witnessSize * 32,
                      "seconds": dt, "step_witnesses_per_s": ns / dt, "GB_generated": ns * wc.witnessSize * 32 / 1e9,
                      "hbm_write_GBps": ns * wc.witnessSize * 32 / dt / 1e9,
                      "note": "b3w_nova_chain end to end from host bytes: H2D, device BLAKE3 tree, chain rows, all step witnesses "
                              "through the HBM ring, z_{i+1}/status/rows D2H (pinned); root == blake3(file)"}), flush=True)
    wc.close()


def config4(ring=False):
    n = 1 << 20
    wc = pkg.builder("blake3_nova_pasta", device=0, chunk=32768, compressible_ring=ring)
    rows = gen.splitmix_nova_inputs(n)
    h_in = pinned(rows)
    h_st = L.b3w_host_alloc(n)
    h_pub = L.b3w_host_alloc(n * 60)
    f = lambda: _lib.check(L.b3w_witness_batch(wc._h, h_in, n, None, h_st, h_pub))
    dt = timed(f, reps=3)
    st = np.ctypeslib.as_array(C.cast(h_st, C.POINTER(C.c_uint8)), shape=(n,))
    assert not st.any()
    print(json.dumps({"config": "configs[3]: blake3_nova_pasta (Pallas Fr) batch 2^20, streamed through the HBM ring" +
                      (" (ring in compressible memory)" if ring else ""),
                      "instances": n, "witness_bytes": wc.witnessSize * 32, "seconds": dt, "witnesses_per_s": n / dt,
                      "GB_generated": n * wc.witnessSize * 32 / 1e9, "hbm_write_GBps": n * wc.witnessSize * 32 / dt / 1e9,
                      "note": "b3w_witness_batch(out=NULL): host pinned inputs H2D, 781 GB of witnesses written to a 2-slot HBM ring, "
                              "status + z_{i+1} D2H"}), flush=True)
    for p in (h_in, h_st, h_pub):
        L.b3w_host_free(p)
    wc.close()


def config5(ring=False):
    """BASELINE configs[4]: 2^24 compression instances sharded by contiguous index range over every visible GPU through
    the C ABI's multi-GPU entry point (one host thread + context per device, no collective), without / with the fused check."""
    n = 1 << 24
    for fused in (False, True):
        m = pkg.MultiGpuCalculator("blake3_compression", devices=None, chunk=32768, fused_check=fused, compressible_ring=ring)
        rows = gen.splitmix_compression_inputs(n)
        h_in = pinned(rows)
        del rows
        h_st = L.b3w_host_alloc(n)
        h_pub = L.b3w_host_alloc(n * 64)
        f = lambda: m.witness_batch_host(h_in, n, None, h_st, h_pub)
        dt = timed(f, reps=2)
        st = np.ctypeslib.as_array(C.cast(h_st, C.POINTER(C.c_uint8)), shape=(n,))
        assert not st.any()
        pub = np.ctypeslib.as_array(C.cast(h_pub, C.POINTER(C.c_uint32)), shape=(n, 16))
        print(json.dumps({"config": "configs[4]: blake3_compression 2^24 instances, %s, %d B200 (contiguous shards, no collective)%s"
                          % ("fused on-device R1CS check" if fused else "no check", m.nDevices,
                             ", rings in compressible memory" if ring else ""),
                          "n_gpus": m.nDevices, "instances": n, "seconds": dt, "witnesses_per_s": n / dt,
                          "TB_generated": n * 770976 / 1e12, "hbm_write_GBps_per_gpu": n * 770976 / dt / 1e9 / m.nDevices,
                          "xor_of_out0": int(np.bitwise_xor.reduce(pub[:, 0])),
                          "note": "b3w_multi_witness_batch(out=NULL) with splitmix inputs (random h,t,d, ragged b): 12.9 TB streamed "
                                  "through the HBM rings, status + out[16] D2H into the shared pinned arrays"}), flush=True)
        for p in (h_in, h_st, h_pub):
            L.b3w_host_free(p)
        m.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "3", "4", "5"]
    if "1" in which:
        config1()
    if "3" in which:
        config3()
    if "4" in which:
        config4()
    if "5" in which:
        config5()
    if "4c" in which:
        config4(ring=True)
    if "5c" in which:
        config5(ring=True)
