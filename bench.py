#!/usr/bin/env python3
"""bench.py -- BASELINE.json metric: blake3_compression witnesses/sec (and witness HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: BASELINE config 2 = 2^16 blake3_compression instances
(the LCG(6429) genRandomChunk sequence, instance 0 = the reference's golden input) PER GPU, witnesses written
in .wtns body layout into a 50.5 GB HBM buffer.  Instances are independent, so N GPUs = N disjoint index
ranges, no collective on the data path (weak scaling; NCCL is used only for the barrier and the max-over-ranks
of the timings).

Keys of the JSON line (rank 0):
  value      whole-job witnesses/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  roofline   the witness kernel against the measured HBM peak (MEASURED_PEAKS.json); algorithmic bytes =
             32*24093 written + 112 read per witness
  e2e        the same metric through the C ABI b3w_witness_batch() with HOST (pinned) buffers: H2D of the
             inputs and D2H of every witness byte + status + public outputs inside the timed region
  e2e_compact  ditto with out=NULL: witnesses only stream through the HBM ring, compact results come back
  e2e_packed   ditto with every witness returned in compact form (its 3 776-byte trace)
  cpu_baseline the reference's own wasm witness program (oracle/_ref, translated to C) on all host cores,
             bounded sample (rank 0, N=1 only)
--impl reference times that CPU path as its own arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WS = 24093
WIT_BYTES = WS * 32                 # 770 976 B written per witness
IN_BYTES = 28 * 4                   # 112 B read per witness
LOG2_BATCH = 16
METRIC = "blake3_compression witnesses/sec"
print_json = print
WORKLOAD = "blake3_compression batch 2^16 random 64-byte blocks, BN254 Fr (BASELINE configs[1]), per GPU"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, read+write bytes; burst)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own witness program on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference_rate(n_witnesses, nthreads, first=0):
    """Times Oracle A (oracle/_ref: the reference .wasm translated to C, witness_calculator.js protocol incl.
    input set-up and read-out) on `nthreads` threads.  Falls back to the C port only if _ref is not there."""
    from hot_proofs_blake3_circom_b200.inputs import lcg_compression_inputs
    from oracle import ref_wasm
    rows = lcg_compression_inputs(n_witnesses, first=first)
    if ref_wasm.available("compression"):
        ref = ref_wasm.RefWasm("compression")
        _, status, secs = ref.batch_u32(rows, nthreads=nthreads, want_out=False)
        assert (status == 0).all()
        return n_witnesses / secs, secs, "reference"
    from oracle import port
    t = time.perf_counter()
    port.witness_batch("compression", rows, nthreads=nthreads, want="sums")
    secs = time.perf_counter() - t
    return n_witnesses / secs, secs, "port"


def run_reference(args, rank, world):
    if rank != 0:
        return
    ncpu = os.cpu_count() or 1
    per_step = 4 * ncpu                      # ~1.2 s of wall clock per step at ~3.5 witnesses/s/core
    for i in range(args.warmup):
        cpu_reference_rate(ncpu, ncpu, first=i * ncpu)
    t_total, kind = 0.0, "reference"
    for k in range(args.steps):
        _, secs, kind = cpu_reference_rate(per_step, ncpu, first=1000 + k * per_step)
        t_total += secs
    value = per_step * args.steps / t_total
    sample = "%d witnesses per step (4 per host thread) of the LCG(6429) sequence, %d steps" % (per_step, args.steps)
    print_json(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "witnesses/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 + Fr256 (BN254)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "note": "CPU arm: bounded sample per step; "
                                        "the reference wasm (V8 unavailable) translated to C, all host threads"},
        "cpu_baseline": {"value": value, "unit": "witnesses/s", "cores": ncpu, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "witnesses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------
def run_own(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import hot_proofs_blake3_circom_b200 as pkg
    from hot_proofs_blake3_circom_b200.inputs import lcg_compression_inputs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = 1 << LOG2_BATCH
    first = rank * n                                            # shard = contiguous index range
    wc = pkg.builder("blake3_compression", device=local_rank, chunk=2048)
    rows = lcg_compression_inputs(n, first=first)
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_out = torch.empty(n * WIT_BYTES, dtype=torch.uint8, device="cuda")       # ordinary (cudaMalloc) memory
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_pub = torch.empty(n * 16, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    # The witnesses of the timed steps go to COMPRESSIBLE device memory (b3w_device_alloc: Blackwell's L2 compresses such
    # lines on their way to HBM; a witness is mostly zero bytes).  Same bytes on read-back; the plain-memory rate is
    # measured next to it.  --plain-output, or a device that does not grant compression, keeps ordinary memory.
    out_ptr, compressible = d_out.data_ptr(), False
    if not args.plain_output:
        try:
            p_c, granted = wc.device_alloc(n * WIT_BYTES, compressible=True)
            if granted:
                out_ptr, compressible = p_c, True
            else:
                wc.device_free(p_c)
        except pkg.B3WError:
            pass

    if world > 1:                                               # every rank takes the same path (collectives follow)
        flag = torch.tensor([1 if compressible else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if compressible and int(flag.item()) == 0:
            wc.device_free(out_ptr)
            out_ptr, compressible = d_out.data_ptr(), False

    def step(ptr=None):
        wc.witness_batch_device(d_in.data_ptr(), n, ptr or out_ptr, d_st.data_ptr(), d_pub.data_ptr(), stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # --- timed region: exactly K steps, one kernel launch each -------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    torch.cuda.synchronize()
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[-1]))
    launch_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    kernel_ms = max_over_ranks(sum(launch_ms) / len(launch_ms))
    gpu_launches = world * args.steps                         # one witness kernel per step per rank
    assert int(d_st.max()) == 0
    pub0 = d_pub[:16].cpu().numpy().view(np.uint32)

    # --- the same kernel into ordinary memory (what every figure before r01j was measured on) -------
    plain_ms = None
    if compressible:
        for _ in range(3):
            step(d_out.data_ptr())
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(5):
            step(d_out.data_ptr())
        p1.record()
        torch.cuda.synchronize()
        plain_ms = max_over_ranks(p0.elapsed_time(p1) / 5)
        wc.device_free(out_ptr)

    # --- pure-store calibration on the same buffer (the write roofline of this very GPU) -----------
    for _ in range(2):
        wc.calib_fill(d_out.data_ptr(), n * WIT_BYTES, stream)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(5):
        wc.calib_fill(d_out.data_ptr(), n * WIT_BYTES, stream)
    c1.record()
    torch.cuda.synchronize()
    fill_gbs = 5 * n * WIT_BYTES / c0.elapsed_time(c1) / 1e6
    for _ in range(2):
        wc.calib_fill(d_out.data_ptr(), n * WIT_BYTES, stream, items=True)
    torch.cuda.synchronize()
    c0.record()
    for _ in range(5):
        wc.calib_fill(d_out.data_ptr(), n * WIT_BYTES, stream, items=True)
    c1.record()
    torch.cuda.synchronize()
    fill_items_gbs = 5 * (n * WIT_BYTES // 32768 * 32768) / c0.elapsed_time(c1) / 1e6
    del d_out
    torch.cuda.empty_cache()

    # --- e2e: the C ABI host-buffer call (what the N-API addon / a user calls) ---------------------
    L = pkg.lib()
    avail = 0
    with open("/proc/meminfo") as f:
        for line in f:
            if line.startswith("MemAvailable"):
                avail = int(line.split()[1]) * 1024
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    budget = int(avail * 0.5 / max(local_world, 1))
    n_e2e = n
    while n_e2e * WIT_BYTES > budget and n_e2e > 1024:
        n_e2e //= 2
    h_out = L.b3w_host_alloc_near(n_e2e * WIT_BYTES, local_rank)       # pinned, on the GPU's own NUMA node
    h_in = L.b3w_host_alloc_near(n_e2e * IN_BYTES, local_rank)
    h_st = L.b3w_host_alloc_near(n_e2e, local_rank)
    h_pub = L.b3w_host_alloc_near(n_e2e * 64, local_rank)
    if not (h_out and h_in and h_st and h_pub):
        raise SystemExit("bench.py: pinned host allocation failed: " + L.b3w_last_error().decode())
    import ctypes as C
    C.memmove(h_in, rows.ctypes.data, n_e2e * IN_BYTES)
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_run(out_ptr, steps, calc=None):
        from hot_proofs_blake3_circom_b200 import _lib
        h = (calc or wc)._h
        for _ in range(1):
            _lib.check(L.b3w_witness_batch(h, h_in, n_e2e, out_ptr, h_st, h_pub))
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            _lib.check(L.b3w_witness_batch(h, h_in, n_e2e, out_ptr, h_st, h_pub))   # returns after the last D2H
        dt = time.perf_counter() - t0
        barrier()
        return max_over_ranks(dt) / steps

    t_full = e2e_run(h_out, e2e_steps)
    got0 = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_uint32)), shape=(WS * 8,))
    assert got0[8] == pub0[0] and got0[0] == 1                  # slot 0 == 1, slot 1 == out[0]
    t_compact = e2e_run(None, e2e_steps)
    # the same two calls with the library's HBM ring in compressible memory (B3W_FLAG_COMPRESSIBLE_RING)
    t_full_c = t_compact_c = None
    if compressible:
        wc_ring = pkg.builder("blake3_compression", device=local_rank, chunk=2048, compressible_ring=True)
        t_compact_c = e2e_run(None, e2e_steps, wc_ring)
        t_full_c = e2e_run(h_out, 2, wc_ring)
        assert got0[8] == pub0[0] and got0[0] == 1
        wc_ring.close()
    L.b3w_host_free(h_out)
    # ... and with the witnesses returned in COMPACT form (the per-instance trace, 3 776 B: every slot is a pure function
    # of it; b3w_unpack_device expands on demand)
    pk_words = wc.packedWords
    h_pk = L.b3w_host_alloc_near(n_e2e * pk_words * 4, local_rank)
    if not h_pk:
        raise SystemExit("bench.py: pinned host allocation failed: " + L.b3w_last_error().decode())

    def packed_run(steps):
        from hot_proofs_blake3_circom_b200 import _lib
        _lib.check(L.b3w_witness_batch_packed(wc._h, h_in, n_e2e, h_pk, h_st, h_pub))
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            _lib.check(L.b3w_witness_batch_packed(wc._h, h_in, n_e2e, h_pk, h_st, h_pub))
        dt = time.perf_counter() - t0
        barrier()
        return max_over_ranks(dt) / steps

    t_packed = packed_run(e2e_steps)
    pk0 = np.ctypeslib.as_array(C.cast(h_pk, C.POINTER(C.c_uint32)), shape=(pk_words,))
    assert pk0[1] == 1 and pk0[30] == pub0[0]                   # trace word 1 = the constant 1, word 30 = out[0]
    for p in (h_pk, h_in, h_st, h_pub):
        L.b3w_host_free(p)

    # --- cpu baseline (rank 0, N=1 only): bounded sample on all host cores -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ncpu = os.cpu_count() or 1
        cpu_reference_rate(ncpu, ncpu)                          # warm-up, discarded
        n_s = 6 * ncpu
        rate, secs, kind = cpu_reference_rate(n_s, ncpu, first=ncpu)
        cpu = {"value": rate, "unit": "witnesses/s", "cores": ncpu, "kind": kind,
               "sample": "%d witnesses (6 per host thread, %d threads) of the same LCG(6429) workload incl. input set-up "
                         "and read-out, %.1f s wall; reference wasm translated to C (no V8 here)" % (n_s, ncpu, secs)}

    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    alg_bytes = n * (WIT_BYTES + IN_BYTES)
    achieved = alg_bytes / kernel_ms / 1e6
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("k_blake3_comp_witness_dram_bytes_per_launch" + ("_compressible" if compressible else ""))
    value = world * n * args.steps / (total_ms / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": "witnesses/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 + Fr256 (BN254)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "instances_per_gpu": n, "witness_bytes": WIT_BYTES,
                   "hbm_out_bytes_per_gpu": n * WIT_BYTES, "l2": "each step writes 50.5 GB per GPU, >> 126 MB L2",
                   "sharding": "contiguous index ranges, no collective", "out0_instance0": int(pub0[0]),
                   "output_memory": "compressible device memory (b3w_device_alloc, CU_MEM_ALLOCATION_COMP_GENERIC): the L2 compresses "
                                    "witness lines on their way to HBM; identical bytes on read-back" if compressible
                                    else "ordinary device memory (cudaMalloc)"},
        "roofline": {"bound": "hbm", "kernel": "k_blake3_comp_witness", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src + " (of measured)"
                     if "MEASURED" in peak_src else peak_src + " (of fallback)",
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms,
                     "pure_store_fill_gbs_same_gpu": fill_gbs,
                     "pure_store_same_stream_shape_gbs": fill_items_gbs, "frac_of_spec_8TBps": achieved / 8000.0,
                     "note": ("witness bytes per second INTO COMPRESSIBLE memory: the HBM interface moves fewer bytes than the "
                              "kernel writes (traffic = ncu dram bytes per launch), so achieved can exceed the interface's peak; "
                              "achieved_plain_memory is the same kernel into cudaMalloc memory") if compressible else None,
                     "achieved_plain_memory": (alg_bytes / plain_ms / 1e6) if plain_ms else None},
        "e2e": {"value": world * n_e2e / t_full, "unit": "witnesses/s", "h2d_bytes_per_step": n_e2e * IN_BYTES,
                "d2h_bytes_per_step": n_e2e * (WIT_BYTES + 1 + 64), "instances_per_step_per_gpu": n_e2e,
                "ms_per_step": 1e3 * t_full, "d2h_gbs_per_gpu": n_e2e * WIT_BYTES / t_full / 1e9,
                "api": "b3w_witness_batch(host pinned in/out): every witness byte copied to the host"},
        "e2e_compact": {"value": world * n_e2e / t_compact, "unit": "witnesses/s", "h2d_bytes_per_step": n_e2e * IN_BYTES,
                        "d2h_bytes_per_step": n_e2e * (1 + 64), "ms_per_step": 1e3 * t_compact,
                        "api": "b3w_witness_batch(out=NULL): witnesses stream through the HBM ring, status + out[16] return"},
        "e2e_compressible_ring": None if t_compact_c is None else {
            "full_copy": world * n_e2e / t_full_c, "compact": world * n_e2e / t_compact_c, "unit": "witnesses/s",
            "api": "the two calls above on a context created with B3W_FLAG_COMPRESSIBLE_RING"},
        "e2e_packed": {"value": world * n_e2e / t_packed, "unit": "witnesses/s", "h2d_bytes_per_step": n_e2e * IN_BYTES,
                       "d2h_bytes_per_step": n_e2e * (pk_words * 4 + 1 + 64), "ms_per_step": 1e3 * t_packed,
                       "api": "b3w_witness_batch_packed(host pinned in/out): every witness returned in compact form "
                              "(%d B trace per instance; expandable to the .wtns body with b3w_unpack_device)" % (pk_words * 4)},
        "gpu_launches": gpu_launches, "clocks": clocks}
    if plain_ms:
        line["value_plain_memory"] = world * n / (plain_ms / 1e3)
    if cpu:
        line["cpu_baseline"] = cpu
    print_json(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--plain-output", action="store_true", help="timed steps write ordinary (cudaMalloc) memory")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself as one rank per GPU
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                                   "--master-port", "29517"] + sys.argv)
    # stdout carries exactly ONE line, the JSON: anything a library prints there (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(real_stdout, "w")
    global print_json
    print_json = lambda line: (json_out.write(line + "\n"), json_out.flush())
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_own(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
