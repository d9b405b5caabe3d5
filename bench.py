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
  config     the workload, identical in both arms; `run` = what this run observed (output memory kind, out[0] of instance 0)
  value      whole-job witnesses/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks; the
             witnesses go to compressible device memory, the library's default (run.output_memory)
  roofline   the HBM record: the same kernel into ORDINARY memory, timed per launch in this run, against the measured
             HBM peak (MEASURED_PEAKS.json); algorithmic bytes = 32*24093 written + 112 read per witness
  roofline_compressible  the timed region itself against its own bound, the SM-side store path (peak = the kernel's
             store stream with no work, measured in this run into the same buffer)
  value_sustained / value_checked  the timed launch for >= 1 s back to back / with the fused R1CS check
  r1cs_check_resident  the stand-alone check of witnesses read back from HBM (all 24 544 rows)
  e2e        the same metric through the C ABI b3w_witness_batch() with HOST (pinned) buffers: H2D of the
             inputs and D2H of every witness byte + status + public outputs inside the timed region
  e2e_compact  ditto with out=NULL: witnesses only stream through the HBM ring, compact results come back
  e2e_packed   ditto with every witness returned in compact form (its 3 776-byte trace)
  e2e_hybrid   every .wtns body in host memory like e2e, but packed records over PCIe + expansion on host threads
  config4 / config5  BASELINE configs[3] / configs[4] through the C ABI, streamed, >= 1 s each (config 5: fused check,
             per-instance witness checksums, a 1 024-instance sample of full witnesses verified against those checksums;
             sums_vs_oracle_b: ALL 2^24 checksums against the committed digests of Oracle B's, tests/golden/)
  fr_batches   b3w_witness_batch_fr (Fr256 rows, all-u32 and 1 % field-valued) next to b3w_witness_batch (N = 1 only)
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


def workload_config(n=1 << LOG2_BATCH):
    """`config` of the JSON line: the workload and nothing measured, so that both arms (this one and --impl reference)
    print the SAME object; what a run found out (output memory kind, instance 0's out[0], checksums) is under `run`."""
    return {"workload": WORKLOAD, "instances_per_gpu": n, "witness_bytes": WIT_BYTES, "input_bytes": IN_BYTES,
            "witness_bytes_per_batch_per_gpu": n * WIT_BYTES,
            "l2": "a batch's witnesses are 50.5 GB per GPU, >> 126 MB L2: nothing of one step survives in L2 for the next",
            "sharding": "contiguous index ranges, no collective"}


def sums_vs_oracle_fixture(sums, first, fixture="compression_sums_2p24.npz"):
    """Every per-instance witness checksum of a streamed run (instances first .. first + len(sums) of the circuit's splitmix
    sequence: config 5's blake3_compression, config 4's blake3_nova_pasta) against Oracle B, through the per-4096-instance digests of Oracle B's checksums committed under
    tests/golden/ (made by tests/golden/make_golden_sums.py; reading a fixture is not running the oracle).  Never raises:
    -> {"match": bool, "blocks": k, ...} or {"skipped": why}."""
    try:
        import hashlib
        path = os.path.join(ROOT, "tests", "golden", fixture)
        if not os.path.exists(path):
            return {"skipped": "tests/golden/%s not present" % fixture}
        g = np.load(path)
        blk, want = int(g["block"]), g["block_digest"]
        n = int(sums.size)
        if first % blk or n % blk or n == 0 or (first + n) // blk > want.size:
            return {"skipped": "instances %d..%d are not whole blocks of the fixture (2^%d instances)" % (first, first + n, int(g["log2_n"]))}
        s = np.ascontiguousarray(sums, "<u8")
        got = np.array([int.from_bytes(hashlib.sha256(s[b:b + blk].tobytes()).digest()[:8], "little") for b in range(0, n, blk)], np.uint64)
        bad = np.nonzero(got != want[first // blk:(first + n) // blk])[0]
        return {"match": bool(bad.size == 0), "blocks": int(got.size), "instances": n, "first_instance": int(first),
                "first_differing_block": (int(bad[0]) + first // blk) if bad.size else None, "fixture": "tests/golden/" + fixture,
                "what": "sha256 digests of Oracle B's checksums per 4096 instances: a match pins every witness checksum of the run"}
    except Exception as e:                                    # the verification must not take the measurement down
        return {"skipped": "error: %r" % (e,)}


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
        "data": "synthetic", "config": workload_config(),
        "run": {"note": "CPU arm: a bounded sample of the workload per step; the reference wasm (V8 unavailable) translated "
                        "to C (oracle/_ref), all host threads", "witnesses_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "witnesses/s", "cores": ncpu, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "witnesses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------
def run_own(args, rank, world, local_rank):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import hot_proofs_blake3_circom_b200 as pkg
    from hot_proofs_blake3_circom_b200 import _lib
    from hot_proofs_blake3_circom_b200 import inputs as gen
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

    def all_true(flag):
        if world == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t.item()))

    L = pkg.lib()
    ncpu = os.cpu_count() or 1
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    n = 1 << LOG2_BATCH
    first = rank * n                                            # shard = contiguous index range
    wc = pkg.builder("blake3_compression", device=local_rank, chunk=2048)
    rows = lcg_compression_inputs(n, first=first)
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_out = torch.empty(n * WIT_BYTES, dtype=torch.uint8, device="cuda")       # ordinary (cudaMalloc) memory
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_pub = torch.empty(n * 16, dtype=torch.int32, device="cuda")
    d_bad = torch.empty(n, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    alg_bytes = n * (WIT_BYTES + IN_BYTES)

    # The witnesses of the timed steps go where the library puts them by default: COMPRESSIBLE device memory
    # (b3w_device_alloc; the default HBM ring of the host-buffer calls): Blackwell's L2 compresses such lines on their way to
    # HBM and a witness is mostly zero bytes.  Same bytes on read-back (tests/test_gpu_compressible.py: every byte vs the
    # oracles).  The HBM roofline record is measured on ordinary memory right after, same kernel, same batch.
    out_c, compressible = None, False
    if not args.plain_output:
        try:
            p_c, granted = wc.device_alloc(n * WIT_BYTES, compressible=True)
            if granted:
                out_c, compressible = p_c, True
            else:
                wc.device_free(p_c)
        except pkg.B3WError:
            pass
    agreed = all_true(compressible)                               # every rank takes the same path (collectives follow)
    if compressible and not agreed:
        wc.device_free(out_c)
        out_c, compressible = None, False
    out_ptr = out_c if compressible else d_out.data_ptr()

    def step(ptr=None, checked=False):
        if checked:
            wc.witness_batch_device_checked(d_in.data_ptr(), n, ptr or out_ptr, d_st.data_ptr(), d_pub.data_ptr(), d_bad.data_ptr(), stream)
        else:
            wc.witness_batch_device(d_in.data_ptr(), n, ptr or out_ptr, d_st.data_ptr(), d_pub.data_ptr(), stream)

    def timed_launches(count, warm=3, **kw):
        """-> (total ms over `count` launches, mean ms per launch), CUDA events on the launching stream, max over ranks"""
        for _ in range(warm):
            step(**kw)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(count + 1)]
        barrier()
        torch.cuda.synchronize()
        ev[0].record()
        for k in range(count):
            step(**kw)
            ev[k + 1].record()
        torch.cuda.synchronize()
        barrier()
        per = [ev[k].elapsed_time(ev[k + 1]) for k in range(count)]
        return max_over_ranks(ev[0].elapsed_time(ev[-1])), max_over_ranks(sum(per) / len(per))

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # --- timed region: exactly K steps, one kernel launch each -------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    total_ms, kernel_ms = timed_launches(args.steps, warm=0)
    clocks = sampler.stop() if rank == 0 else None
    gpu_launches = world * args.steps                         # one witness kernel per step per rank
    assert int(d_st.max()) == 0
    pub0 = d_pub[:16].cpu().numpy().view(np.uint32)
    # the timed buffer is looked at: per-instance checksums of the bytes in it == those of the same kernel's output in
    # ordinary memory (identity of the two memory kinds; the oracle comparison of both is in tests/)
    d_s1 = torch.empty(n, dtype=torch.int64, device="cuda")
    d_s2 = torch.empty(n, dtype=torch.int64, device="cuda")
    wc.checksum_device(out_ptr, n, d_s1.data_ptr(), stream)
    step(d_out.data_ptr())
    wc.checksum_device(d_out.data_ptr(), n, d_s2.data_ptr(), stream)
    torch.cuda.synchronize()
    assert torch.equal(d_s1, d_s2), "witness bytes differ between compressible and ordinary memory"
    sums_xor = int(np.bitwise_xor.reduce(d_s1.cpu().numpy().view(np.uint64)))
    del d_s1, d_s2

    # --- sustained (>= 1 s of device time) and checked variants of the same launch ------------------
    reps_1s = int(1100.0 / kernel_ms) + 1
    sus_ms, _ = timed_launches(reps_1s, warm=1)
    chk_total, chk_ms = timed_launches(reps_1s, warm=2, checked=True)
    assert int(d_st.max()) == 0 and int(d_bad.min()) == -1                # B3W_NO_ROW everywhere
    # --- the HBM roofline record: the same kernel into ordinary (cudaMalloc) memory -------------------
    plain_total, plain_ms = (total_ms, kernel_ms)
    chk_plain_ms = None
    if compressible:
        plain_total, plain_ms = timed_launches(max(args.steps, 10), warm=3, ptr=d_out.data_ptr())
        _, chk_plain_ms = timed_launches(10, warm=2, ptr=d_out.data_ptr(), checked=True)

    # --- pure-store calibration (the write ceilings of this very GPU, same access shape) -------------
    def fill_rate(ptr, items):
        for _ in range(2):
            wc.calib_fill(ptr, n * WIT_BYTES, stream, items=items)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            wc.calib_fill(ptr, n * WIT_BYTES, stream, items=items)
        c1.record()
        torch.cuda.synchronize()
        nb = n * WIT_BYTES if not items else n * WIT_BYTES // 32768 * 32768
        return 5 * nb / c0.elapsed_time(c1) / 1e6
    fill_gbs = fill_rate(d_out.data_ptr(), False)
    fill_items_gbs = fill_rate(d_out.data_ptr(), True)
    fill_items_c_gbs = fill_rate(out_c, True) if compressible else None

    # --- the stand-alone R1CS check of witnesses where they lie (reads HBM): 2^15 of the witnesses just written ---------
    n_chk = 1 << 15
    step(d_out.data_ptr())
    if compressible:
        step(out_c)

    def check_rate(ptr, calc=None):
        calc = calc or wc
        for _ in range(2):
            calc.r1cs_check_device(ptr, n_chk, d_st.data_ptr(), d_bad.data_ptr(), stream)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            calc.r1cs_check_device(ptr, n_chk, d_st.data_ptr(), d_bad.data_ptr(), stream)
        c1.record()
        torch.cuda.synchronize()
        assert int(d_st[:n_chk].max()) == 0
        return max_over_ranks(c0.elapsed_time(c1) / 5)
    chk_hbm_ms = check_rate(d_out.data_ptr())
    chk_hbm_c_ms = check_rate(out_c) if compressible else None
    if compressible:
        wc.device_free(out_c)
    del d_out
    torch.cuda.empty_cache()
    # ... and of the nova step circuits' witnesses: the O2 build rust_fold loads (blake3_nova_pasta, Pallas Fr) and the O1 build
    chk_nova = {}
    from hot_proofs_blake3_circom_b200.inputs import splitmix_nova_inputs
    for nm in ("blake3_nova_pasta", "blake3_nova_o1"):
        wn = pkg.builder(nm, device=local_rank)
        dn_in = torch.from_numpy(splitmix_nova_inputs(n_chk, first=rank * n_chk).view(np.int32)).cuda()
        dn_out = torch.empty(n_chk * wn.witnessSize * 32, dtype=torch.uint8, device="cuda")
        wn.witness_batch_device(dn_in.data_ptr(), n_chk, dn_out.data_ptr(), d_st.data_ptr(), 0, stream)
        ms = check_rate(dn_out.data_ptr(), wn)
        nb = n_chk * wn.witnessSize * 32
        info = wn.r1cs_program_info()
        chk_nova[nm] = {"value": world * n_chk / (ms / 1e3), "unit": "witnesses/s", "kernel_ms": ms, "rows": info["rows"], "witness_bytes": wn.witnessSize * 32,
                        "read_gbs": nb / ms / 1e6, "frac_of_measured_hbm_peak": nb / ms / 1e6 / measured_peaks()[0]}
        del dn_out, dn_in
        wn.close()
        torch.cuda.empty_cache()

    # --- e2e: the C ABI host-buffer calls (what the N-API addon / a user calls) ---------------------
    avail = 0
    with open("/proc/meminfo") as f:
        for line in f:
            if line.startswith("MemAvailable"):
                avail = int(line.split()[1]) * 1024
    budget = int(avail * 0.4 / max(local_world, 1))
    n_e2e = n
    while n_e2e * WIT_BYTES > budget and n_e2e > 1024:
        n_e2e //= 2
    h_out = L.b3w_host_alloc_near(n_e2e * WIT_BYTES, local_rank)       # pinned, on the GPU's own NUMA node
    h_in = L.b3w_host_alloc_near(n_e2e * IN_BYTES, local_rank)
    h_st = L.b3w_host_alloc_near(n_e2e, local_rank)
    h_pub = L.b3w_host_alloc_near(n_e2e * 64, local_rank)
    if not (h_out and h_in and h_st and h_pub):
        raise SystemExit("bench.py: pinned host allocation failed: " + L.b3w_last_error().decode())
    C.memmove(h_in, rows.ctypes.data, n_e2e * IN_BYTES)
    e2e_steps = max(2, min(args.steps, 5))

    def wall(call, steps, warm=1):
        for _ in range(warm):
            _lib.check(call())
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            _lib.check(call())                                   # returns after the last D2H / host store
        dt = time.perf_counter() - t0
        barrier()
        return max_over_ranks(dt) / steps

    got0 = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_uint32)), shape=(WS * 8,))
    t_full = wall(lambda: L.b3w_witness_batch(wc._h, h_in, n_e2e, h_out, h_st, h_pub), e2e_steps)
    assert got0[8] == pub0[0] and got0[0] == 1                  # slot 0 == 1, slot 1 == out[0]
    t_compact = wall(lambda: L.b3w_witness_batch(wc._h, h_in, n_e2e, None, h_st, h_pub), e2e_steps)
    # hybrid: the same host buffer filled with every .wtns body, but only packed records cross PCIe; host threads expand
    got0[:] = 0
    hy_threads = max(1, ncpu // max(local_world, 1))
    t_hybrid = wall(lambda: L.b3w_witness_batch_hybrid(wc._h, h_in, n_e2e, h_out, h_st, h_pub, hy_threads), 2)
    assert got0[8] == pub0[0] and got0[0] == 1
    hy_t = wc.lastTiming()
    L.b3w_host_free(h_out)
    # ... and with the witnesses returned in COMPACT form (the per-instance trace, 3 776 B: every slot is a pure function
    # of it; b3w_unpack_device / b3w_unpack_host expand on demand)
    pk_words = wc.packedWords
    h_pk = L.b3w_host_alloc_near(n_e2e * pk_words * 4, local_rank)
    if not h_pk:
        raise SystemExit("bench.py: pinned host allocation failed: " + L.b3w_last_error().decode())
    t_packed = wall(lambda: L.b3w_witness_batch_packed(wc._h, h_in, n_e2e, h_pk, h_st, h_pub), e2e_steps)
    pk0 = np.ctypeslib.as_array(C.cast(h_pk, C.POINTER(C.c_uint32)), shape=(pk_words,))
    assert pk0[1] == 1 and pk0[30] == pub0[0]                   # trace word 1 = the constant 1, word 30 = out[0]
    for p in (h_pk, h_in, h_st, h_pub):
        L.b3w_host_free(p)
    wc.close()

    # --- BASELINE configs[3]: 2^20 blake3_nova_pasta steps, streamed through the HBM ring (whole job = 2^20: strong) ----
    def streamed(name, rows_fn, n_total, fused, reps_target_s, n_samples, tag, byte_check=False):
        n_r = n_total // world
        calc = pkg.builder(name, device=local_rank, chunk=16384, fused_check=fused, byte_check=byte_check)
        r = gen.parallel_rows(rows_fn, n_r, first=rank * n_r, threads=max(2, ncpu // max(local_world, 1)))
        hin = L.b3w_host_alloc_near(r.nbytes, local_rank)
        C.memmove(hin, r.ctypes.data, r.nbytes)
        del r
        hst, hpub = L.b3w_host_alloc_near(n_r, local_rank), L.b3w_host_alloc_near(n_r * calc.nPublic * 4, local_rank)
        hsum = L.b3w_host_alloc_near(n_r * 8, local_rank)
        rng = np.random.default_rng(1234 + rank)
        idx = np.unique(np.concatenate([[0, n_r - 1], rng.integers(0, n_r, n_samples - 2)])).astype(np.uint64) if n_samples else np.zeros(0, np.uint64)
        wb = calc.witnessSize * 32
        hsmp = L.b3w_host_alloc_near(max(idx.size, 1) * wb, local_rank)          # pinned: 1 024 copies of 771 KB out of the ring
        smp = np.ctypeslib.as_array(C.cast(hsmp, C.POINTER(C.c_uint8)), shape=(idx.size, wb)) if idx.size else np.zeros((0, wb), np.uint8)
        ex = _lib.BatchExtras()
        ex.sums = hsum
        if idx.size:
            ex.sample_idx, ex.n_samples, ex.sample_out = idx.ctypes.data, idx.size, hsmp
        call = lambda: L.b3w_witness_batch_ex(calc._h, hin, n_r, None, hst, hpub, C.byref(ex))
        t1 = wall(call, 1, warm=1)
        reps = max(1, int(reps_target_s / t1 + 0.999))
        dt = wall(call, reps, warm=0)
        st = np.ctypeslib.as_array(C.cast(hst, C.POINTER(C.c_uint8)), shape=(n_r,))
        sums = np.ctypeslib.as_array(C.cast(hsum, C.POINTER(C.c_uint64)), shape=(n_r,))
        assert not st.any(), "%s: status != 0" % tag
        # every sampled witness, copied out of the ring, has the checksum the kernel reported for it
        ok = bool(np.array_equal(gen.witness_checksums(smp, calc.witnessSize), sums[idx.astype(np.int64)])) if idx.size else None
        assert ok is not False, "%s: sample checksums differ" % tag
        tm = calc.lastTiming()
        fixture = {"blake3_compression": "compression_sums_2p24.npz", "blake3_nova_pasta": "nova_pasta_o2_sums_2p20.npz"}.get(name)
        vs_oracle = sums_vs_oracle_fixture(sums, rank * n_r, fixture) if fixture else None
        res = {"value": n_total / dt, "unit": "witnesses/s", "seconds_per_pass": dt, "passes_timed": reps, "instances": n_total,
               "instances_per_gpu": n_r, "witness_bytes": calc.witnessSize * 32, "generated_GB_per_pass": n_total * calc.witnessSize * 32 / 1e9,
               "ring_write_GBps_per_gpu": n_r * calc.witnessSize * 32 / dt / 1e9, "fused_check": fused, "byte_check": byte_check,
               "kernel_ms_sum_last_pass": tm["kernel_ms"], "launches_per_pass": tm["launches"], "d2h_bytes_per_pass_per_gpu": tm["d2h_bytes"],
               "kernel_ms_note": "b3w_last_timing: sum of the CUDA-event durations of the pass's launches; launches alternate between the two ring "
                                 "streams and overlap, so the sum exceeds the wall time of the pass",
               "sums_xor_rank0": int(np.bitwise_xor.reduce(sums)), "samples_per_gpu": int(idx.size), "samples_match_their_sums": ok}
        if vs_oracle is not None:                                # rank 0's record + the verdict of all ranks
            res["sums_vs_oracle_b"] = vs_oracle
            res["sums_match_oracle_b_on_all_ranks"] = all_true(vs_oracle.get("match") is not False) and "match" in vs_oracle
        del smp
        for p in (hin, hst, hpub, hsum, hsmp):
            L.b3w_host_free(p)
        calc.close()
        return res
    cfg4 = streamed("blake3_nova_pasta", gen.splitmix_nova_inputs, 1 << 20, False, 1.0, 0, "config4")
    cfg4["api"] = ("b3w_witness_batch_ex(out=NULL, sums): BASELINE configs[3], 2^20 blake3_nova_pasta (Pallas Fr) step witnesses streamed "
                   "through the HBM ring; status + z_{i+1} + checksums D2H")
    # --- BASELINE configs[4]: 2^24 blake3_compression instances over the N GPUs (2^24 / N each), fused R1CS check ------
    cfg5 = streamed("blake3_compression", gen.splitmix_compression_inputs, 1 << args.log2_config5, True, 1.0, 1024 // world, "config5")
    cfg5["api"] = ("b3w_witness_batch_ex(out=NULL, sums, %d samples per GPU) on a context with B3W_FLAG_FUSED_CHECK: BASELINE configs[4], "
                   "contiguous index ranges, no collective; per instance status + out[16] + a 64-bit witness checksum come back, plus the "
                   "full witnesses of the sample" % (1024 // world))

    # --- the same stream with B3W_FLAG_BYTE_CHECK: every chunk is read back from the ring and all 24 544 rows are evaluated
    # on its bytes before anything leaves the GPU (2^22 instances over the N GPUs) ------------------------------------------
    cfg5b = streamed("blake3_compression", gen.splitmix_compression_inputs, 1 << min(22, args.log2_config5), False, 0.5, 0, "config5_byte_check",
                     byte_check=True)
    cfg5b["api"] = ("b3w_witness_batch_ex(out=NULL, sums) on a context with B3W_FLAG_BYTE_CHECK: like config5, but instead of the fused "
                    "check on the trace the stand-alone checker (k_r1cs_check_fast) re-reads every witness from the HBM ring and evaluates "
                    "every row on the bytes that were stored")

    # --- BASELINE configs[2]: all Nova step witnesses of a synthetic 1 MiB file through b3w_nova_chain (every rank chains its
    # own file: weak), and a 64 MiB file for a sustained figure -------------------------------------------------------------
    def chained(mib, reps_target_s):
        import blake3
        from hot_proofs_blake3_circom_b200.witness_calculator import pinned_array
        calc = pkg.builder("blake3_nova", device=local_rank, chunk=16384)
        data = gen.splitmix_words(0xB3B30003 + rank, np.arange(mib << 18, dtype=np.uint64), 1)[:, 0].tobytes()
        nc, ns = C.c_uint64(), C.c_uint64()
        _lib.check(L.b3w_nova_chain_size(len(data), C.byref(nc), C.byref(ns)))
        ns = ns.value
        rows, status, pubz = pinned_array((ns, 32), np.uint32), pinned_array((ns,), np.uint8), pinned_array((ns, 15), np.uint32)
        h_data = pinned_array((len(data),), np.uint8)
        h_data[:] = np.frombuffer(data, np.uint8)
        step_off, root = np.zeros(nc.value + 1, np.uint64), np.zeros(32, np.uint8)
        call = lambda: L.b3w_nova_chain(calc._h, h_data.ctypes.data, len(data), None, status.ctypes.data, pubz.ctypes.data,
                                        rows.ctypes.data, step_off.ctypes.data, root.ctypes.data)
        t1 = wall(call, 1, warm=2)
        reps = max(3, int(reps_target_s / t1 + 0.999))
        dt = wall(call, reps, warm=0)
        digest = blake3.blake3(data).digest()
        last = step_off[1:].astype(np.int64) - 1
        folds = bool((pubz[last, 2:10].view(np.uint8).reshape(len(last), 32) == np.frombuffer(digest, np.uint8)).all())
        assert root.tobytes() == digest and folds and not status.any(), "config3: chain does not fold to BLAKE3(file)"
        res = {"value": world * ns / dt, "unit": "step witnesses/s", "file_MiB_per_gpu": mib, "chunks_per_gpu": nc.value, "step_witnesses_per_gpu": ns,
               "seconds_per_file": dt, "files_timed": reps, "witness_bytes": calc.witnessSize * 32,
               "ring_write_GBps_per_gpu": ns * calc.witnessSize * 32 / dt / 1e9, "root_is_blake3_of_file": True, "every_chunk_folds_to_root": folds}
        calc.close()
        return res
    cfg3 = chained(1, 0.3)
    cfg3["api"] = ("b3w_nova_chain(out=NULL): BASELINE configs[2], host bytes in -> device BLAKE3 tree, step rows, all 26 624 step witnesses "
                   "(blake3_nova, BN254 O2 layout) through the HBM ring, z_{i+1} / status / rows back (pinned)")
    cfg3["sustained_64MiB"] = chained(64, 0.5)

    # --- field-element rows (b3w_witness_batch_fr) next to u32 rows: 2^20 blake3_nova_pasta, N = 1 only ------------------
    fr_line = None
    if world == 1 and not args.no_fr:
        n_f = 1 << 20
        calc = pkg.builder("blake3_nova_pasta", device=local_rank, chunk=16384)
        r = gen.parallel_rows(gen.splitmix_nova_inputs, n_f, threads=ncpu)
        hin = L.b3w_host_alloc(r.nbytes)
        C.memmove(hin, r.ctypes.data, r.nbytes)
        hfr = L.b3w_host_alloc(n_f * 1024)
        fr = np.ctypeslib.as_array(C.cast(hfr, C.POINTER(C.c_uint8)), shape=(n_f, 32, 32))
        fr[:] = 0
        fr[:, :, 0:4] = r.view(np.uint8).reshape(n_f, 32, 4)
        hst, hpub = L.b3w_host_alloc(n_f), L.b3w_host_alloc(n_f * 60)
        t_u32 = wall(lambda: L.b3w_witness_batch(calc._h, hin, n_f, None, hst, hpub), 3)
        t_fr = wall(lambda: L.b3w_witness_batch_fr(calc._h, hfr, n_f, None, hst, hpub), 3)
        p = calc.prime
        for i in range(50, n_f, 100):                                # 1 %: n_blocks becomes a genuine field element
            fr[i, 0] = np.frombuffer(((p - 1 - i) % p).to_bytes(32, "little"), np.uint8)
        t_mix = wall(lambda: L.b3w_witness_batch_fr(calc._h, hfr, n_f, None, hst, hpub), 3)
        st = np.ctypeslib.as_array(C.cast(hst, C.POINTER(C.c_uint8)), shape=(n_f,))
        assert not st.any()
        fr_line = {"unit": "witnesses/s", "instances": n_f, "circuit": "blake3_nova_pasta", "u32_rows_b3w_witness_batch": n_f / t_u32,
                   "fr_rows_all_u32_b3w_witness_batch_fr": n_f / t_fr, "fr_rows_1pct_field_valued": n_f / t_mix,
                   "h2d_bytes_u32": n_f * 128, "h2d_bytes_fr": n_f * 1024,
                   "note": "host pinned rows in, out=NULL; Fr256 rows are converted on the device, the 1 % field-valued instances run on the "
                           "general kernel into the same ring slots"}
        for q in (hin, hfr, hst, hpub):
            L.b3w_host_free(q)
        calc.close()

    # --- cpu baseline (rank 0, N=1 only): bounded sample on all host cores -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_reference_rate(ncpu, ncpu)                          # warm-up, discarded
        n_s = 6 * ncpu
        rate, secs, kind = cpu_reference_rate(n_s, ncpu, first=ncpu)
        cpu = {"value": rate, "unit": "witnesses/s", "cores": ncpu, "kind": kind,
               "sample": "%d witnesses (6 per host thread, %d threads) of the same LCG(6429) workload incl. input set-up "
                         "and read-out, %.1f s wall; reference wasm translated to C (no V8 here)" % (n_s, ncpu, secs)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    traffic, traffic_c = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        traffic = tj.get("k_blake3_comp_witness_dram_bytes_per_launch")
        traffic_c = tj.get("k_blake3_comp_witness_dram_bytes_per_launch_compressible")
    mem_kind = ("compressible device memory (b3w_device_alloc, CU_MEM_ALLOCATION_COMP_GENERIC; the library's default ring): the L2 compresses "
                "witness lines on their way to HBM; identical bytes on read-back") if compressible else "ordinary device memory (cudaMalloc)"
    value = world * n * args.steps / (total_ms / 1e3)
    achieved_plain = alg_bytes / plain_ms / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": "witnesses/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 + Fr256 (BN254)", "data": "synthetic",
        "config": workload_config(n),
        "run": {"output_memory": mem_kind, "out0_instance0": int(pub0[0]), "witness_checksums_xor_rank0": sums_xor},
        # the HBM record: the witness kernel into ORDINARY memory, where every algorithmic byte crosses the HBM interface
        "roofline": {"bound": "hbm", "kernel": "k_blake3_comp_witness", "achieved": achieved_plain, "peak": peak,
                     "unit": "GB/s", "frac": achieved_plain / peak, "traffic": traffic,
                     "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture "
                                       "of this kernel on ordinary memory (committed profile); NOT measured in this run",
                     "peak_source": peak_src + (" (of measured)" if "MEASURED" in peak_src else " (of fallback)"),
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": plain_ms, "launches_timed": max(args.steps, 10),
                     "measured_on": "ordinary (cudaMalloc) output, same kernel and batch as the timed region, CUDA events per launch in this run",
                     "frac_of_spec_8TBps": achieved_plain / 8000.0, "pure_store_fill_gbs_same_gpu": fill_gbs,
                     "pure_store_same_stream_shape_gbs": fill_items_gbs, "frac_of_pure_store_same_shape": achieved_plain / fill_items_gbs},
        "e2e": {"value": world * n_e2e / t_full, "unit": "witnesses/s", "h2d_bytes_per_step": n_e2e * IN_BYTES,
                "d2h_bytes_per_step": n_e2e * (WIT_BYTES + 1 + 64), "instances_per_step_per_gpu": n_e2e,
                "ms_per_step": 1e3 * t_full, "d2h_gbs_per_gpu": n_e2e * WIT_BYTES / t_full / 1e9,
                "api": "b3w_witness_batch(host pinned in/out): every witness byte copied to the host"},
        "e2e_compact": {"value": world * n_e2e / t_compact, "unit": "witnesses/s", "h2d_bytes_per_step": n_e2e * IN_BYTES,
                        "d2h_bytes_per_step": n_e2e * (1 + 64), "ms_per_step": 1e3 * t_compact,
                        "api": "b3w_witness_batch(out=NULL): witnesses stream through the HBM ring, status + out[16] return"},
        "e2e_packed": {"value": world * n_e2e / t_packed, "unit": "witnesses/s", "h2d_bytes_per_step": n_e2e * IN_BYTES,
                       "d2h_bytes_per_step": n_e2e * (pk_words * 4 + 1 + 64), "ms_per_step": 1e3 * t_packed,
                       "api": "b3w_witness_batch_packed(host pinned in/out): every witness returned in compact form "
                              "(%d B trace per instance; expandable to the .wtns body with b3w_unpack_device / b3w_unpack_host)" % (pk_words * 4)},
        "e2e_hybrid": {"value": world * n_e2e / t_hybrid, "unit": "witnesses/s", "h2d_bytes_per_step": n_e2e * IN_BYTES,
                       "d2h_bytes_per_step": n_e2e * (pk_words * 4 + 1 + 64), "ms_per_step": 1e3 * t_hybrid, "host_threads_per_gpu": hy_threads,
                       "host_unpack_ms_per_step": hy_t["host_ms"], "host_store_gbs_per_gpu": n_e2e * WIT_BYTES / t_hybrid / 1e9,
                       "api": "b3w_witness_batch_hybrid: every .wtns body in the caller's host buffer like e2e, but only the packed records cross "
                              "PCIe; the expansion runs on host threads (non-temporal stores)"},
        "value_sustained": {"value": world * n * reps_1s / (sus_ms / 1e3), "unit": "witnesses/s", "launches": reps_1s, "device_seconds": sus_ms / 1e3,
                            "note": "the timed region's launch repeated back to back for >= 1 s"},
        "value_checked": {"value": world * n / (chk_ms / 1e3), "unit": "witnesses/s", "kernel": "k_blake3_comp_witness<CHECK> (fused R1CS check)",
                          "kernel_ms": chk_ms, "launches": reps_1s, "device_seconds": chk_total / 1e3, "output_memory": "as the timed region",
                          "value_plain_memory": world * n / (chk_plain_ms / 1e3) if chk_plain_ms else None},
        "r1cs_check_resident": {"kernel": "k_r1cs_check_fast (b3w_r1cs_check_device): all 24 544 rows on witnesses read back from HBM",
                                "instances": n_chk, "value": world * n_chk / (chk_hbm_ms / 1e3), "unit": "witnesses/s", "kernel_ms": chk_hbm_ms,
                                "read_gbs": n_chk * WIT_BYTES / chk_hbm_ms / 1e6, "frac_of_measured_hbm_peak": n_chk * WIT_BYTES / chk_hbm_ms / 1e6 / peak,
                                "compressible_buffer": None if chk_hbm_c_ms is None else {
                                    "value": world * n_chk / (chk_hbm_c_ms / 1e3), "kernel_ms": chk_hbm_c_ms, "read_gbs": n_chk * WIT_BYTES / chk_hbm_c_ms / 1e6},
                                "nova": chk_nova},
        "config3": cfg3, "config4": cfg4, "config5": cfg5, "config5_byte_check": cfg5b,
        "gpu_launches": gpu_launches, "clocks": clocks}
    if compressible:
        achieved_c = alg_bytes / kernel_ms / 1e6
        line["roofline_compressible"] = {
            "bound": "sm-store-path", "kernel": "k_blake3_comp_witness", "achieved": achieved_c, "peak": fill_items_c_gbs, "unit": "GB/s",
            "frac": achieved_c / fill_items_c_gbs, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes,
            "peak_source": "k_fill_items into the same compressible buffer, measured in this run: the witness kernels' store stream "
                           "(same items, same grid, 1 KiB warp stores) with none of their work",
            "traffic": traffic_c, "traffic_source": "profiles/traffic.json (ncu capture on compressible memory, committed); NOT measured in this run",
            "note": "the timed region: witness bytes per second INTO COMPRESSIBLE memory.  The HBM interface carries about a third of them "
                    "(traffic), so this is not an HBM fraction: the SM-side store path is what bounds it"}
        line["value_plain_memory"] = world * n / (plain_ms / 1e3)
    if fr_line:
        line["fr_batches"] = fr_line
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
    ap.add_argument("--no-fr", action="store_true", help="skip the field-element-row comparison (N = 1 only)")
    ap.add_argument("--log2-config5", type=int, default=24, help="log2 of config 5's whole-job instance count")
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
