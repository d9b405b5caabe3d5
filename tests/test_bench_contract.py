"""The benchmark contract, checked WITHOUT a GPU: the JSON lines bench.py printed on the B200 boxes (committed under
profiles/) carry every key the driver reads, with coherent values; the reference arm's line has its own shape; and
bench.py refuses to run its own arm without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


@pytest.mark.parametrize("name,n", [("r02z_bench_own.json", 1), ("r02z_bench_n2.json", 2), ("r02z_bench_n8.json", 8)])
def test_own_arm_line(name, n):
    d = line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "blake3_compression witnesses/sec" and d["unit"] == "witnesses/s" and d["n_gpus"] == n
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"] and d["warmup"] >= 3 and d["gpu_launches"] == d["steps"] * n   # one launch per step and GPU
    per_gpu = d["config"]["instances_per_gpu"]
    assert abs(d["value"] - n * per_gpu / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6       # value = whole-job units / max-over-ranks time
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms"] / 1e3) / 1e9) / r["achieved"] < 1e-6
    assert 0.9 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.1                        # DRAM traffic ~ algorithmic bytes: nothing re-read
    e = d["e2e"]
    assert e["unit"] == "witnesses/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] >= per_gpu * 770976
    assert e["value"] < d["value"] / 50                                                        # PCIe-bound, not a copy of `value`
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        b = d["cpu_baseline"]
        assert b["kind"] == "reference" and b["cores"] >= 1 and b["unit"] == "witnesses/s" and b["sample"]
        for k in ("value_checked", "r1cs_check_resident", "config3", "config4", "config5", "config5_byte_check", "fr_batches", "e2e_hybrid"):
            assert k in d, k
        assert d["config5"]["samples_match_their_sums"] is True and d["config3"]["root_is_blake3_of_file"] is True


def test_reference_arm_line():
    d = line("r02z_bench_ref.json")
    assert d["impl"] == "reference" and d["metric"] == "blake3_compression witnesses/sec" and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    own = line("r02z_bench_own.json")
    assert d["config"]["workload"] == own["config"]["workload"] and d["unit"] == own["unit"]


def test_own_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box WITHOUT a GPU")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and not p.stdout.strip().startswith("{")                          # no number without the CUDA path


def test_both_arms_print_the_same_config():
    """The driver compares the arms' `config` objects: both come from bench.workload_config(), which holds the workload
    and nothing a run measured (that is under `run`).  The reference arm is run here (1 step, a few seconds of CPU)."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import ref_wasm
    cfg = bench.workload_config()
    assert cfg["workload"] == bench.WORKLOAD and cfg["instances_per_gpu"] == 1 << 16 and cfg["witness_bytes"] == 24093 * 32
    assert "l2" in cfg and "model" not in cfg
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": workload_config(') == 2                                        # own arm and reference arm
    if not ref_wasm.available("compression"):
        pytest.skip("oracle/_ref not built")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-400:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["config"] == cfg and d["higher_is_better"] is True and d["unit"] == "witnesses/s"
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and d["e2e"]["value"] == d["value"] > 0


def test_bench_checks_streamed_checksums_against_the_oracle_fixture(built):
    """bench.py holds every checksum of config 5's streamed run to committed digests of Oracle B's (tests/golden/); the
    helper is exercised here with Oracle B's own sums standing in for the GPU's."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    from oracle import port
    from hot_proofs_blake3_circom_b200.inputs import splitmix_compression_inputs
    first = 101 * 4096
    sums = port.witness_batch("compression", splitmix_compression_inputs(8192, first=first), want="sums")
    for fixture in ("compression_sums_2p20.npz", "compression_sums_2p24.npz"):
        if not os.path.exists(os.path.join(ROOT, "tests", "golden", fixture)):
            assert "skipped" in bench.sums_vs_oracle_fixture(sums, first, fixture)
            continue
        r = bench.sums_vs_oracle_fixture(sums, first, fixture)
        assert r["match"] is True and r["blocks"] == 2 and r["first_differing_block"] is None
        bad = sums.copy()
        bad[4096 + 7] ^= 1
        r = bench.sums_vs_oracle_fixture(bad, first, fixture)
        assert r["match"] is False and r["first_differing_block"] == 102
        assert "skipped" in bench.sums_vs_oracle_fixture(sums, first + 1, fixture)              # not whole blocks
        assert "skipped" in bench.sums_vs_oracle_fixture(sums[:100], first, fixture)
        assert "skipped" in bench.sums_vs_oracle_fixture(sums, (1 << 24), fixture)               # beyond the fixture
        json.dumps(r)
    assert "skipped" in bench.sums_vs_oracle_fixture(sums, first, "no_such_fixture.npz")


def test_committed_bench_line_agrees_with_oracle_b_on_2p24_checksums():
    """The XOR of all per-instance witness checksums the B200 reported for config 5 (2^24 blake3_compression) and config 4
    (2^20 blake3_nova_pasta) in the committed single-GPU bench line equals the XOR of Oracle B's checksums of the same
    instances (tests/golden/*_sums_2p*.npz, generated AFTER that GPU run): 64 bits over 12.9 TB / 0.78 TB of witnesses."""
    import numpy as np
    d = line("r02z_bench_own.json")
    g5 = np.load(os.path.join(ROOT, "tests", "golden", "compression_sums_2p24.npz"))
    g4 = np.load(os.path.join(ROOT, "tests", "golden", "nova_pasta_o2_sums_2p20.npz"))
    assert d["config5"]["instances"] == 1 << 24 and d["config5"]["sums_xor_rank0"] == int(g5["xor"])
    assert d["config4"]["instances"] == 1 << 20 and d["config4"]["sums_xor_rank0"] == int(g4["xor"])


@pytest.mark.parametrize("name", ["r02z_bench_own.json", "r02z_bench_n2.json", "r02z_bench_n8.json"])
def test_committed_bench_lines_rank0_shards_agree_with_oracle_b(name):
    """Rank 0's shard of a streamed run is a prefix of the instance sequence: the XOR of the checksums the B200s reported for
    it (config 5: 2^24 / N instances with the fused check; the byte-check run: 2^22 / N) equals the XOR of Oracle B's
    checksums of that prefix (tests/golden/compression_sums_prefix_xor.npz, computed after those runs)."""
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "compression_sums_prefix_xor.npz"))
    want = {1 << int(k): int(x) for k, x in zip(g["log2_n"], g["xor"])}
    g24 = np.load(os.path.join(ROOT, "tests", "golden", "compression_sums_2p24.npz"))
    want[1 << 24] = int(g24["xor"])
    d = line(name)
    for key in ("config5", "config5_byte_check"):
        per_gpu = d[key]["instances_per_gpu"]
        assert per_gpu * d["n_gpus"] == d[key]["instances"] and per_gpu in want, (key, per_gpu)
        assert d[key]["sums_xor_rank0"] == want[per_gpu], (name, key, per_gpu)
    # config 4 (2^20 blake3_nova_pasta steps, Pallas scalar field): 2^20 / N per rank
    g = np.load(os.path.join(ROOT, "tests", "golden", "nova_pasta_o2_sums_prefix_xor.npz"))
    want = {1 << int(k): int(x) for k, x in zip(g["log2_n"], g["xor"])}
    want[1 << 20] = int(np.load(os.path.join(ROOT, "tests", "golden", "nova_pasta_o2_sums_2p20.npz"))["xor"])
    per_gpu = d["config4"]["instances_per_gpu"]
    assert per_gpu * d["n_gpus"] == 1 << 20 and d["config4"]["sums_xor_rank0"] == want[per_gpu], (name, per_gpu)


def test_committed_bench_lines_headline_batch_agrees_with_oracle_b(built):
    """The timed region's own batch (2^16 LCG(6429) instances, rank 0 = instances 0 .. 2^16): the XOR of the checksums the
    B200 computed from the bytes in the timed COMPRESSIBLE buffer equals Oracle B's, re-derived live (14 s on 8 cores)."""
    import numpy as np
    from oracle import port
    from hot_proofs_blake3_circom_b200.inputs import lcg_compression_inputs
    sums = port.witness_batch("compression", lcg_compression_inputs(1 << 16), want="sums")
    want = int(np.bitwise_xor.reduce(sums))
    for name in ("r02z_bench_own.json", "r02z_bench_n2.json", "r02z_bench_n8.json"):
        d = line(name)
        got = d.get("run", d["config"])["witness_checksums_xor_rank0"]          # under `config` in lines printed before `run` existed
        assert got == want, name
