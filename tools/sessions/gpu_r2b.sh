#!/bin/bash
# round 2, second GPU session: checker timings after the Fr / 64-bit rework, store-mode experiment, the bench (both arms),
# ncu launch list of the bench command, full captures of the witness kernel (ordinary + compressible) and of the checker
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_r1cs.py tests/test_gpu_compressible.py::test_tma_store_mode_gives_the_same_bytes -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2b_pytest.log
timeout 300 python tools/r1cs_quickbench.py 2>&1 | tee gpurun_out/r2b_r1cs.jsonl
timeout 300 python tools/store_mode_bench.py 2>&1 | tee gpurun_out/r2b_store_mode.jsonl
(time python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_own.json 2> gpurun_out/r2b_bench_own.err) 2>&1 | tail -3 | tee gpurun_out/r2b_bench_wall.txt
tail -5 gpurun_out/r2b_bench_own.err; cat gpurun_out/r2b_bench_own.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref.json; cat gpurun_out/r2b_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fr --log2-config5 20 > gpurun_out/r2b_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blake3_comp -s 2 -c 1 -f -o gpurun_out/r2b_prof_comp python tools/prof_run.py 16 4 > gpurun_out/r2b_ncu_full.log 2>&1; tail -2 gpurun_out/r2b_ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blake3_comp -s 2 -c 1 -f -o gpurun_out/r2b_prof_comp_c python tools/prof_run.py 16 4 blake3_compression plain compressible > gpurun_out/r2b_ncu_full_c.log 2>&1; tail -2 gpurun_out/r2b_ncu_full_c.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_check_fast -s 1 -c 1 -f -o gpurun_out/r2b_prof_r1cs python tools/prof_run.py 15 3 blake3_compression r1cs > gpurun_out/r2b_ncu_r1cs.log 2>&1; tail -2 gpurun_out/r2b_ncu_r1cs.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_check_fast -s 1 -c 1 -f -o gpurun_out/r2b_prof_r1cs_nova python tools/prof_run.py 15 3 blake3_nova_pasta r1cs > gpurun_out/r2b_ncu_r1cs_nova.log 2>&1; tail -2 gpurun_out/r2b_ncu_r1cs_nova.log
ls -la gpurun_out | tail -20
