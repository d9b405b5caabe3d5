#!/bin/bash
# round 2, last GPU session: bench (both arms), ncu launch list of the bench command, full captures of the witness kernel
# (ordinary + compressible memory) and of the reworked stand-alone checker on the three systems, checker sweep, configs 0/2,
# then the whole GPU suite.  (compute-sanitizer: tools/sessions/gpu_r2z_sanitize.sh, a call of its own.)
mkdir -p gpurun_out
(time python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench_own.json 2> gpurun_out/r2z_bench_own.err) 2>&1 | tail -3 | tee gpurun_out/r2z_bench_wall.txt
tail -5 gpurun_out/r2z_bench_own.err; cut -c1-600 gpurun_out/r2z_bench_own.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_ref.json; cat gpurun_out/r2z_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fr --log2-config5 20 > gpurun_out/r2z_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blake3_comp -s 2 -c 1 -f -o gpurun_out/r2z_prof_comp python tools/prof_run.py 16 4 > gpurun_out/r2z_ncu_full.log 2>&1; tail -1 gpurun_out/r2z_ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blake3_comp -s 2 -c 1 -f -o gpurun_out/r2z_prof_comp_c python tools/prof_run.py 16 4 blake3_compression plain compressible > gpurun_out/r2z_ncu_full_c.log 2>&1; tail -1 gpurun_out/r2z_ncu_full_c.log
for c in blake3_compression blake3_nova_pasta blake3_nova_o1; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_check_fast -s 1 -c 1 -f -o gpurun_out/r2z_prof_r1cs_$c python tools/prof_run.py 15 3 $c r1cs > gpurun_out/r2z_ncu_r1cs_$c.log 2>&1; tail -1 gpurun_out/r2z_ncu_r1cs_$c.log
done
python tools/r1cs_sweep.py 2>&1 | tee gpurun_out/r2z_final_r1cs_sweep.jsonl
timeout 600 python tools/bench_configs.py 1 3 2>&1 | tee gpurun_out/r2z_configs_0_2.jsonl
timeout 1500 python -m pytest tests -q -m gpu --durations=15 2>&1 | tail -24 | tee gpurun_out/r2z_pytest.log
ls -la gpurun_out | tail -12
