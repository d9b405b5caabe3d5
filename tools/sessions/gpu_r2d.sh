#!/bin/bash
# round 2, 2-GPU session: the multi-GPU tests on two real devices, then the bench under torchrun as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2d_gpus.txt
timeout 600 python -m pytest tests/test_gpu_extras.py tests/test_gpu_parity.py tests/test_gpu_chain.py -q -m gpu -k "multi or restore or serialises or counter or imperfect" 2>&1 | tail -5 | tee gpurun_out/r2d_pytest.log
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err) 2>&1 | tail -3
tail -5 gpurun_out/r2d_bench_n2.err; cat gpurun_out/r2d_bench_n2.json | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2d_bench_ref_n2.json 2>/dev/null; cat gpurun_out/r2d_bench_ref_n2.json | cut -c1-300
