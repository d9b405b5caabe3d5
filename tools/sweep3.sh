#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_chain.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
python tools/bench_configs.py 5 2>&1 | tee gpurun_out/config5_n$N.log
