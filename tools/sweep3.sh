#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
(nvidia-smi topo -m 2>&1 | head -14; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa $(cat $d/numa_node)"; fi; done; lscpu | grep -i numa) > gpurun_out/topo_n$N.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_numa_n$N.json 2> gpurun_out/bench_numa_n$N.err
tail -2 gpurun_out/bench_numa_n$N.err; cat gpurun_out/bench_numa_n$N.json
