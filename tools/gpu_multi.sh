#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, one rank per GPU), ours then the reference arm.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_multi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; tail -2 gpurun_out/bench_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; tail -1 gpurun_out/bench_ref_n$N.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_same_box.log 2>&1; tail -1 gpurun_out/bench_n1_same_box.log
