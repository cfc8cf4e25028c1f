#!/bin/bash
# two-GPU sanity of the list-sharded scan with the batched kernels (NCCL): correctness check + one bench line
TAG=${1:-r1l}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_gpu_check_$TAG.json 2> gpurun_out/multi_gpu_check_$TAG.err
tail -c 600 gpurun_out/multi_gpu_check_$TAG.json; tail -2 gpurun_out/multi_gpu_check_$TAG.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_${TAG}_2gpu.json 2> gpurun_out/bench_${TAG}_2gpu.err
cut -c1-2500 gpurun_out/bench_${TAG}_2gpu.json; tail -3 gpurun_out/bench_${TAG}_2gpu.err
