#!/bin/bash
TAG=${1:-r1w}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_${TAG}_2gpu.json 2> gpurun_out/bench_${TAG}_2gpu.err
echo "stdout lines: $(wc -l < gpurun_out/bench_${TAG}_2gpu.json)"; head -c 200 gpurun_out/bench_${TAG}_2gpu.json; echo; grep -c "NCCL version" gpurun_out/bench_${TAG}_2gpu.err; tail -2 gpurun_out/bench_${TAG}_2gpu.err | cut -c1-200
