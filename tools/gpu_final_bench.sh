#!/bin/bash
# the headline line at the committed state
TAG=${1:-r1u}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 600 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err
cut -c1-3200 gpurun_out/bench_${TAG}_default.json; tail -2 gpurun_out/bench_${TAG}_default.err
