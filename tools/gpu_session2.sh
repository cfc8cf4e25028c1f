#!/bin/bash
# round-2 session 2 (1 GPU): tests that changed, then both bench arms like the driver runs them, then the other configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_hnsw_build.py tests/test_gpu_kmeans.py tests/test_gpu_hnsw.py -m gpu -q --timeout=600 > gpurun_out/r2_tests2.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests2.log; tail -8 gpurun_out/r2_tests2.log
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r2_bench_B_ref.json 2> gpurun_out/r2_bench_B_ref.err; tail -3 gpurun_out/r2_bench_B_ref.err
( time timeout 1200 python bench.py ) > gpurun_out/r2_bench_B.json 2> gpurun_out/r2_bench_B.err; tail -5 gpurun_out/r2_bench_B.err; cut -c1-600 gpurun_out/r2_bench_B.json
( time timeout 600 python bench.py --config A ) > gpurun_out/r2_bench_A.json 2> gpurun_out/r2_bench_A.err; tail -3 gpurun_out/r2_bench_A.err
( time timeout 900 python bench.py --config C ) > gpurun_out/r2_bench_C.json 2> gpurun_out/r2_bench_C.err; tail -3 gpurun_out/r2_bench_C.err
( time timeout 1200 python bench.py --config E ) > gpurun_out/r2_bench_E.json 2> gpurun_out/r2_bench_E.err; tail -3 gpurun_out/r2_bench_E.err
( time timeout 1200 python bench.py --config D ) > gpurun_out/r2_bench_D.json 2> gpurun_out/r2_bench_D.err; tail -3 gpurun_out/r2_bench_D.err
for f in A C E D; do echo "== $f"; cut -c1-400 gpurun_out/r2_bench_$f.json; done
