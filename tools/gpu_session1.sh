#!/bin/bash
# round-2 session 1: full GPU test suite + HNSW build checks
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --ignore=tests/test_gpu_hnsw_build.py > gpurun_out/r2_tests1.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests1.log
tail -15 gpurun_out/r2_tests1.log
timeout 900 python -m pytest tests/test_gpu_hnsw_build.py -m gpu -q --timeout=300 > gpurun_out/r2_tests1_build.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests1_build.log
tail -15 gpurun_out/r2_tests1_build.log
timeout 300 python tools/hnsw_build_check.py --config C --rows 200000 > gpurun_out/r2_hnswbuild_C200k.json 2> gpurun_out/r2_hnswbuild_C200k.err
cat gpurun_out/r2_hnswbuild_C200k.json; tail -3 gpurun_out/r2_hnswbuild_C200k.err
timeout 600 python tools/hnsw_build_check.py --config C --rows 1000000 > gpurun_out/r2_hnswbuild_C1M.json 2> gpurun_out/r2_hnswbuild_C1M.err
cat gpurun_out/r2_hnswbuild_C1M.json; tail -3 gpurun_out/r2_hnswbuild_C1M.err
timeout 600 python tools/hnsw_build_check.py --config E --rows 2000000 > gpurun_out/r2_hnswbuild_E2M.json 2> gpurun_out/r2_hnswbuild_E2M.err
cat gpurun_out/r2_hnswbuild_E2M.json; tail -3 gpurun_out/r2_hnswbuild_E2M.err
