#!/bin/bash
# Third-pass GPU session: full GPU test suite, then the secondary workloads after the
# paired-row / packed-half bulk scan, the tiled exact scan and the sync-free k-means++.
TAG=${1:-r1g}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
run() {  # name, args...
  local name=$1; shift
  timeout 400 python tools/bench_extra.py "$@" > gpurun_out/extra_${name}_$TAG.json 2> gpurun_out/extra_${name}_$TAG.err
  cut -c1-1000 gpurun_out/extra_${name}_$TAG.json; tail -2 gpurun_out/extra_${name}_$TAG.err
}
run ivf_halfvec ivf --elem halfvec
run ivf_bit ivf --elem bit
run exact1m exact --rows 1000000 --dim 128 --queries 512
run exact1q exact --rows 1000000 --dim 128 --queries 32
run kmeans kmeans --k 4096 --iters 10
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:"members|sum_centers|finish_centers|write_centers|keep_previous|count_and_diff|assign|pack_planes|row_sqnorm|DeviceScan" \
    --csv --log-file gpurun_out/launches_kmeans_$TAG.csv python tools/bench_extra.py kmeans --k 4096 --iters 3 > gpurun_out/ncu_kmeans_$TAG.log 2>&1
tail -2 gpurun_out/ncu_kmeans_$TAG.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err
cut -c1-1500 gpurun_out/bench_${TAG}_default.json; tail -2 gpurun_out/bench_${TAG}_default.err
