#!/bin/bash
# round-2 session 4 (1 GPU): selection-kernel A/B, launch lists, build-phase diagnostics, ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_headline.py tests/test_gpu_ivfflat.py tests/test_gpu_hnsw.py tests/test_gpu_hnsw_build.py tests/test_gpu_kmeans.py tests/test_gpu_ops.py -m gpu -q --timeout=600 > gpurun_out/r2_tests4.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests4.log; tail -8 gpurun_out/r2_tests4.log
for f in 0 1 2; do
  VB_FUSED_REFINE=$f timeout 600 python bench.py --no-extras --no-cpu --law rank16 > gpurun_out/r2_ab_fused$f.json 2> gpurun_out/r2_ab_fused$f.err
done
python - <<'PY'
import json
for n in ("0","1","2"):
    try:
        d=json.load(open(f"gpurun_out/r2_ab_fused{n}.json"))
        print("fused_refine", n, "qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"], json.dumps(d["roofline"]["other_kernels_ms_per_step"]), "list_tc", round(d["roofline"]["avg_launch_ms"],4), "frac", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_B.csv python bench.py --no-extras --no-cpu --no-recall --law rank16 --steps 2 --warmup 1 > gpurun_out/r2_ncu_B.log 2>&1
timeout 600 python bench.py --config C --no-cpu > gpurun_out/r2_bench_C3.json 2> gpurun_out/r2_bench_C3.err; cut -c1-200 gpurun_out/r2_bench_C3.json
timeout 900 python bench.py --config E --no-cpu > gpurun_out/r2_bench_E3.json 2> gpurun_out/r2_bench_E3.err; cut -c1-200 gpurun_out/r2_bench_E3.json
timeout 600 python tools/diag_build_phases.py > gpurun_out/r2_diag_build.json 2> gpurun_out/r2_diag_build.err; cat gpurun_out/r2_diag_build.json; tail -2 gpurun_out/r2_diag_build.err
ROUNDS=24 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_build.csv python tools/diag_build_phases.py > gpurun_out/r2_ncu_build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:list_tc_kernel -s 9 -c 1 -o gpurun_out/r2_listtc -f python bench.py --no-extras --no-cpu --no-recall --law rank16 --queries 4096 --steps 2 --warmup 1 > gpurun_out/r2_ncu_listtc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 1 -o gpurun_out/r2_hnsw_C -f python bench.py --config C --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_ncu_hnsw.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null
