#!/bin/bash
# round-2 session 3 (1 GPU): fused refine A/B, launch list of a step, HNSW build quality, config D again
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_headline.py tests/test_gpu_ivfflat.py tests/test_gpu_hnsw.py tests/test_gpu_hnsw_build.py tests/test_ext_harness.py -m gpu -q --timeout=600 > gpurun_out/r2_tests3.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests3.log; tail -6 gpurun_out/r2_tests3.log
VB_FUSED_REFINE=0 timeout 600 python bench.py --no-extras --no-cpu --law rank16 > gpurun_out/r2_ab_unfused.json 2> gpurun_out/r2_ab_unfused.err
timeout 600 python bench.py --no-extras --no-cpu --law rank16 > gpurun_out/r2_ab_fused.json 2> gpurun_out/r2_ab_fused.err
python - <<'PY'
import json
for n in ("unfused","fused"):
    try:
        d=json.load(open(f"gpurun_out/r2_ab_{n}.json"))
        print(n, "qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"], json.dumps(d["roofline"]["other_kernels_ms_per_step"]), "list_tc", d["roofline"]["avg_launch_ms"], "frac", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_B.csv python bench.py --no-extras --no-cpu --no-recall --law rank16 --steps 2 --warmup 1 > gpurun_out/r2_ncu_B.log 2>&1
tail -2 gpurun_out/r2_ncu_B.log | cut -c1-200
timeout 900 python tools/hnsw_quality.py 0 1 > gpurun_out/r2_hnsw_quality.jsonl 2> gpurun_out/r2_hnsw_quality.err; cat gpurun_out/r2_hnsw_quality.jsonl; tail -3 gpurun_out/r2_hnsw_quality.err
timeout 900 python bench.py --config C > gpurun_out/r2_bench_C2.json 2> gpurun_out/r2_bench_C2.err; cut -c1-200 gpurun_out/r2_bench_C2.json
timeout 900 python bench.py --config E > gpurun_out/r2_bench_E2.json 2> gpurun_out/r2_bench_E2.err; cut -c1-200 gpurun_out/r2_bench_E2.json
timeout 900 python bench.py --config D > gpurun_out/r2_bench_D.json 2> gpurun_out/r2_bench_D.err; tail -3 gpurun_out/r2_bench_D.err; cut -c1-300 gpurun_out/r2_bench_D.json
timeout 1200 python tools/hnsw_quality.py 2 3 >> gpurun_out/r2_hnsw_quality.jsonl 2>> gpurun_out/r2_hnsw_quality.err; tail -2 gpurun_out/r2_hnsw_quality.jsonl
