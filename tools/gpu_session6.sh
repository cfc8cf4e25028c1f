#!/bin/bash
# round-2 session 6 (1 GPU): slab-minimum selection A/B, HNSW after the instruction / cache changes, seeding after the persistent pass
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_headline.py tests/test_gpu_ivfflat.py tests/test_gpu_hnsw.py tests/test_gpu_hnsw_build.py tests/test_gpu_kmeans.py tests/test_ext_harness.py -m gpu -q --timeout=900 > gpurun_out/r2_tests6.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests6.log; tail -12 gpurun_out/r2_tests6.log
for f in 0 1; do
  VB_SLAB_SELECT=$f timeout 600 python bench.py --no-extras --no-cpu --law rank16 > gpurun_out/r2_ab_slab$f.json 2> gpurun_out/r2_ab_slab$f.err
done
python - <<'PY'
import json
for n in ("0","1"):
    try:
        d=json.load(open(f"gpurun_out/r2_ab_slab{n}.json"))
        print("slab_select", n, "qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"], json.dumps(d["roofline"]["other_kernels_ms_per_step"]), "list_tc", round(d["roofline"]["avg_launch_ms"],4), "frac", round(d["roofline"]["frac"],3), "recall", d.get("parity",{}).get("recall_at_10"))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_B2.csv python bench.py --no-extras --no-cpu --no-recall --law rank16 --steps 2 --warmup 1 > gpurun_out/r2_ncu_B2.log 2>&1
timeout 600 python bench.py --config C --no-cpu > gpurun_out/r2_bench_C4.json 2> gpurun_out/r2_bench_C4.err; cut -c1-200 gpurun_out/r2_bench_C4.json
timeout 900 python bench.py --config E --no-cpu > gpurun_out/r2_bench_E4.json 2> gpurun_out/r2_bench_E4.err; cut -c1-200 gpurun_out/r2_bench_E4.json
touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu
VB_NVCC_DEFS="VB_HNSW_EVICT_FIRST=0" python -m pgvector_b200.build > /dev/null 2> gpurun_out/r2_ef0_build.err
timeout 600 python bench.py --config C --no-cpu > gpurun_out/r2_bench_C4_ef0.json 2> gpurun_out/r2_bench_C4_ef0.err
timeout 900 python bench.py --config E --no-cpu > gpurun_out/r2_bench_E4_ef0.json 2> gpurun_out/r2_bench_E4_ef0.err
touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu; python -m pgvector_b200.build > /dev/null 2>&1
python - <<'PY'
import json
for n in ("C4","E4","C4_ef0","E4_ef0"):
    try:
        d=json.load(open(f"gpurun_out/r2_bench_{n}.json"))
        print(n,"qps",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(d["roofline"]["frac"],3),"build_s",d.get("build_s"),"recall",d.get("recall_at_10") or d.get("parity"))
    except Exception as e:
        print(n,"failed",e)
PY
timeout 600 python tools/diag_build_phases.py > gpurun_out/r2_diag_build3.json 2> gpurun_out/r2_diag_build3.err; cat gpurun_out/r2_diag_build3.json; tail -2 gpurun_out/r2_diag_build3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 1 -o gpurun_out/r2_hnsw_E2 -f python bench.py --config E --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_ncu_hnswE2.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null
