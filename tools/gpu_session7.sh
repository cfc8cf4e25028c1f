#!/bin/bash
# round-2 session 7 (1 GPU): which of the HNSW kernel changes pay (one switch off at a time), slab selection after the parallel prologue
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_hnsw_build.py tests/test_gpu_headline.py -m gpu -q --timeout=900 > gpurun_out/r2_tests7.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests7.log; tail -8 gpurun_out/r2_tests7.log
timeout 600 python bench.py --no-cpu > gpurun_out/r2_bench_B7.json 2> gpurun_out/r2_bench_B7.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_B7.json"))
    print("B qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"], json.dumps(d["roofline"]["other_kernels_ms_per_step"]), "list_tc", round(d["roofline"]["avg_launch_ms"],4), "frac", round(d["roofline"]["frac"],3), "recall", d.get("recall_at_10"))
except Exception as e:
    print("B failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_B3.csv python bench.py --no-extras --no-cpu --no-recall --law rank16 --steps 2 --warmup 1 > gpurun_out/r2_ncu_B3.log 2>&1
for v in ALLON VB_AB_PACKED VB_AB_PINGPONG VB_AB_RANKSORT VB_AB_INPLACE VB_AB_VIS; do
  touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu
  if [ "$v" = "ALLON" ]; then defs=""; else defs="$v=0"; fi
  VB_NVCC_DEFS="$defs" python -m pgvector_b200.build > /dev/null 2> gpurun_out/r2_ab_build_$v.err
  timeout 600 python bench.py --config C --no-cpu > gpurun_out/r2_ab_${v}_C.json 2> gpurun_out/r2_ab_${v}_C.err
  timeout 900 python bench.py --config E --no-cpu > gpurun_out/r2_ab_${v}_E.json 2> gpurun_out/r2_ab_${v}_E.err
done
python - <<'PY'
import json
for v in ("ALLON","VB_AB_PACKED","VB_AB_PINGPONG","VB_AB_RANKSORT","VB_AB_INPLACE","VB_AB_VIS"):
    for c in "CE":
        try:
            d=json.load(open(f"gpurun_out/r2_ab_{v}_{c}.json"))
            print(v+("=0" if v!="ALLON" else ""),c,"qps",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(d["roofline"]["frac"],3),"recall",d.get("recall_at_10"))
        except Exception as e:
            print(v,c,"failed",e)
PY
touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu; python -m pgvector_b200.build > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 1 -o gpurun_out/r2_hnsw_C3 -f python bench.py --config C --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_ncu_hnswC3.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -3
