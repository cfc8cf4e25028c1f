#!/bin/bash
# round-2 session 5 (1 GPU): iterative-scan + refactored search tests, seeding after the bank-conflict fix, register-target A/B of the HNSW kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_hnsw_build.py tests/test_gpu_kmeans.py tests/test_gpu_ops.py -m gpu -q --timeout=600 > gpurun_out/r2_tests5.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests5.log; tail -12 gpurun_out/r2_tests5.log
timeout 600 python tools/diag_build_phases.py > gpurun_out/r2_diag_build2.json 2> gpurun_out/r2_diag_build2.err; cat gpurun_out/r2_diag_build2.json; tail -2 gpurun_out/r2_diag_build2.err
ROUNDS=100 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_build2.csv python tools/diag_build_phases.py > gpurun_out/r2_ncu_build2.log 2>&1
for m in 6 0 4 5 8; do
  touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu
  VB_NVCC_DEFS="VB_HNSW_MINB=$m" python -m pgvector_b200.build > /dev/null 2> gpurun_out/r2_minb_build$m.err
  timeout 600 python bench.py --config C --no-cpu > gpurun_out/r2_minb${m}_C.json 2> gpurun_out/r2_minb${m}_C.err
  timeout 900 python bench.py --config E --no-cpu > gpurun_out/r2_minb${m}_E.json 2> gpurun_out/r2_minb${m}_E.err
done
python - <<'PY'
import json
for m in (6,0,4,5,8):
    for c in "CE":
        try:
            d=json.load(open(f"gpurun_out/r2_minb{m}_{c}.json"))
            print("minb",m,c,"qps",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(d["roofline"]["frac"],3),"recall",d.get("parity",{}).get("recall_at_10"))
        except Exception as e:
            print(m,c,"failed",e)
PY
touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu; python -m pgvector_b200.build > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 1 -o gpurun_out/r2_hnsw_E -f python bench.py --config E --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_ncu_hnswE.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 1 -o gpurun_out/r2_hnsw_C2 -f python bench.py --config C --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_ncu_hnswC2.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null
