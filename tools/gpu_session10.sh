#!/bin/bash
# round-2 session 10 (1 GPU): sparsevec parity tests, where the end-to-end step goes, HNSW register target 8 CTAs / SM
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q --timeout=300 > $O/r2_tests_sparse.log 2>&1; echo "pytest exit $?" >> $O/r2_tests_sparse.log; tail -25 $O/r2_tests_sparse.log
timeout 600 python tools/diag_e2e.py > $O/r2_diag_e2e.json 2> $O/r2_diag_e2e.err; echo "diag exit $?"; tail -3 $O/r2_diag_e2e.err; cat $O/r2_diag_e2e.json
touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu
VB_NVCC_DEFS="VB_HNSW_MINB=8" python -m pgvector_b200.build > /dev/null 2> $O/r2_s10_build_minb8.err
grep -A2 "hnsw_search_kernel" pgvector_b200/build/ptxas.log | grep -E "registers|spill" | tail -4
for c in C E; do
  timeout 900 python bench.py --config $c --no-cpu > $O/r2_s10_minb8_$c.json 2> $O/r2_s10_minb8_$c.err
done
python - <<'PY'
import json
for c in "CE":
    try:
        d=json.load(open(f"gpurun_out/r2_s10_minb8_{c}.json"))
        print("MINB=8",c,"qps",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(d["roofline"]["frac"],3),"recall",d.get("recall_at_10"),"build_s",round(d["build"]["seconds"],2))
    except Exception as e:
        print(c,"failed",e)
PY
