#!/bin/bash
# round-2 session 20 (1 GPU): config C with all three words per lane of W rows in flight (VB_AB_WIDE1), with and without a 96-register target
mkdir -p gpurun_out
O=gpurun_out
ab() {  # name defs
  touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu
  VB_NVCC_DEFS="$2" python -m pgvector_b200.build > /dev/null 2> $O/r2_s20_build_$1.err
  timeout 600 python -m pytest tests/test_gpu_hnsw.py -m gpu -q --timeout=300 -x > $O/r2_s20_tests_$1.log 2>&1; echo "$1 pytest exit $?"; tail -1 $O/r2_s20_tests_$1.log
  timeout 600 python bench.py --config C --no-cpu > $O/r2_s20_$1_C.json 2> $O/r2_s20_$1_C.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_s20_$1_C.json"))
    print("$1","C","qps",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(d["roofline"]["frac"],3),"recall",d.get("recall_at_10"),"parity",d["parity"]["queries_with_identical_ids"],"build_s",round(d["build"]["seconds"],2))
except Exception as e:
    print("$1","C","failed",e)
PY
}
ab base ""
ab wide3 "VB_AB_WIDE1=3"
ab wide4 "VB_AB_WIDE1=4"
ab wide4_minb5 "VB_AB_WIDE1=4 VB_HNSW_MINB=5"
ab rpi2_minb8 "VB_AB_RPI_WIDE=2 VB_HNSW_MINB=8"
ab rpi3_minb7 "VB_AB_RPI_WIDE=3 VB_HNSW_MINB=7"
touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu; python -m pgvector_b200.build > /dev/null 2>&1
timeout 300 python tools/bench_extra.py sparse > $O/r2_extra_sparse.json 2> $O/r2_extra_sparse.err; cut -c1-900 $O/r2_extra_sparse.json; tail -2 $O/r2_extra_sparse.err
