#!/bin/bash
# round-2 session 12 (1 GPU): config B with the 200 ms clock sampler; HNSW A/B: one-word-per-lane gathers for bit(1024) rows, neighbour-list prefetch; ncu of config E
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py > $O/r2_bench_B_v2.json 2> $O/r2_bench_B_v2.err; echo "B exit $?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_B_v2.json"))
    print("B qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],4), "clocks", d["clocks"], "list_tc", d["roofline"].get("avg_launch_ms"), "frac", d["roofline"]["frac"])
except Exception as e:
    print("B failed", e)
PY
ab() {  # name defs configs
  touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu
  VB_NVCC_DEFS="$2" python -m pgvector_b200.build > /dev/null 2> $O/r2_s12_build_$1.err
  timeout 600 python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_hnsw_build.py -m gpu -q --timeout=300 -x > $O/r2_s12_tests_$1.log 2>&1; echo "$1 pytest exit $?"; tail -2 $O/r2_s12_tests_$1.log
  for c in $3; do
    timeout 900 python bench.py --config $c --no-cpu > $O/r2_s12_$1_$c.json 2> $O/r2_s12_$1_$c.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_s12_$1_$c.json"))
    print("$1","$c","qps",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(d["roofline"]["frac"],3),"recall",d.get("recall_at_10"),"parity",d["parity"]["queries_with_identical_ids"],"build_s",round(d["build"]["seconds"],2))
except Exception as e:
    print("$1","$c","failed",e)
PY
  done
}
ab lpr8 "VB_AB_LPR_V8=8" "E"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 1 -o $O/r2_hnsw_E -f python bench.py --config E --steps 2 --warmup 1 --no-cpu > $O/r2_ncu_hnswE.log 2>&1; echo "ncu E exit $?"
ab lpr8_nbrpf "VB_AB_LPR_V8=8 VB_AB_NBRPF=1" "C E"
touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu; python -m pgvector_b200.build > /dev/null 2>&1
ls -la $O/*.ncu-rep | tail -3
