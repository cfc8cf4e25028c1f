#!/bin/bash
# round-2 session 13 (1 GPU): full GPU suite on the new defaults (one-word gathers + speculative scoring for narrow HNSW rows, lane-parallel
# addresses in select_refine), config B / C / E lines, the end-to-end diagnostic with and without the clock sampler, A/B of the speculation
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > $O/r2_tests13.log 2>&1; echo "pytest exit $?" >> $O/r2_tests13.log; tail -6 $O/r2_tests13.log
timeout 900 python bench.py > $O/r2_bench_B_v3.json 2> $O/r2_bench_B_v3.err; echo "B exit $?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_B_v3.json"))
    print("B qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],4), "clocks", d["clocks"], "list_tc", d["roofline"].get("avg_launch_ms"), "frac", d["roofline"]["frac"], json.dumps(d["roofline"]["other_kernels_ms_per_step"]), "parity", d["parity"])
except Exception as e:
    print("B failed", e)
PY
timeout 600 python tools/diag_e2e.py > $O/r2_diag_e2e_v2.json 2> $O/r2_diag_e2e_v2.err; echo "diag exit $?"; tail -2 $O/r2_diag_e2e_v2.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_diag_e2e_v2.json"))
    for k,v in d.items():
        if isinstance(v,dict) and "event_ms_per_step" in v: print(f"{k:70s} wall {v['wall_ms_per_step']:.4f} event {v['event_ms_per_step']:.4f}")
    print(d.get("sampler"))
except Exception as e:
    print("diag failed", e)
PY
for c in C E; do
  timeout 900 python bench.py --config $c > $O/r2_bench_${c}_v2.json 2> $O/r2_bench_${c}_v2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_${c}_v2.json"))
    print("$c","qps",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(d["roofline"]["frac"],3),"recall",d.get("recall_at_10"),"parity",d["parity"]["queries_with_identical_ids"],d["parity"]["n_dist_equal_on_identical_walks"],"build_s",round(d["build"]["seconds"],2),"cpu",d["cpu_baseline"] and round(d["cpu_baseline"]["value"]))
except Exception as e:
    print("$c","failed",e)
PY
done
touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu
VB_NVCC_DEFS="VB_AB_SPEC=0" python -m pgvector_b200.build > /dev/null 2> $O/r2_s13_build_spec0.err
timeout 900 python bench.py --config E --no-cpu > $O/r2_s13_spec0_E.json 2> $O/r2_s13_spec0_E.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_s13_spec0_E.json"))
    print("SPEC=0 E","qps",round(d["value"]),"frac",round(d["roofline"]["frac"],3),"build_s",round(d["build"]["seconds"],2))
except Exception as e:
    print("spec0 failed",e)
PY
touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu; python -m pgvector_b200.build > /dev/null 2>&1
