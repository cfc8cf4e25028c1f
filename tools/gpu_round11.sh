#!/bin/bash
# Eleventh-pass GPU session: two-level tensor-core filter (hi plane first), HNSW with 8 gathers in flight.
TAG=${1:-r1o}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest_$TAG.log | cut -c1-220
for ARGS in "" "--components 1000 --no-cpu"; do
  N=$(echo "$ARGS" | tr -cd 'a-z0-9' | cut -c1-12); N=${N:-default}
  timeout 600 python bench.py $ARGS > gpurun_out/bench_${TAG}_$N.json 2> gpurun_out/bench_${TAG}_$N.err
  python - "$N" <<PY
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/bench_${TAG}_%s.json" % n)); r = d["roofline"]
    print(n, "qps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "scan_ms", round(r["avg_launch_ms"], 3), r["other_kernels_ms_per_step"],
          "fallbacks", r.get("certificate_fallback_queries"), "l1 fallbacks", r.get("level1_fallback_queries"), "recall", d["recall_at_10"], "agree", (d.get("cpu_baseline") or {}).get("gpu_vs_oracle_id_agreement"),
          (d.get("cpu_baseline") or {}).get("gpu_vs_oracle_max_rel_dist_err"), d["clocks"])
except Exception as e:
    print(n, "unreadable", e)
PY
  tail -2 gpurun_out/bench_${TAG}_$N.err
done
B="python bench.py --queries 4096 --batch 2048 --steps 2 --warmup 1 --no-cpu --no-recall"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:list_tc_kernel -s 9 -c 1 \
    -o gpurun_out/prof_listtc_l1_$TAG -f $B > gpurun_out/ncu_listtc_l1_$TAG.log 2>&1
tail -1 gpurun_out/ncu_listtc_l1_$TAG.log | cut -c1-200
timeout 400 python tools/bench_extra.py hnsw --elem halfvec --rows 100000 --ef 100 > gpurun_out/extra_hnsw_half_$TAG.json 2> gpurun_out/extra_hnsw_half_$TAG.err; cut -c1-420 gpurun_out/extra_hnsw_half_$TAG.json; tail -2 gpurun_out/extra_hnsw_half_$TAG.err
