#!/bin/bash
# gpurun session, ~1.5 GPU-minutes: IVFFlat parity, the headline bench, and a workload that needs filter level 2
TAG=${1:-r1q}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_ivfflat.py -x -q --timeout 120 > gpurun_out/pytest_ivf_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_ivf_$TAG.log | cut -c1-220
for ARGS in "--no-cpu" "--components 1000 --no-cpu --no-recall"; do
  N=$(echo "$ARGS" | tr -cd 'a-z0-9' | cut -c1-12)
  timeout 600 python bench.py $ARGS > gpurun_out/bench_${TAG}_$N.json 2> gpurun_out/bench_${TAG}_$N.err
  python - "$N" <<PY
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/bench_${TAG}_%s.json" % n)); r = d["roofline"]
    print(n, "qps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "scan_ms", round(r["avg_launch_ms"], 3), r["other_kernels_ms_per_step"],
          "fallbacks", r.get("certificate_fallback_queries"), "l1 fallbacks", r.get("level1_fallback_queries"), "recall", d["recall_at_10"], d["clocks"])
except Exception as e:
    print(n, "unreadable", e)
PY
  tail -2 gpurun_out/bench_${TAG}_$N.err
done
