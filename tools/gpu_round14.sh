#!/bin/bash
# pipelined host path: parity + the headline line
TAG=${1:-r1x}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_ivfflat.py -x -q --timeout 100 > gpurun_out/pytest_ivf_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_ivf_$TAG.log | cut -c1-200
timeout 400 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_default.json")); r = d["roofline"]
    print("default qps", round(d["value"]), "e2e", d["e2e"], "ms/step", round(d["ms_per_step"], 3), "scan_ms", round(r["avg_launch_ms"], 3),
          "agree", (d.get("cpu_baseline") or {}).get("gpu_vs_oracle_id_agreement"), d["clocks"])
except Exception as e:
    print("unreadable", e)
PY
tail -3 gpurun_out/bench_${TAG}_default.err
