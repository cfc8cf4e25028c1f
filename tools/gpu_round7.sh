#!/bin/bash
# Seventh-pass GPU session: everything at the new defaults (automatic = tensor-core filter for batched small-k scans).
TAG=${1:-r1k}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err
cut -c1-3000 gpurun_out/bench_${TAG}_default.json; tail -2 gpurun_out/bench_${TAG}_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
cut -c1-600 gpurun_out/bench_${TAG}_reference.json; tail -2 gpurun_out/bench_${TAG}_reference.err
for E in halfvec; do
  timeout 400 python tools/bench_extra.py ivf --elem $E > gpurun_out/extra_ivf_${E}_$TAG.json 2> gpurun_out/extra_ivf_${E}_$TAG.err
  cut -c1-900 gpurun_out/extra_ivf_${E}_$TAG.json; tail -2 gpurun_out/extra_ivf_${E}_$TAG.err
done
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --batch 512 > gpurun_out/bench_${TAG}_b512.json 2> gpurun_out/bench_${TAG}_b512.err
python - <<'PY'
import json
for n in ("b512",):
    try:
        d = json.load(open("gpurun_out/bench_r1k_%s.json" % n)); r = d["roofline"]
        print(n, "qps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "scan_ms", round(r["avg_launch_ms"], 3), r["other_kernels_ms_per_step"])
    except Exception as e:
        print(n, "unreadable", e)
PY
