#!/bin/bash
# Ninth-pass GPU session: tensor-core centre scan, HNSW single-wave grid; full suite + headline + HNSW rows + launch list.
TAG=${1:-r1m}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_$TAG.log | cut -c1-220
timeout 600 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_default.json")); r = d["roofline"]
    print("default qps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "scan_ms", round(r["avg_launch_ms"], 3), r["other_kernels_ms_per_step"],
          "dram", r.get("dram"), "fallbacks", r.get("certificate_fallback_queries"), "cpu", (d.get("cpu_baseline") or {}).get("gpu_vs_oracle_id_agreement"), d["clocks"])
except Exception as e:
    print("unreadable", e)
PY
tail -2 gpurun_out/bench_${TAG}_default.err
timeout 400 python tools/bench_extra.py hnsw --elem halfvec --rows 100000 --ef 100 > gpurun_out/extra_hnsw_half_$TAG.json 2> gpurun_out/extra_hnsw_half_$TAG.err; cut -c1-700 gpurun_out/extra_hnsw_half_$TAG.json; tail -2 gpurun_out/extra_hnsw_half_$TAG.err
timeout 400 python tools/bench_extra.py hnsw --elem bit --rows 200000 --ef 200 > gpurun_out/extra_hnsw_bit_$TAG.json 2> gpurun_out/extra_hnsw_bit_$TAG.err; cut -c1-700 gpurun_out/extra_hnsw_bit_$TAG.json; tail -2 gpurun_out/extra_hnsw_bit_$TAG.err
B="python bench.py --queries 4096 --batch 2048 --steps 2 --warmup 1 --no-cpu --no-recall"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:vb:: --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -1 gpurun_out/ncu_launches_$TAG.log | cut -c1-200
