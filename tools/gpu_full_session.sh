#!/bin/bash
# gpurun session, ~12 GPU-minutes: full GPU suite, smoke, headline, reference arm, assign / k-means with
# the converged-leader tensor-core kernels, ncu capture + launch list of the headline command.
TAG=${1:-r1t}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_$TAG.log | cut -c1-220
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_default.json")); r = d["roofline"]
    print("default qps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "scan_ms", round(r["avg_launch_ms"], 3), r["other_kernels_ms_per_step"],
          "fallbacks", r.get("certificate_fallback_queries"), r.get("level1_fallback_queries"), "agree", (d.get("cpu_baseline") or {}).get("gpu_vs_oracle_id_agreement"),
          (d.get("cpu_baseline") or {}).get("gpu_vs_oracle_max_rel_dist_err"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d["clocks"])
except Exception as e:
    print("unreadable", e)
PY
tail -2 gpurun_out/bench_${TAG}_default.err
timeout 300 python tools/bench_extra.py assign --rows 1250000 --k 4096 > gpurun_out/extra_assign_$TAG.json 2> gpurun_out/extra_assign_$TAG.err; cut -c1-800 gpurun_out/extra_assign_$TAG.json; tail -2 gpurun_out/extra_assign_$TAG.err
timeout 400 python tools/bench_extra.py kmeans --k 4096 --iters 10 > gpurun_out/extra_kmeans_$TAG.json 2> gpurun_out/extra_kmeans_$TAG.err; cut -c1-800 gpurun_out/extra_kmeans_$TAG.json; tail -2 gpurun_out/extra_kmeans_$TAG.err
timeout 400 python tools/bench_extra.py ivf --elem halfvec > gpurun_out/extra_ivf_halfvec_$TAG.json 2> gpurun_out/extra_ivf_halfvec_$TAG.err; cut -c1-600 gpurun_out/extra_ivf_halfvec_$TAG.json; tail -2 gpurun_out/extra_ivf_halfvec_$TAG.err
B="python bench.py --queries 4096 --batch 2048 --steps 2 --warmup 1 --no-cpu --no-recall"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:list_tc_kernel -s 9 -c 1 \
    -o gpurun_out/prof_listtc_l1_$TAG -f $B > gpurun_out/ncu_listtc_l1_$TAG.log 2>&1
tail -1 gpurun_out/ncu_listtc_l1_$TAG.log | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:vb:: --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -1 gpurun_out/ncu_launches_$TAG.log | cut -c1-200
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
cut -c1-300 gpurun_out/bench_${TAG}_reference.json
