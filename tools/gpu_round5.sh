#!/bin/bash
# Fifth-pass GPU session: list_tile_kernel v2 (row per lane) -- parity, headline bench, full ncu capture.
TAG=${1:-r1i}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_ivfflat.py -x -q --timeout 120 > gpurun_out/pytest_ivf_$TAG.log 2>&1; echo "pytest ivf rc=$?"; tail -4 gpurun_out/pytest_ivf_$TAG.log
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]
    print(sys.argv[1], "qps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(r["frac"], 3), "scan_ms", round(r["avg_launch_ms"], 2),
          "GB/launch", round(r["bytes_per_launch"] / 1e9, 1), "other", {k: round(v, 2) for k, v in r["other_kernels_ms_per_step"].items()},
          "recall", d["recall_at_10"], "| cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("gpu_vs_oracle_id_agreement"),
          "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
timeout 600 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err
show gpurun_out/bench_${TAG}_default.json; tail -2 gpurun_out/bench_${TAG}_default.err
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --batch 8192 --queries 16384 > gpurun_out/bench_${TAG}_b8192.json 2> gpurun_out/bench_${TAG}_b8192.err
show gpurun_out/bench_${TAG}_b8192.json; tail -2 gpurun_out/bench_${TAG}_b8192.err
timeout 400 python tools/bench_extra.py ivf --elem halfvec > gpurun_out/extra_ivf_halfvec_$TAG.json 2> gpurun_out/extra_ivf_halfvec_$TAG.err
cut -c1-1000 gpurun_out/extra_ivf_halfvec_$TAG.json; tail -2 gpurun_out/extra_ivf_halfvec_$TAG.err
timeout 400 python tools/bench_extra.py kmeans --k 4096 --iters 10 > gpurun_out/extra_kmeans_$TAG.json 2> gpurun_out/extra_kmeans_$TAG.err
cut -c1-1000 gpurun_out/extra_kmeans_$TAG.json; tail -2 gpurun_out/extra_kmeans_$TAG.err
echo "== ncu"
B="python bench.py --queries 4096 --batch 2048 --steps 2 --warmup 1 --no-cpu --no-recall"
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:list_tile_kernel -s 4 -c 1 \
    -o gpurun_out/prof_listtile_$TAG -f $B > gpurun_out/ncu_listtile_$TAG.log 2>&1
tail -1 gpurun_out/ncu_listtile_$TAG.log | cut -c1-200
