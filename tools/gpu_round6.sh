#!/bin/bash
# Sixth-pass GPU session: tensor-core filter of the batched list scan (vb_list_tc.cu) -- parity first, then bench + ncu.
TAG=${1:-r1j}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 400 python -m pytest tests/test_gpu_ivfflat.py -x -q --timeout 100 > gpurun_out/pytest_ivf_$TAG.log 2>&1; RC=$?; echo "pytest ivf rc=$RC"; tail -25 gpurun_out/pytest_ivf_$TAG.log | cut -c1-200
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]
    print(sys.argv[1], "qps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(r["frac"], 3), "scan_ms", round(r["avg_launch_ms"], 2),
          "other", {k: round(v, 2) for k, v in r["other_kernels_ms_per_step"].items()}, "fallbacks", r.get("certificate_fallback_queries"),
          "recall", d["recall_at_10"], "| cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("gpu_vs_oracle_id_agreement"),
          (d.get("cpu_baseline") or {}).get("gpu_vs_oracle_max_rel_dist_err"), "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
timeout 500 python bench.py --scan-impl 4 > gpurun_out/bench_${TAG}_tc.json 2> gpurun_out/bench_${TAG}_tc.err
show gpurun_out/bench_${TAG}_tc.json; tail -3 gpurun_out/bench_${TAG}_tc.err
timeout 400 python bench.py --scan-impl 4 --steps 5 --warmup 3 --no-cpu --batch 8192 --queries 16384 > gpurun_out/bench_${TAG}_tc_b8192.json 2> gpurun_out/bench_${TAG}_tc_b8192.err
show gpurun_out/bench_${TAG}_tc_b8192.json; tail -2 gpurun_out/bench_${TAG}_tc_b8192.err
echo "== ncu"
B="python bench.py --scan-impl 4 --queries 4096 --batch 2048 --steps 2 --warmup 1 --no-cpu --no-recall"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:vb:: --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -1 gpurun_out/ncu_launches_$TAG.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:list_tc_kernel -s 4 -c 1 \
    -o gpurun_out/prof_listtc_$TAG -f $B > gpurun_out/ncu_listtc_$TAG.log 2>&1
tail -1 gpurun_out/ncu_listtc_$TAG.log | cut -c1-200
