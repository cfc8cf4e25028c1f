#!/bin/bash
# Second-pass GPU session: headline bench on the balanced data law (three scan-kernel settings),
# ncu captures of the list scan, assign launch list.
TAG=${1:-r1c}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]
    print(sys.argv[1], "qps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(r["frac"], 3), "scan_ms", round(r["avg_launch_ms"], 2),
          "GB/launch", round(r["bytes_per_launch"] / 1e9, 1), "other", {k: round(v, 2) for k, v in r["other_kernels_ms_per_step"].items()},
          "recall", d["recall_at_10"], "|", d["config"]["index_build"], "| cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
for I in 0 1; do
  timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --scan-impl $I > gpurun_out/bench_${TAG}_impl$I.json 2> gpurun_out/bench_${TAG}_impl$I.err
  show gpurun_out/bench_${TAG}_impl$I.json; tail -2 gpurun_out/bench_${TAG}_impl$I.err
done
# the default configuration, with the CPU baseline and recall: this is the line the driver will see
timeout 600 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err
show gpurun_out/bench_${TAG}_default.json; tail -2 gpurun_out/bench_${TAG}_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
cut -c1-700 gpurun_out/bench_${TAG}_reference.json; tail -2 gpurun_out/bench_${TAG}_reference.err

echo "== ncu"
B="python bench.py --queries 4096 --batch 2048 --steps 2 --warmup 1 --no-cpu --no-recall"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:vb:: --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches_$TAG.log 2>&1
for I in 0 1; do
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"scan_kernel|scan_bulk_kernel" -s 1002 -c 1 \
      -o gpurun_out/prof_listscan_impl${I}_$TAG -f $B --scan-impl $I > gpurun_out/ncu_listscan_impl${I}_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_listscan_impl${I}_$TAG.log | cut -c1-160
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:vb:: --csv --log-file gpurun_out/launches_assign_$TAG.csv \
    python tools/bench_extra.py assign --rows 600000 --k 4096 > gpurun_out/ncu_assign_launches_$TAG.log 2>&1
timeout 300 python tools/bench_extra.py assign --rows 1250000 --k 4096 > gpurun_out/extra_assign_$TAG.json 2> gpurun_out/extra_assign_$TAG.err; cut -c1-900 gpurun_out/extra_assign_$TAG.json
timeout 400 python tools/bench_extra.py hnsw --elem halfvec --rows 100000 --ef 100 > gpurun_out/extra_hnsw_half_$TAG.json 2> gpurun_out/extra_hnsw_half_$TAG.err; cut -c1-900 gpurun_out/extra_hnsw_half_$TAG.json; tail -2 gpurun_out/extra_hnsw_half_$TAG.err
timeout 400 python tools/bench_extra.py hnsw --elem bit --rows 200000 --ef 200 > gpurun_out/extra_hnsw_bit_$TAG.json 2> gpurun_out/extra_hnsw_bit_$TAG.err; cut -c1-900 gpurun_out/extra_hnsw_bit_$TAG.json; tail -2 gpurun_out/extra_hnsw_bit_$TAG.err
ls gpurun_out | wc -l
