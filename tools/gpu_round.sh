#!/bin/bash
# One GPU session: tests, headline bench on both scan kernels, secondary benches, ncu captures.
# Usage (through gpurun):  bash tools/gpu_round.sh <tag> [quick]
# Everything lands in gpurun_out/ (merged back by gpurun); summaries are made on the CPU box with
# tools/summarize_ncu.py and committed under profiles/.
TAG=${1:-r1}
QUICK=${2:-}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }

echo "== pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1
tail -3 gpurun_out/pytest_$TAG.log | cut -c1-300

show() {
python - "$1" <<'EOF'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]
    print(sys.argv[1], "qps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(r["frac"], 3), "scan_ms", round(r["avg_launch_ms"], 2),
          "GB/launch", round(r["bytes_per_launch"] / 1e9, 1), "other", {k: round(v, 2) for k, v in r["other_kernels_ms_per_step"].items()},
          "recall", d["recall_at_10"], "|", d["config"]["index_build"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
EOF
}

echo "== headline bench, both scan kernels"
for I in 0 1; do
  timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --scan-impl $I > gpurun_out/bench_${TAG}_impl$I.json 2> gpurun_out/bench_${TAG}_impl$I.err
  show gpurun_out/bench_${TAG}_impl$I.json; tail -2 gpurun_out/bench_${TAG}_impl$I.err
done

if [ -z "$QUICK" ]; then
echo "== secondary benches"
timeout 300 python tools/bench_extra.py assign --rows 1250000 --k 4096 > gpurun_out/extra_assign_$TAG.json 2> gpurun_out/extra_assign_$TAG.err; cut -c1-900 gpurun_out/extra_assign_$TAG.json; tail -2 gpurun_out/extra_assign_$TAG.err
timeout 400 python tools/bench_extra.py hnsw --elem halfvec --rows 50000 --ef 100 > gpurun_out/extra_hnsw_half_$TAG.json 2> gpurun_out/extra_hnsw_half_$TAG.err; cut -c1-900 gpurun_out/extra_hnsw_half_$TAG.json; tail -2 gpurun_out/extra_hnsw_half_$TAG.err
timeout 400 python tools/bench_extra.py hnsw --elem bit --rows 100000 --ef 200 > gpurun_out/extra_hnsw_bit_$TAG.json 2> gpurun_out/extra_hnsw_bit_$TAG.err; cut -c1-900 gpurun_out/extra_hnsw_bit_$TAG.json; tail -2 gpurun_out/extra_hnsw_bit_$TAG.err
timeout 200 python tools/bench_extra.py exact > gpurun_out/extra_exact_$TAG.json 2> gpurun_out/extra_exact_$TAG.err; cut -c1-600 gpurun_out/extra_exact_$TAG.json

echo "== ncu"
B="python bench.py --queries 4096 --batch 2048 --steps 2 --warmup 1 --no-cpu --no-recall"
# launch list of the timed command (all kernels of this library)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:vb:: --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches_$TAG.log 2>&1
# the dominant kernel, once per variant: skip the k-means++ seeding scans (lists - 1 of them) and the warm-up step
for I in 0 1; do
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"scan_kernel|scan_bulk_kernel" -s 1003 -c 2 \
      -o gpurun_out/prof_listscan_impl${I}_$TAG -f $B --scan-impl $I > gpurun_out/ncu_listscan_impl${I}_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_listscan_impl${I}_$TAG.log | cut -c1-160
done
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:assign_tc_kernel -s 1 -c 1 \
    -o gpurun_out/prof_assign_tc_$TAG -f python tools/bench_extra.py assign --rows 300000 --k 4096 > gpurun_out/ncu_assign_$TAG.log 2>&1
tail -1 gpurun_out/ncu_assign_$TAG.log | cut -c1-160
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:hnsw_search_kernel -s 2 -c 1 \
    -o gpurun_out/prof_hnsw_$TAG -f python tools/bench_extra.py hnsw --elem halfvec --rows 20000 --ef 100 --queries 4096 > gpurun_out/ncu_hnsw_$TAG.log 2>&1
tail -1 gpurun_out/ncu_hnsw_$TAG.log | cut -c1-160
fi
ls -la gpurun_out/ | tail -30
