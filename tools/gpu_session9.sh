#!/bin/bash
# round-2 session 9 (1 GPU): the evidence run — bench lines of every config, launch list, ncu captures, full GPU test suite
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/r2_gpu.txt
timeout 900 python bench.py > $O/r2_bench_B.json 2> $O/r2_bench_B.err; echo "B exit $?"; tail -2 $O/r2_bench_B.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_B.json"))
    print("B qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"], json.dumps(d["roofline"].get("other_kernels_ms_per_step")), "list_tc", d["roofline"].get("avg_launch_ms"), "frac", d["roofline"]["frac"], "recall", d.get("recall_at_10"), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]))
    print("north_star", json.dumps(d.get("north_star_kernel"))[:400])
    print("sweep", json.dumps(d.get("batch_sweep"))[:600])
    print("laws", json.dumps(d.get("laws"))[:600])
except Exception as e:
    print("B failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_B.csv python bench.py --no-extras --no-cpu --no-recall --law rank16 --steps 2 --warmup 1 > $O/r2_ncu_launches_B.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:list_tc_kernel -s 2 -c 2 -o $O/r2_listtc -f python bench.py --no-extras --no-cpu --no-recall --law rank16 --steps 2 --warmup 1 > $O/r2_ncu_listtc.log 2>&1; echo "ncu list_tc exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_bulk_kernel -s 1 -c 2 -o $O/r2_scanbulk -f python bench.py --scan-impl 1 --no-extras --no-cpu --no-recall --law rank16 --steps 2 --warmup 1 > $O/r2_ncu_scanbulk.log 2>&1; echo "ncu scan_bulk exit $?"
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 --durations=12 > $O/r2_tests.log 2>&1
echo "pytest exit $?" >> $O/r2_tests.log; tail -22 $O/r2_tests.log
timeout 900 python bench.py --config C > $O/r2_bench_C.json 2> $O/r2_bench_C.err; echo "C exit $?"; tail -2 $O/r2_bench_C.err; cut -c1-1500 $O/r2_bench_C.json
timeout 1200 python bench.py --config E > $O/r2_bench_E.json 2> $O/r2_bench_E.err; echo "E exit $?"; tail -2 $O/r2_bench_E.err; cut -c1-1500 $O/r2_bench_E.json
timeout 300 python bench.py --config A > $O/r2_bench_A.json 2> $O/r2_bench_A.err; echo "A exit $?"; tail -2 $O/r2_bench_A.err; cut -c1-1200 $O/r2_bench_A.json
timeout 900 python bench.py --config D > $O/r2_bench_D.json 2> $O/r2_bench_D.err; echo "D exit $?"; tail -2 $O/r2_bench_D.err; cut -c1-1500 $O/r2_bench_D.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 1 -o $O/r2_hnsw_C -f python bench.py --config C --steps 2 --warmup 1 --no-cpu > $O/r2_ncu_hnswC.log 2>&1; echo "ncu hnsw exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_B_reference.json 2> $O/r2_bench_B_reference.err; echo "ref exit $?"; cut -c1-400 $O/r2_bench_B_reference.json
ls -la $O | tail -30
