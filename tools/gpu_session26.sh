#!/bin/bash
# round-2 final session (1 GPU): full GPU suite, smoke(), the bench lines of every config, launch list, reference arm -- on the final tree
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > $O/r2_tests_final.log 2>&1; echo "pytest exit $?" >> $O/r2_tests_final.log; tail -5 $O/r2_tests_final.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/r2_bench_B_final.json 2> $O/r2_bench_B_final.err; echo "B exit $?"; tail -2 $O/r2_bench_B_final.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_B_final.json"))
    print("B qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],4), "h2d", d["e2e"].get("h2d_gbs"), "launches/step", d["gpu_launches"]/d["steps"], "list_tc", d["roofline"].get("avg_launch_ms"), "frac", d["roofline"]["frac"], json.dumps(d["roofline"]["other_kernels_ms_per_step"]), "parity", d["parity"], "recall", d["recall_at_10"], "clocks", d["clocks"])
    m=d["laws"]["mixture"]; print("mixture", round(m["value"]), round(m["e2e"]["value"]), m["roofline"]["frac"], m["parity"])
    print("sweep", [(x["batch"], round(x["ms_per_batch"],3)) for x in d["batch_sweep"]["device_resident"]], json.dumps(d["batch_sweep"]["single_query"]))
    print("north", {k:(round(v["value"]), round(v["roofline"]["frac"],3)) for k,v in d["north_star_kernel"].items()}, "cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print("B failed", e)
PY
for c in C E A D; do
  timeout 600 python bench.py --config $c > $O/r2_bench_${c}_final.json 2> $O/r2_bench_${c}_final.err; echo "$c exit $?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_${c}_final.json"))
    print("$c","value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(d["roofline"]["frac"],3),"recall",d.get("recall_at_10"),"parity",d.get("parity"),"cpu",d.get("cpu_baseline") and round(d["cpu_baseline"]["value"]), d.get("phases_s"), d.get("single_query"))
except Exception as e:
    print("$c","failed",e)
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_B_final.csv python bench.py --no-extras --no-cpu --no-recall --law rank16 --steps 2 --warmup 1 > $O/r2_ncu_launches_B_final.log 2>&1; echo "launch list exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_B_reference_final.json 2> /dev/null; cut -c1-300 $O/r2_bench_B_reference_final.json
