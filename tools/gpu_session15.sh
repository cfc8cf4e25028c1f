#!/bin/bash
# round-2 session 15 (1 GPU): the CTA-per-query refine kernel -- parity tests, config B line, launch list, smoke()
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_ivfflat.py tests/test_ext_harness.py -m gpu -q --timeout=600 -x > $O/r2_tests15.log 2>&1; echo "pytest exit $?" >> $O/r2_tests15.log; tail -12 $O/r2_tests15.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > $O/r2_bench_B_v4.json 2> $O/r2_bench_B_v4.err; echo "B exit $?"; tail -2 $O/r2_bench_B_v4.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_B_v4.json"))
    print("B qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],4), "h2d", d["e2e"].get("h2d_gbs"), d["e2e"].get("pinned_near_gpu"), "launches/step", d["gpu_launches"]/d["steps"], "list_tc", d["roofline"].get("avg_launch_ms"), "frac", d["roofline"]["frac"], json.dumps(d["roofline"]["other_kernels_ms_per_step"]), "parity", d["parity"], "recall", d["recall_at_10"])
    m=d["laws"]["mixture"]; print("mixture", round(m["value"]), round(m["e2e"]["value"]), m["roofline"]["frac"], m["parity"], m["roofline"]["level1_fallback_queries"], m["roofline"]["certificate_fallback_queries"])
    print("sweep", [(x["batch"], round(x["ms_per_batch"],3)) for x in d["batch_sweep"]["device_resident"]], d["batch_sweep"]["single_query"]["latency_us_p50"])
except Exception as e:
    print("B failed", e)
PY
VB_FUSED_REFINE=1 timeout 600 python bench.py --no-extras --no-cpu --no-recall --law rank16 > $O/r2_bench_B_v4_fused1.json 2> /dev/null
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_B_v4_fused1.json"))
    print("fused_refine=1: B qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), json.dumps(d["roofline"]["other_kernels_ms_per_step"]))
except Exception as e:
    print("fused1 failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_B_v4.csv python bench.py --no-extras --no-cpu --no-recall --law rank16 --steps 2 --warmup 1 > $O/r2_ncu_launches_B_v4.log 2>&1; echo "launch list exit $?"
