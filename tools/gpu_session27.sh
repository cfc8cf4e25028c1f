#!/bin/bash
# round-2 session 27 (1 GPU, short): the final tree after the last two changes (Lloyd state in one arena, plain radix histogram)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=600 > $O/r2_tests_final2.log 2>&1; echo "pytest exit $?" >> $O/r2_tests_final2.log; tail -3 $O/r2_tests_final2.log | cut -c1-200
timeout 300 python bench.py --config D > $O/r2_bench_D_final2.json 2> $O/r2_bench_D_final2.err; echo "D exit $?"
timeout 300 python bench.py --no-cpu --law rank16 > $O/r2_bench_B_final2.json 2> $O/r2_bench_B_final2.err; echo "B exit $?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_D_final2.json")); print("D", round(d["value"]), d["phases_s"], d["roofline"]["frac"], d.get("recall_at_10"))
except Exception as e: print("D failed", e)
try:
    d=json.load(open("gpurun_out/r2_bench_B_final2.json")); s=d["batch_sweep"]
    print("B", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["roofline"]["other_kernels_ms_per_step"], [(x["batch"], round(x["ms_per_batch"],4)) for x in s["device_resident"]], s["single_query"]["latency_us_p50"], s["single_query"]["one_call_vb_ivf_search_latency_us"], s["single_query"]["general_path"])
except Exception as e: print("B failed", e)
PY
