#!/bin/bash
# round-2 session 22 (1 GPU): the fused one-query IVFFlat kernels (tests, latency against the general path), and line-level
# ncu captures of hnsw_search_kernel with the bucketed visited set (configs E and C)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ivf_one.py -m gpu -q --timeout=300 -x > $O/r2_s22_tests_one.log 2>&1; echo "one-query tests exit $?"; tail -15 $O/r2_s22_tests_one.log
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > $O/r2_s22_tests.log 2>&1; echo "pytest exit $?" >> $O/r2_s22_tests.log; tail -6 $O/r2_s22_tests.log
timeout 600 python bench.py --no-cpu --law rank16 > $O/r2_s22_B.json 2> $O/r2_s22_B.err; echo "B exit $?"; tail -2 $O/r2_s22_B.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_s22_B.json"))
    print("B qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"], "frac", d["roofline"]["frac"], "parity", d.get("parity"))
    print("sweep", [(x["batch"], round(x["ms_per_batch"],4)) for x in d["batch_sweep"]["device_resident"]])
    print("single", json.dumps(d["batch_sweep"]["single_query"]))
except Exception as e:
    print("B failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 1 -o $O/r2_hnsw_E_v2 -f python bench.py --config E --steps 2 --warmup 1 --no-cpu > $O/r2_ncu_hnswE_v2.log 2>&1; echo "ncu E exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 1 -o $O/r2_hnsw_C_v2 -f python bench.py --config C --steps 2 --warmup 1 --no-cpu > $O/r2_ncu_hnswC_v2.log 2>&1; echo "ncu C exit $?"
ls -la $O/*.ncu-rep
