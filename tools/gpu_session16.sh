#!/bin/bash
# round-2 session 16 (1 GPU): CTA-per-query refine with the parallel gather / radix direct selection: parity tests, config B, A/B against fused_refine 1
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_ivfflat.py -m gpu -q --timeout=600 -x > $O/r2_tests16.log 2>&1; echo "pytest exit $?" >> $O/r2_tests16.log; tail -8 $O/r2_tests16.log
for f in 3 1 3 1; do
VB_FUSED_REFINE=$f timeout 600 python bench.py --no-extras --no-cpu --no-recall --law rank16 > $O/r2_bench_B_v5_fused$f.json 2> /dev/null
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_B_v5_fused$f.json"))
    print("fused_refine=$f: B qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), json.dumps(d["roofline"]["other_kernels_ms_per_step"]), "launches/step", d["gpu_launches"]/d["steps"])
except Exception as e:
    print("fused $f failed", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_B_v5.csv python bench.py --no-extras --no-cpu --no-recall --law rank16 --steps 2 --warmup 1 > $O/r2_ncu_launches_B_v5.log 2>&1; echo "launch list exit $?"
