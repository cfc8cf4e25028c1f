#!/bin/bash
# round-2 session 17 (1 GPU): CTA refine with shared-memory addresses / parallel slab-minimum loads: parity tests + config B
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_ivfflat.py -m gpu -q --timeout=600 -x > $O/r2_tests17.log 2>&1; echo "pytest exit $?" >> $O/r2_tests17.log; tail -6 $O/r2_tests17.log
for f in 3 3; do
timeout 600 python bench.py --no-extras --no-cpu --no-recall --law rank16 > $O/r2_bench_B_v6.json 2> /dev/null
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_B_v6.json"))
    print("B qps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), json.dumps(d["roofline"]["other_kernels_ms_per_step"]), "launches/step", d["gpu_launches"]/d["steps"])
except Exception as e:
    print("failed", e)
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cta_refine_kernel -s 5 -c 2 -o $O/r2_ctarefine -f python bench.py --no-extras --no-cpu --no-recall --law rank16 --steps 2 --warmup 1 > $O/r2_ncu_ctarefine.log 2>&1; echo "ncu exit $?"
