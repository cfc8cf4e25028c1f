#!/bin/bash
# round-2 session 21 (1 GPU): bucketed visited set + dynamic query hand-out (the new defaults) against the previous kernel,
# variants built beforehand by tools/build_variants.sh (no compile time on the GPU box)
mkdir -p gpurun_out
O=gpurun_out
V=pgvector_b200/variants
use() { cp $V/libvecb200_$1.so pgvector_b200/libvecb200.so; }
use default
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > $O/r2_s21_tests.log 2>&1; echo "pytest exit $?" >> $O/r2_s21_tests.log; tail -4 $O/r2_s21_tests.log
run() {  # variant config
  use $1
  timeout 500 python bench.py --config $2 --no-cpu > $O/r2_s21_$1_$2.json 2> $O/r2_s21_$1_$2.err; echo "$1 $2 exit $?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_s21_$1_$2.json"))
    print("$1","$2","qps",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(d["roofline"]["frac"],3),"kern_ms",round(d["roofline"]["avg_launch_ms"],3),"recall",d.get("recall_at_10"),"parity",d["parity"]["queries_with_identical_ids"],d["parity"]["n_dist_equal_on_identical_walks"],"build_s",round(d["build"]["seconds"],2))
except Exception as e:
    print("$1","$2","failed",e)
PY
}
run default E
run base E
run default C
run base C
run wide3 C
use default
