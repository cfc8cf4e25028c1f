#!/bin/bash
# round-2 session 23 (1 GPU): compute-sanitizer on the one-query scan (the sticky "invalid address space" error of session 22),
# and the next-candidate preload of the HNSW search (VB_AB_NEXTPF) against the default on configs E and C
mkdir -p gpurun_out
O=gpurun_out
V=pgvector_b200/variants
cp $V/libvecb200_default.so pgvector_b200/libvecb200.so
timeout 300 python tools/diag_one_query.py 3 > $O/r2_s23_diag3.log 2>&1; echo "diag dim 3 exit $?"; tail -6 $O/r2_s23_diag3.log
timeout 300 python tools/diag_one_query.py 96 > $O/r2_s23_diag96.log 2>&1; echo "diag dim 96 exit $?"; tail -3 $O/r2_s23_diag96.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/diag_one_query.py 3 > $O/r2_s23_memcheck.log 2>&1; echo "memcheck exit $?"; grep -v "^$" $O/r2_s23_memcheck.log | head -70
run() {  # variant config
  cp $V/libvecb200_$1.so pgvector_b200/libvecb200.so
  timeout 500 python bench.py --config $2 --no-cpu > $O/r2_s23_$1_$2.json 2> $O/r2_s23_$1_$2.err; echo "$1 $2 exit $?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_s23_$1_$2.json"))
    print("$1","$2","qps",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(d["roofline"]["frac"],3),"kern_ms",round(d["roofline"]["avg_launch_ms"],3),"recall",d.get("recall_at_10"),"parity",d["parity"]["queries_with_identical_ids"],d["parity"]["n_dist_equal_on_identical_walks"],"build_s",round(d["build"]["seconds"],2))
except Exception as e:
    print("$1","$2","failed",e)
PY
}
run nextpf E
run default E
run nextpf C
run default C
cp $V/libvecb200_default.so pgvector_b200/libvecb200.so
