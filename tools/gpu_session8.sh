#!/bin/bash
# round-2 session 8 (1 GPU): visited-table size x persisting L2 window, shared-memory visited cache, merge variant
mkdir -p gpurun_out
run() {  # name defs l2
  touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu
  VB_NVCC_DEFS="$2" python -m pgvector_b200.build > /dev/null 2> gpurun_out/r2_s8_build_$1.err
  for l2 in $3; do
    VB_HNSW_L2=$l2 timeout 600 python bench.py --config C --no-cpu > gpurun_out/r2_s8_$1_l2${l2}_C.json 2> gpurun_out/r2_s8_$1_l2${l2}_C.err
    VB_HNSW_L2=$l2 timeout 900 python bench.py --config E --no-cpu > gpurun_out/r2_s8_$1_l2${l2}_E.json 2> gpurun_out/r2_s8_$1_l2${l2}_E.err
  done
}
run vis0 "VB_AB_VIS=0" "0 1"
run vis1 "VB_AB_VIS=1" "0 1"
run vis0_vcache "VB_AB_VIS=0 VB_AB_VCACHE=1" "1"
run vis0_twobuf "VB_AB_VIS=0 VB_AB_INPLACE=0" "1"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_s8_*_[CE].json")):
    try:
        d=json.load(open(f))
        print(f.split("r2_s8_")[1][:-5],"qps",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms/step",round(d["ms_per_step"],3),"frac",round(d["roofline"]["frac"],3),"recall",d.get("recall_at_10"))
    except Exception as e:
        print(f,"failed",e)
PY
touch pgvector_b200/csrc/vb_hnsw.cu pgvector_b200/csrc/vb_hnsw_iter.cu pgvector_b200/csrc/vb_hnsw_build.cu; python -m pgvector_b200.build > /dev/null 2>&1
timeout 600 python tools/diag_build_phases.py > gpurun_out/r2_diag_build4.json 2> gpurun_out/r2_diag_build4.err; cat gpurun_out/r2_diag_build4.json
timeout 900 python bench.py --config D --no-cpu > gpurun_out/r2_bench_D8.json 2> gpurun_out/r2_bench_D8.err; cut -c1-300 gpurun_out/r2_bench_D8.json
