#!/bin/bash
# round-2 8-GPU session (run under gpurun --gpus 8): config D at its stated size (10M x 1536, lists 4096, 8 x B200), then config B if time is left
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 210 $RUN --master-port 29513 bench.py --config D --gpus $N > gpurun_out/r2_bench_D_n$N.json 2> gpurun_out/r2_bench_D_n$N.err; echo "D exit $?"; tail -2 gpurun_out/r2_bench_D_n$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_D_n$N.json"))
    print("D n=$N rows/s", round(d["value"]), "ms/step", round(d["ms_per_step"],1), d["phases_s"], "recall", d["recall_at_10"], "roof", d["roofline"] and {k:d["roofline"].get(k) for k in ("achieved","frac","useful_tflops")})
except Exception as e:
    print("D n=$N failed", e)
PY
