#!/bin/bash
# cfg 4 on 4 GPUs: in-graph milestone exchange vs one all-reduce after the step (one run each)
mkdir -p gpurun_out
export SEGTRAN_BENCH_WATCHDOG_S=150
for mode in "" "--no-overlap"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
    --master-port $((29500 + RANDOM % 2000)) bench.py --gpus 4 --steps 30 --warmup 5 $mode \
    > "gpurun_out/r2_n4${mode}.json" 2> "gpurun_out/r2_n4${mode}.err" || echo "mode '$mode' failed rc=$?"
  python - "gpurun_out/r2_n4${mode}.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], "%.3f ms/step" % d["ms_per_step"], "%.4g" % d["value"], d.get("dp_param_checksums_agree"))
except Exception as e:
    print(sys.argv[1], "NO LINE", e)
PY
done
