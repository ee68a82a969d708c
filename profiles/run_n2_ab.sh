#!/bin/bash
# cfg 4 on 2 GPUs: milestone-overlapped all-reduce inside the captured step vs one all-reduce after the step, two
# repetitions each (interleaved), strong scaling, and N=1 on the same box.  usage: bash profiles/run_n2_ab.sh [tag]
tag=${1:-r2_n2}
mkdir -p gpurun_out
export SEGTRAN_BENCH_WATCHDOG_S=200
run2() {  # name, extra flags
  timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port $((29500 + RANDOM % 2000)) bench.py --gpus 2 --steps 40 --warmup 5 $2 \
    > gpurun_out/${tag}_$1.json 2> gpurun_out/${tag}_$1.err || echo "$1 failed rc=$?"
}
run2 overlap_a ""
run2 nooverlap_a "--no-overlap"
run2 overlap_b ""
run2 nooverlap_b "--no-overlap"
run2 strong "--scaling strong"
timeout 260 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-fp32-equivalent \
  > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-36s N=%d %-6s %.3f ms/step  %.4g %s  checksums %s" % (f.split("/")[-1], d["n_gpus"], d["scaling"],
              d["ms_per_step"], d["value"], d["unit"], d.get("dp_param_checksums_agree")))
    except Exception as e:
        print(f, "NO LINE", e)
PY
