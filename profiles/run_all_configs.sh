#!/bin/bash
# One bench line per BASELINE.json config on one GPU (device-timed value, e2e, roofline, PyTorch-CUDA eager baseline, CPU
# baseline with a 40 s budget per config so that the whole script stays within ~10 GPU-minutes).
# usage: bash profiles/run_all_configs.sh [tag]
tag=${1:-r2}
mkdir -p gpurun_out
for c in 4 1 2 3 5; do
  SEGTRAN_CPU_BUDGET_S=40 timeout 500 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/${tag}_cfg${c}.json 2> gpurun_out/${tag}_cfg${c}.err || echo "cfg $c failed rc=$?"
done
for c in 4 5; do
  timeout 400 python bench.py --config $c --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/${tag}_cfg${c}_bf16.json 2> gpurun_out/${tag}_cfg${c}_bf16.err || echo "cfg $c bf16 failed rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_cfg*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "NO LINE", e); continue
    eb = d.get("cuda_eager_baseline", {})
    print("%-32s %-6s %8.3f ms  %.3e %s  e2e %.1f ms  roofline %.3f (%s)  eager-fp32 %s x%s" % (
        f.split("/")[-1], d["dtype"], d["ms_per_step"], d["value"], d["unit"], d["e2e"]["ms_per_step"], d["roofline"]["frac"],
        d["roofline"]["kernel"][:12], ("%.1f ms" % eb["fp32"]["ms_per_step"]) if "fp32" in eb else "-",
        ("%.1f" % eb["speedup_vs_fp32"]) if "speedup_vs_fp32" in eb else "-"))
PY
