timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
SEGTRAN_BENCH_VERBOSE=1 timeout 300 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2d_cfg4.json 2> gpurun_out/r2d_cfg4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2d_cfg4.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["roofline"]["frac"], d.get("fp32_equivalent"))
print(d["kernel_breakdown"]["sx_gemm"])
PY
grep "^GEMM" gpurun_out/r2d_cfg4.err | head -8
