# quick GPU validation of bench.py's fp32_equivalent figure (tf32x3 step replayed as a CUDA graph): cfg 4 and cfg 3
mkdir -p gpurun_out
for c in 4 3; do
  timeout 200 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2g_cfg$c.json 2> gpurun_out/r2g_cfg$c.err
  python - $c <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2g_cfg%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("cfg", sys.argv[1], "%.3f ms/step" % d["ms_per_step"], d.get("fp32_equivalent"))
PY
done
