#!/bin/bash
# round-end validation on one GPU: GPU test suite, smoke(), the default bench line, and the ncu launch list of the same
# bench command (graph mode; numbers printed under ncu are never bench values)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 500 python bench.py > gpurun_out/r2_default.json 2> gpurun_out/r2_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_default.json").read().strip().splitlines()[-1])
print("default bench: %.3f ms/step  %.4g %s  e2e %.1f ms  roofline %.3f  launches %d  cpu %s  clocks %s" % (
    d["ms_per_step"], d["value"], d["unit"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"],
    d.get("cpu_baseline"), d["clocks"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2_launches_final.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-fp32-equivalent > gpurun_out/r2_launches_final.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_final.csv | head -12
