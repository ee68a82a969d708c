#!/bin/bash
# ncu evidence of a round (run under gpurun on ONE GPU; numbers printed by a run under ncu are never bench values):
#   1. launch list of the default bench command (graph mode): per-kernel share of the step
#   2. DRAM bytes of every sx_gemm launch of one eager step (roofline.traffic)
#   3. --set full capture of the fused attention kernel and of the dominant GEMM
tag=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/${tag}_launches_graph.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/${tag}_launches_graph.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sx_gemm_kernel \
    --csv --log-file gpurun_out/${tag}_gemm_dram.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline \
    --no-eager-baseline > gpurun_out/${tag}_gemm_dram.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sx_attn_probs -s 3 -c 1 -o gpurun_out/${tag}_attn_full \
    python profiles/run_attn_kernel.py cfg4 2 > gpurun_out/${tag}_attn_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sx_gemm_kernel -s 40 -c 12 -o gpurun_out/${tag}_gemm_full \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline > gpurun_out/${tag}_gemm_full.log 2>&1
ls -la gpurun_out/${tag}_*
