timeout 600 ncu --set full --clock-control none --import-source on -k regex:_cta -s 8 -c 8 -o gpurun_out/r2_rows_cta -f python profiles/run_row_cta.py > gpurun_out/r2_rows_cta.log 2>&1
tail -3 gpurun_out/r2_rows_cta.log
