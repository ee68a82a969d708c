# compute-sanitizer memcheck over smoke() (tiny Segtran3d hot path: fused attention kernel, CTA row kernels, GEMMs, head)
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_smoke.log 2>&1
echo "rc=$?"; tail -6 gpurun_out/r2_sanitizer_smoke.log
