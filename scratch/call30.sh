set -x
timeout 300 python benchmarks/glu_epilogue_bench.py 2>&1 | grep "^{" | tee gpurun_out/r2_glu_epilogue_bench_30.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm2_kernel --launch-skip 1 -c 1 -f -o gpurun_out/r2_ncu_gemm_glu_bwd_30 python benchmarks/ncu_glu_target.py bwd > gpurun_out/ncu30.log 2>&1
tail -3 gpurun_out/ncu30.log
ls -la gpurun_out/*_30*
