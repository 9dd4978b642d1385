set -x
timeout 400 python -m pytest tests/test_gpu_gemm_lora.py -x -q -k "glu" 2>&1 | tail -12 | tee gpurun_out/r2_pytest_glu_31.log
timeout 300 python benchmarks/glu_epilogue_bench.py 2>&1 | grep "^{" | tee gpurun_out/r2_glu_epilogue_bench_31.log
i=0
for m in 1 0 1 0; do i=$((i+1)); UB200_FUSED_GLU=$m timeout 400 python bench.py --no-cpu-baseline --no-gpu-reference 2>gpurun_out/bench31_$i.err | tail -1 > gpurun_out/r2_bench_call31_glu${m}_$i.json; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench_call31_glu*.json")):
    try:
        d=json.load(open(f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["gpu_launches"], d["loss"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/bench31_1.err
