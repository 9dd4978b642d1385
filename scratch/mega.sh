set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_pytest_gpu_26.log; cat gpurun_out/r2_pytest_gpu_26.log
for sp in 0 1 0 1; do UB200_STEP_PLAN=$sp timeout 400 python bench.py --no-cpu-baseline --no-gpu-reference 2>/dev/null | tail -1 > gpurun_out/r2_bench_call26_plan${sp}_$RANDOM.json; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench_call26_plan*.json")):
    d=json.load(open(f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["config"].get("step_plan"), d["gpu_launches"])
PY
timeout 120 python benchmarks/kernel_bench.py gemv 2>&1 | grep gemv_nf4 | tee gpurun_out/r2_gemv_26.log
timeout 400 python benchmarks/ref_triton_bench.py --ops 2>&1 | grep "^{" | cut -c1-260 | tee gpurun_out/r2_ref_triton_ops_26.log
UB200_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_26.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-gpu-reference > /dev/null 2>&1
wc -l gpurun_out/r2_launches_26.csv
for t in attn_fwd attn_bwd; do timeout 300 ncu --set full --clock-control none -k regex:attn_ -c 2 --csv --page raw --log-file gpurun_out/r2_ncu_${t}_26.csv python benchmarks/ncu_targets.py $t > /dev/null 2>&1; done
ls -la gpurun_out/*_26*
