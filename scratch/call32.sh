set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_pytest_gpu_32.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/r2_smoke_32.log
timeout 100 python benchmarks/glu_epilogue_bench.py 1 2>&1 | grep "^{" | tee gpurun_out/r2_glu_epilogue_bench_32_geglu.log
UB200_PROFILE_RANGE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_32.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-gpu-reference > /dev/null 2>&1
wc -l gpurun_out/r2_launches_32.csv
timeout 400 python bench.py 2>gpurun_out/bench32.err | tail -1 > gpurun_out/r2_bench_call32_default.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_call32_default.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], (d.get("gpu_reference") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
PY
timeout 120 compute-sanitizer --tool memcheck python benchmarks/sanitize_glu.py 2>&1 | tail -4 | tee gpurun_out/r2_sanitizer_glu_memcheck.log
