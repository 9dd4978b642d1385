set -x
timeout 60 python -m pytest tests/test_gpu_kernels.py -x -q -k "cast_pad_multi" > gpurun_out/r2_pytest_castpad_33.log 2>&1; tail -3 gpurun_out/r2_pytest_castpad_33.log
UB200_PROFILE_RANGE=1 timeout 90 ncu --profile-from-start off -k regex:multi_kernel --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_multi_33.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-gpu-reference > /dev/null 2>&1
grep -c multi_kernel gpurun_out/r2_launches_multi_33.csv
timeout 60 python bench.py --no-cpu-baseline --no-gpu-reference 2>/dev/null | tail -1 > gpurun_out/r2_bench_call33.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_call33.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])"
timeout 150 python -m pytest tests/test_gpu_vs_reference.py -x -q > gpurun_out/r2_pytest_vs_reference_33.log 2>&1; tail -4 gpurun_out/r2_pytest_vs_reference_33.log
