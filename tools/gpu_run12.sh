mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py -x -q -m gpu -k "search or neighbor or config3 or lockstep" > gpurun_out/r2u_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2u_tests.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-extra --no-cpu-baseline --workload config3 > gpurun_out/r2u_bench_c3.json 2> gpurun_out/r2u_bench_c3.err; echo "rc=$?" >> gpurun_out/r2u_bench_c3.err
tail -n 2 gpurun_out/r2u_tests.log gpurun_out/r2u_bench_c3.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2u_bench_c3.json')); print('c3 value %.3e ms %.3f'%(b['value'],b['ms_per_step']), b['roofline']['kernel_ms_per_episode'], b['roofline']['frac'])
PY
