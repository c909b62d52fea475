mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_stress.py tests/test_gpu_policy.py -x -q -m gpu > gpurun_out/r2w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2w_tests.log
for i in 1 2; do timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2w_bench$i.json 2> gpurun_out/r2w_bench$i.err; done
tail -n 2 gpurun_out/r2w_tests.log
python - <<'PY'
import json
for i in (1,2):
    b=json.load(open('gpurun_out/r2w_bench%d.json'%i)); print(i, 'value %.4e ms %.3f kernel_us %.1f fresh %.3e traced %.3e'%(b['value'],b['ms_per_step'],b['roofline']['avg_launch_us'],b['value_fresh_streams']['value'],b['value_traced']['value']))
PY
