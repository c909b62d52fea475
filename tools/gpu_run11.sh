mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_synthetic.py -x -q -m gpu > gpurun_out/r2v_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "rc=$?" >> gpurun_out/r2v_bench.err
tail -n 2 gpurun_out/r2v_tests.log gpurun_out/r2v_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2v_bench.json')); print('value %.3e ms %.3f fresh %.3e traced %.3e'%(b['value'],b['ms_per_step'],b['value_fresh_streams']['value'],b['value_traced']['value']))
PY
