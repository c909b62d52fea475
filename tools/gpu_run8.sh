mkdir -p gpurun_out
VDS_NQ=1 timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2j_nq.log 2>&1; echo "nq rc=$?" >> gpurun_out/r2j_nq.log
VDS_NQ=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2j_bench_nq.json 2> gpurun_out/r2j_bench_nq.err; echo "rc=$?" >> gpurun_out/r2j_bench_nq.err
tail -n 3 gpurun_out/r2j_nq.log gpurun_out/r2j_bench_nq.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2j_bench_nq.json')); print('value %.3e ms %.3f fresh %.3e traced %.3e frac %.3f'%(b['value'],b['ms_per_step'],b['value_fresh_streams']['value'],b['value_traced']['value'],b['roofline']['frac']), b['roofline']['kernel'])
PY
VDS_NQ=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_nq -c 1 -o gpurun_out/prof_r2_nq4 -f python profiles/prof_driver.py config2 > gpurun_out/prof_r2_nq4.log 2>&1
tail -n 2 gpurun_out/prof_r2_nq4.log
