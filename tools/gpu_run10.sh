mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2o_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "rc=$?" >> gpurun_out/r2o_bench.err
VDS_NQ=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_nq -c 1 -o gpurun_out/prof_r2_nq_c5 -f python profiles/prof_driver.py config5 296 > gpurun_out/prof_r2_nq_c5.log 2>&1
tail -n 3 gpurun_out/r2o_tests.log gpurun_out/r2o_bench.err gpurun_out/prof_r2_nq_c5.log
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2o_bench.json')); print('value %.3e ms %.3f e2e %.3e fresh %.3e traced %.3e frac %.3f'%(b['value'],b['ms_per_step'],b['e2e']['value'],b['value_fresh_streams']['value'],b['value_traced']['value'],b['roofline']['frac']), b['roofline']['kernel'])
PY
