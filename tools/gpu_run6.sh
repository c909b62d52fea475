mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "rc=$?" >> gpurun_out/r2g_bench.err
tail -n 3 gpurun_out/r2g_tests.log gpurun_out/r2g_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2g_bench.json')); print('value %.3e ms %.3f e2e %.3e fresh %.3e traced %.3e frac %.3f parity %s'%(b['value'],b['ms_per_step'],b['e2e']['value'],b['value_fresh_streams']['value'],b['value_traced']['value'],b['roofline']['frac'],b['parity_check_vs_oracle']))
for k,v in b['extra_workloads'].items(): print(k, {kk:v.get(kk) for kk in ('value','ms_per_step','roofline_frac','error')})
PY
