mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2z_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_tests.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err; echo "rc=$?" >> gpurun_out/r2z_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "rc=$?" >> gpurun_out/r2z_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-graph > gpurun_out/r2z_launch_bench.log 2>&1
tail -n 3 gpurun_out/r2z_smoke.log gpurun_out/r2z_tests.log gpurun_out/r2z_ref.err gpurun_out/r2z_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2z_bench.json')); print('value %.3e ms %.3f e2e %.3e fresh %.3e traced %.3e frac %.3f parity %s'%(b['value'],b['ms_per_step'],b['e2e']['value'],b['value_fresh_streams']['value'],b['value_traced']['value'],b['roofline']['frac'],b['parity_check_vs_oracle']))
for k,v in b['extra_workloads'].items(): print(k, {kk:v.get(kk) for kk in ('value','ms_per_step','roofline_frac','error')})
r=json.load(open('gpurun_out/r2z_ref.json')); print('ref', r['value'], r.get('reference_python',{}).get('single_process'), r.get('reference_python',{}).get('all_cores'))
PY
