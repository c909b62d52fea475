mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_boundary.py -x -q -m gpu > gpurun_out/r2f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_tests.log
python profiles/graph_tick_experiment.py > gpurun_out/r2f_graph_tick.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "rc=$?" >> gpurun_out/r2f_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-graph > gpurun_out/r2f_launch_bench.log 2>&1
tail -n 3 gpurun_out/r2f_tests.log gpurun_out/r2f_bench.err; tail -n 1 gpurun_out/r2f_graph_tick.txt
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2f_bench.json')); print('value %.3e ms %.3f fresh %.3e traced %.3e frac %.3f'%(b['value'],b['ms_per_step'],b['value_fresh_streams']['value'],b['value_traced']['value'],b['roofline']['frac']))
PY
