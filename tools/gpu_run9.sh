mkdir -p gpurun_out
VDS_NQ=1 timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_stress.py -x -q -m gpu > gpurun_out/r2p_nq.log 2>&1; echo "nq rc=$?" >> gpurun_out/r2p_nq.log
VDS_NQ=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2p_bench_nq.json 2> gpurun_out/r2p_bench_nq.err; echo "rc=$?" >> gpurun_out/r2p_bench_nq.err
tail -n 3 gpurun_out/r2p_nq.log gpurun_out/r2p_bench_nq.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2p_bench_nq.json')); print('value %.3e ms %.3f fresh %.3e traced %.3e frac %.3f'%(b['value'],b['ms_per_step'],b['value_fresh_streams']['value'],b['value_traced']['value'],b['roofline']['frac']), b['roofline']['kernel'])
PY
VDS_NQ=1 timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_boundary.py -x -q -m gpu > gpurun_out/r2p_nq2.log 2>&1; echo "nq2 rc=$?" >> gpurun_out/r2p_nq2.log; tail -n 2 gpurun_out/r2p_nq2.log
VDS_NQ=1 timeout 600 python bench.py --steps 5 --warmup 2 --no-extra --no-cpu-baseline --workload config5 > gpurun_out/r2p_bench_c5.json 2> gpurun_out/r2p_bench_c5.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2p_bench_c5.json')); print('c5 value %.3e ms %.3f frac %.3f'%(b['value'],b['ms_per_step'],b['roofline']['frac']), b['roofline']['kernel'])
PY
