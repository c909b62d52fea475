mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2d_nq.log 2>&1; echo "nq rc=$?" >> gpurun_out/r2d_nq.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2d_bench_nq.json 2> gpurun_out/r2d_bench_nq.err; echo "rc=$?" >> gpurun_out/r2d_bench_nq.err
timeout 600 python bench.py --steps 5 --warmup 2 --no-extra --no-cpu-baseline --workload config5 > gpurun_out/r2d_bench_c5.json 2> gpurun_out/r2d_bench_c5.err; echo "rc=$?" >> gpurun_out/r2d_bench_c5.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_nq -c 1 -o gpurun_out/prof_r2_nq3 -f python profiles/prof_driver.py config2 > gpurun_out/prof_r2_nq3.log 2>&1
tail -n 4 gpurun_out/r2d_nq.log gpurun_out/r2d_bench_nq.err gpurun_out/r2d_bench_c5.err gpurun_out/prof_r2_nq3.log
python - <<'PY'
import json
for f in ('gpurun_out/r2d_bench_nq.json','gpurun_out/r2d_bench_c5.json'):
    try:
        b=json.load(open(f)); print(f, 'value %.3e ms %.3f fresh %.3e traced %.3e frac %.3f'%(b['value'],b['ms_per_step'],b['value_fresh_streams']['value'],b['value_traced']['value'],b['roofline']['frac']), b['returns_sha256']['sha256'][:12])
    except Exception as e: print(f, e)
PY
