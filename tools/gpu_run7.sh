mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "rc=$?" >> gpurun_out/r2h_bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo "rc=$?" >> gpurun_out/r2h_bench_n1.err
timeout 300 python -m pytest tests/test_gpu_synthetic.py -x -q -m gpu > gpurun_out/r2h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_tests.log
tail -n 3 gpurun_out/r2h_bench_n2.err gpurun_out/r2h_bench_n1.err gpurun_out/r2h_tests.log
python - <<'PY'
import json
for f in ('gpurun_out/r2h_bench_n1.json','gpurun_out/r2h_bench_n2.json'):
    b=json.load(open(f)); print(f, 'N',b['n_gpus'],'value %.4e ms %.3f host_ms %.3f ovh %.3f e2e %.4e fresh %.3e kernel_us %.1f sha %s'%(b['value'],b['ms_per_step'],b['host_ms_per_step'],b['launch_overhead_ms_per_step'],b['e2e']['value'],b['value_fresh_streams']['value'],b['roofline']['avg_launch_us'],b['returns_sha256']['sha256'][:12]))
PY
