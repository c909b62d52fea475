mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2x_ref.json 2> gpurun_out/r2x_ref.err
for i in 1 2 3; do timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2x_bench$i.json 2> gpurun_out/r2x_bench$i.err; done
python - <<'PY'
import json
for i in (1,2,3):
    b=json.load(open('gpurun_out/r2x_bench%d.json'%i)); print(i, 'value %.4e ms %.3f host %.3f ovh %.3f kernel_us %.1f e2e %.4e'%(b['value'],b['ms_per_step'],b['host_ms_per_step'],b['launch_overhead_ms_per_step'],b['roofline']['avg_launch_us'],b['e2e']['value']))
PY
