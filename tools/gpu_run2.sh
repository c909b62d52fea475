mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_stress.py -x -q -m gpu > gpurun_out/r2b_nq.log 2>&1; echo "nq rc=$?" >> gpurun_out/r2b_nq.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_rollout.py --deselect tests/test_gpu_stress.py > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2b_bench_nq.json 2> gpurun_out/r2b_bench_nq.err; echo "rc=$?" >> gpurun_out/r2b_bench_nq.err
VDS_NO_NQ=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2b_bench_classic.json 2> gpurun_out/r2b_bench_classic.err; echo "rc=$?" >> gpurun_out/r2b_bench_classic.err
timeout 600 python bench.py --steps 5 --warmup 2 --no-extra --no-cpu-baseline --workload config5 > gpurun_out/r2b_bench_c5.json 2> gpurun_out/r2b_bench_c5.err; echo "rc=$?" >> gpurun_out/r2b_bench_c5.err
tail -n 4 gpurun_out/r2b_nq.log gpurun_out/r2b_tests.log gpurun_out/r2b_bench_nq.err gpurun_out/r2b_bench_classic.err gpurun_out/r2b_bench_c5.err
