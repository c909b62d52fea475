mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2y_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_tests.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_rollout.py -q -m gpu -k "golden or search" >> gpurun_out/r2y_tests2.log 2>&1; done
tail -n 3 gpurun_out/r2y_tests.log; grep -E "passed|failed" gpurun_out/r2y_tests2.log
