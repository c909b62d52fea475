mkdir -p gpurun_out
VDS_TMA=1 timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_policy.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2e_tma_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_tma_tests.log
VDS_TMA=0 python profiles/tma_experiment.py > gpurun_out/r2e_tma0.txt 2>&1
VDS_TMA=1 python profiles/tma_experiment.py > gpurun_out/r2e_tma1.txt 2>&1
tail -n 3 gpurun_out/r2e_tma_tests.log; tail -n 1 gpurun_out/r2e_tma0.txt gpurun_out/r2e_tma1.txt
