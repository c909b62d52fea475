mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests/test_gpu_boundary.py -x -q -m gpu > gpurun_out/r2a_boundary.log 2>&1; echo "boundary rc=$?" >> gpurun_out/r2a_boundary.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_boundary.py > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?" >> gpurun_out/r2a_ref.err
tail -3 gpurun_out/r2a_boundary.log gpurun_out/r2a_tests.log gpurun_out/r2a_bench.err gpurun_out/r2a_ref.err
