mkdir -p gpurun_out
VDS_NQ=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_nq -c 1 -o gpurun_out/prof_r2_nq5 -f python profiles/prof_driver.py config2 > gpurun_out/prof_r2_nq5.log 2>&1
tail -3 gpurun_out/prof_r2_nq5.log
