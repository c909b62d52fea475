mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_nq -c 1 -o gpurun_out/prof_r2_nq1 -f python profiles/prof_driver.py config2 > gpurun_out/prof_r2_nq1.log 2>&1
tail -3 gpurun_out/prof_r2_nq1.log
