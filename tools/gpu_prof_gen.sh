mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gen_prepare -c 1 -o gpurun_out/prof_r2_genprep -f python profiles/prof_gen_driver.py > gpurun_out/prof_r2_genprep.log 2>&1
tail -3 gpurun_out/prof_r2_genprep.log
