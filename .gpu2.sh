python profiles/profile_driver.py 2 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_unbinned_mma -c 2 -f -o gpurun_out/prof_r1i python profiles/profile_driver.py 1 > gpurun_out/ncu_r1i.log 2>&1
tail -3 gpurun_out/ncu_r1i.log
