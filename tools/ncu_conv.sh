# full ncu capture of the LoG passes only (after a plain run of the same command)
python tools/prof_step.py C2 > gpurun_out/plain_prof.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:conv_strided|conv_y_kernel' -s 3 -c 3 -f -o gpurun_out/conv_prof python tools/prof_step.py C2 > gpurun_out/ncu_full.log 2>&1; echo ncu_full=$?
