python tools/profile_step.py 20000 2 4200 5000 > gpurun_out/r02f_plain_s2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:align_locate_nb_kernel -s 1 -c 1 -o gpurun_out/r02f_nb_s2 python tools/profile_step.py 20000 2 4200 5000 > gpurun_out/r02f_ncu_s2.log 2>&1
tail -3 gpurun_out/r02f_plain_s2.log | cut -c1-300; tail -5 gpurun_out/r02f_ncu_s2.log | cut -c1-200
