set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02l_gputest.log 2>&1; tail -4 gpurun_out/r02l_gputest.log
python tools/profile_step.py 100000 2 > gpurun_out/r02l_k3_traffic_plain.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:align_locate -s 20 -c 20 --csv --log-file gpurun_out/r02l_k3_traffic.csv python tools/profile_step.py 100000 2 > gpurun_out/r02l_k3_traffic_run.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02l_bench_plain.json 2> gpurun_out/r02l_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02l_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02l_ncu_bench.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02l_bench_reference_arm.json 2> gpurun_out/r02l_bench_reference_arm.err
tail -c 300 gpurun_out/r02l_bench_n1.json
