set -x
timeout 900 python tools/config4_bench.py --steps 3 --sample 32 > gpurun_out/r03d_config4_n1.log 2>&1; tail -6 gpurun_out/r03d_config4_n1.log | cut -c1-600
timeout 900 python tools/config4_bench.py --steps 2 --sample 0 --serial > gpurun_out/r03d_config4_n1_serial.log 2>&1; tail -4 gpurun_out/r03d_config4_n1_serial.log | cut -c1-400
