set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "weighted or thread_kernel" > gpurun_out/r03a_wtests.log 2>&1; tail -5 gpurun_out/r03a_wtests.log
timeout 900 python tools/dp_sweep.py 4096 --weighted --check 2 > gpurun_out/r03a_dp_sweep_weighted.log 2>&1; tail -8 gpurun_out/r03a_dp_sweep_weighted.log
