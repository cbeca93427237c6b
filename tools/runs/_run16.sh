set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "align or config3 or aligner or thread_kernel" > gpurun_out/r03b_aligntests.log 2>&1; tail -5 gpurun_out/r03b_aligntests.log
timeout 600 python tools/dp_sweep.py 4096 --check 4 > gpurun_out/r03b_dp_sweep_unit.log 2>&1; tail -8 gpurun_out/r03b_dp_sweep_unit.log
