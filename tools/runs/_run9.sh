python -m pytest tests -m gpu -x -q > gpurun_out/r02u_gputest.log 2>&1; tail -3 gpurun_out/r02u_gputest.log
