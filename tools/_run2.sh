python tools/profile_step.py 100000 3 2>&1 | grep -a "^step\|rror" | sed -e 's/.*wall, align/align/' | cut -c1-260 > gpurun_out/r02c_step.log 2>&1; cat gpurun_out/r02c_step.log
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gputest1.log 2>&1; tail -15 gpurun_out/r02c_gputest1.log
