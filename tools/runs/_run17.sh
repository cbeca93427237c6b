set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "align or config3 or aligner or thread_kernel" > gpurun_out/r03c_aligntests.log 2>&1; tail -5 gpurun_out/r03c_aligntests.log
timeout 600 python tools/dp_sweep.py 4096 --check 4 > gpurun_out/r03c_dp_sweep_unit.log 2>&1; tail -8 gpurun_out/r03c_dp_sweep_unit.log
for pt in "2000 64" "2000 128" "10000 128"; do
  for m in 0 100000; do echo "== $pt thread_min $m"; PB_THREAD_MIN_ITEMS=$m python tools/sweep_point.py $pt 8192 3 2>&1 | tail -1; done
done > gpurun_out/r03c_crossover.log 2>&1
cat gpurun_out/r03c_crossover.log
