set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "align or config3 or aligner" > gpurun_out/r02w_aligntests.log 2>&1; tail -5 gpurun_out/r02w_aligntests.log
L=gpurun_out/r02w_thread.log
: > $L
for pt in "1000 32" "5000 32" "2000 64" "2000 128" "10000 128"; do
  python tools/sweep_point.py $pt 4096 3 2>&1 | tail -1 >> $L
done
python tools/sweep_point.py 2000 32 65536 3 2>&1 | tail -1 >> $L
python tools/sweep_point.py 2000 128 65536 3 2>&1 | tail -1 >> $L
cat $L
