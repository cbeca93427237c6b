set -x
L=gpurun_out/r02x_thread.log
: > $L
for pt in "1000 32" "5000 32"; do
  python tools/sweep_point.py $pt 4096 3 2>&1 | tail -1 >> $L
  PB_LIB=build/exp/libpb_notb.so python tools/sweep_point.py $pt 4096 3 2>&1 | tail -1 >> $L
done
cat $L
timeout 300 ncu --set full --import-source on --clock-control none -k regex:align_pairs_thread -c 1 -s 1 -o gpurun_out/r02x_thread_w3 -f python tools/sweep_point.py 5000 32 4096 2 > gpurun_out/r02x_ncu.log 2>&1
tail -3 gpurun_out/r02x_ncu.log
