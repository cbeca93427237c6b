# thin-warp knob of the one-alignment-per-thread kernels: sweep points at 4096 pairs, lanes per warp 32..1; forward / traceback split
set -x
L=gpurun_out/r02v_thread_tl.log
: > $L
for pt in "1000 32" "5000 32" "2000 64" "2000 128" "10000 128"; do
  for tl in 32 16 8 4 2 1; do
    echo "== point $pt min_tl $tl" >> $L
    PB_THREAD_MIN_TL=$tl PB_THREAD_WARPS_PER_SM=1000 python tools/sweep_point.py $pt 4096 3 2>&1 | tail -1 >> $L
  done
done
for tl in 32 4; do
  echo "== NOTB point 2000 32 min_tl $tl" >> $L
  PB_LIB=build/exp/libpb_notb.so PB_THREAD_MIN_TL=$tl PB_THREAD_WARPS_PER_SM=1000 python tools/sweep_point.py 2000 32 4096 3 2>&1 | tail -1 >> $L
done
echo "== 65536 pairs 2000 32 default" >> $L
python tools/sweep_point.py 2000 32 65536 3 2>&1 | tail -1 >> $L
cat $L
PB_THREAD_MIN_TL=4 PB_THREAD_WARPS_PER_SM=1000 timeout 300 ncu --set full --import-source on --clock-control none -k regex:align_pairs_thread -c 1 -s 1 -o gpurun_out/r02v_thread_w3 -f python tools/sweep_point.py 2000 32 4096 2 > gpurun_out/r02v_ncu.log 2>&1
tail -3 gpurun_out/r02v_ncu.log
