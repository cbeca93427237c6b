L=gpurun_out/r03h_thread_knobs.log
: > $L
for v in base u4 u16 u32 col30; do
  for pt in "2000 32" "2000 64" "5000 128"; do
    echo "== $v $pt" >> $L
    if [ $v = base ]; then python tools/sweep_point.py $pt 4096 3 2>&1 | tail -1 >> $L; else PB_LIB=build/exp/libpb_$v.so python tools/sweep_point.py $pt 4096 3 2>&1 | tail -1 >> $L; fi
  done
done
cat $L | paste - -
