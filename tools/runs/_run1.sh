set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gputest0.log 2>&1; tail -3 gpurun_out/r02b_gputest0.log
for v in base notb nost; do
  echo "== $v"; PB_LIB=build/exp/libpb_$v.so python tools/profile_step.py 100000 3 2>&1 | grep -a "^step 2" | sed -e 's/.*wall, align/align/' | cut -c1-200
done > gpurun_out/r02b_ab_split.log 2>&1
cat gpurun_out/r02b_ab_split.log
python tools/probe_bulk.py > gpurun_out/r02b_probe_peak.log 2>&1; cat gpurun_out/r02b_probe_peak.log
