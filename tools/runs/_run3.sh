for v in main unr4 unr16 unr32; do
  lib=build/exp/libpb_$v.so; [ $v = main ] && lib=pacbioassembly_b200/libpacbio_b200.so
  echo "== $v"; PB_LIB=$lib python tools/profile_step.py 100000 3 2>&1 | grep -a "^step 2" | sed -e 's/.*wall, align/align/' | cut -c1-60
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
