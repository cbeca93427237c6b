set -x
PB_LIB=build/exp/libpb_bounds.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stress or locate or overlap or allpairs or assemble or strip or redo" > gpurun_out/r02g_bounds_check.log 2>&1; tail -3 gpurun_out/r02g_bounds_check.log
python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; tail -c 600 gpurun_out/r02g_bench_n1.err
python bench.py --impl reference > gpurun_out/r02g_bench_reference_arm.json 2> gpurun_out/r02g_bench_reference_arm.err
PB_TRACE=1 python tools/profile_step.py 100000 2 > gpurun_out/r02g_pb_trace_100k.log 2>&1
python tools/allpairs_bench.py > gpurun_out/r02g_allpairs_config5_n1.json 2> gpurun_out/r02g_allpairs_config5_n1.log; tail -4 gpurun_out/r02g_allpairs_config5_n1.log | cut -c1-400
python tools/config4_bench.py > gpurun_out/r02g_config4_n1.json 2> gpurun_out/r02g_config4_n1.log; tail -4 gpurun_out/r02g_config4_n1.log | cut -c1-300
