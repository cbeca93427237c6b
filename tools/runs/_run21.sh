set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "weighted" > gpurun_out/r03g_wtests.log 2>&1; tail -5 gpurun_out/r03g_wtests.log
timeout 900 python tools/dp_sweep.py 4096 --weighted --check 3 > gpurun_out/r03g_dp_sweep_weighted.log 2>&1; tail -8 gpurun_out/r03g_dp_sweep_weighted.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r03g_bench_n1.json 2> gpurun_out/r03g_bench_n1.err; tail -c 400 gpurun_out/r03g_bench_n1.json
