set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02y_gputest.log 2>&1; tail -3 gpurun_out/r02y_gputest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02y_bench_n1.json 2> gpurun_out/r02y_bench_n1.err; tail -c 600 gpurun_out/r02y_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02y_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'cpu',d.get('cpu_baseline'),'parity',d.get('parity_vs_reference_cpu'))
print('dev walls',d['device_leg']['call_wall_ms'])
"
timeout 600 python tools/dp_sweep.py 4096 --check 4 > gpurun_out/r02y_dp_sweep_unit.log 2>&1; tail -8 gpurun_out/r02y_dp_sweep_unit.log
