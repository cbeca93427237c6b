set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02z_bench_n8.json 2> gpurun_out/r02z_bench_n8.err
python -c "
import json
d=json.loads(open('gpurun_out/r02z_bench_n8.json').read().strip().splitlines()[-1])
print('N',d['n_gpus'],'value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'parity',d.get('parity_vs_reference_cpu'))
print('dev walls',d['device_leg']['call_wall_ms'])
print('e2e walls',d['e2e']['call_wall_ms'])
print('K3',d['roofline']['kernel_ms_per_step'])
"
