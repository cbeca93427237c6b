set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r03j_bench_n4.json 2> gpurun_out/r03j_bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r03j_bench_n2.json 2> gpurun_out/r03j_bench_n2.err
python -c "
import json
for n in (4,2):
    d=json.loads(open(f'gpurun_out/r03j_bench_n{n}.json').read().strip().splitlines()[-1])
    print('N',d['n_gpus'],'value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'parity',d.get('parity_vs_reference_cpu'))
"
