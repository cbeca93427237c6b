N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02u_bench_n$N.json 2> gpurun_out/r02u_bench_n$N.err
tail -c 1500 gpurun_out/r02u_bench_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02u_bench_n$N.json').read().strip().splitlines()[-1])
print("N",d['n_gpus'],"value",d['value'],d['ms_per_step'],"e2e",d['e2e']['value'],d['e2e']['ms_per_step'],"parity",d.get('parity_vs_reference_cpu'))
P
