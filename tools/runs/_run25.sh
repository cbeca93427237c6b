for rep in 1 2; do
for mode in 0 1; do
PB_SPIN_WAIT=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$mode bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r03k_n2_spin${mode}_$rep.json 2> gpurun_out/r03k_n2_spin${mode}_$rep.err
python -c "
import json
d=json.loads(open('gpurun_out/r03k_n2_spin${mode}_$rep.json').read().strip().splitlines()[-1])
print('spin',$mode,'rep',$rep,'value',round(d['value']),round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']), 'max wall dev',max(d['device_leg']['call_wall_ms']),'max wall e2e',max(d['e2e']['call_wall_ms'][1:]))
"
done
done
