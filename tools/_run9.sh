python bench.py --steps 20 --warmup 5 > gpurun_out/r02r_bench_n1.json 2> gpurun_out/r02r_bench_n1.err; tail -c 500 gpurun_out/r02r_bench_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02r_bench_n1.json').read().strip().splitlines()[-1])
print("value",d['value'],d['ms_per_step'],"serial",d['device_leg']['serial']['ms_per_step'],"e2e",d['e2e']['value'],d['e2e']['ms_per_step'],"K3",d['roofline']['kernel_ms'],d['roofline']['traffic'],d['gpu_launches'])
print(d['device_leg']['call_wall_ms'])
print(d['parity_vs_reference_cpu'], d['cpu_baseline'])
P
