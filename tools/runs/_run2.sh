python tools/profile_step.py 100000 3 2>&1 | grep -a "^step 2\|rror" | sed -e 's/.*wall, align/align/' | cut -c1-260
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gputest2.log 2>&1; tail -15 gpurun_out/r02c_gputest2.log
python bench.py --steps 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02h_bench.json').read().strip().splitlines()[-1])
print("value",d['value'],d['ms_per_step'],"e2e",d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['stage_ms'])
P
python tools/allpairs_bench.py --targets 2 2>&1 | grep -a "stage ms\|step 1\|bit-exact" | tail -4 | cut -c1-300
