set -x
timeout 300 ncu --set full --import-source on --clock-control none -k regex:align_pairs_thread -c 1 -s 1 -o gpurun_out/r03i_thread_w3 -f python tools/sweep_point.py 5000 32 4096 2 > gpurun_out/r03i_ncu_thread.log 2>&1
tail -2 gpurun_out/r03i_ncu_thread.log
cat > /tmp/wpoint.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, workload
from pacbioassembly_b200 import Context
alen, band, npairs = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(5)
P = [workload.sweep_pair(1000 * band + alen, k, alen, band) for k in range(npairs)]
A, B, R = [x[0] for x in P], [x[1] for x in P], P[0][2]
WA = np.concatenate([rng.integers(1, 5, size=len(a)).astype(np.uint8) for a in A]); WB = np.concatenate([rng.integers(1, 5, size=len(b)).astype(np.uint8) for b in B])
a_len = np.array([len(x) for x in A], dtype=np.int32); b_len = np.array([len(x) for x in B], dtype=np.int32)
a_off = np.zeros(npairs, dtype=np.int64); np.cumsum(a_len[:-1], out=a_off[1:])
b_off = np.zeros(npairs, dtype=np.int64); np.cumsum(b_len[:-1], out=b_off[1:])
ctx = Context(0)
for rep in range(2):
    recs, _ = ctx.align_weighted_batch(b"".join(A), WA, a_off, a_len, b"".join(B), WB, b_off, b_len, R, 4.0, 26000, 6000)
    t = ctx.timings(); cells = int(recs["cells"].sum())
    print(f"weighted len {alen} band {band} pairs {npairs}: aligned {int((recs['ret'] >= 0).sum())}, K3 {t['align']:.3f} ms, {cells / t['align'] / 1e6:.0f} GCUPS")
ctx.close()
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:alignw_reg -c 1 -s 1 -o gpurun_out/r03i_alignw_cpl17 -f python /tmp/wpoint.py 2000 512 4096 > gpurun_out/r03i_ncu_w17.log 2>&1
tail -2 gpurun_out/r03i_ncu_w17.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:alignw_reg -c 1 -s 1 -o gpurun_out/r03i_alignw_cpl2 -f python /tmp/wpoint.py 2000 32 4096 > gpurun_out/r03i_ncu_w2.log 2>&1
tail -2 gpurun_out/r03i_ncu_w2.log
