set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/config4_bench.py --steps 4 --sample 32 > gpurun_out/r03e_config4_n8.log 2>&1; grep -v "^\[rank\|NCCL\|^W" gpurun_out/r03e_config4_n8.log | tail -8 | cut -c1-500
