N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/config4_bench.py > gpurun_out/r02n_config4_n$N.log 2>&1; tail -4 gpurun_out/r02n_config4_n$N.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/allpairs_bench.py > gpurun_out/r02n_allpairs_config5_n$N.log 2>&1; tail -3 gpurun_out/r02n_allpairs_config5_n$N.log | cut -c1-400
