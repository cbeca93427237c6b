python tools/allpairs_bench.py > gpurun_out/r02m_allpairs_config5_n1.log 2>&1; tail -3 gpurun_out/r02m_allpairs_config5_n1.log | cut -c1-300
