PB_HOST_TRACE=1 python tools/pipe_probe.py 100000 5 text 2>&1 | tail -40 | cut -c1-250
