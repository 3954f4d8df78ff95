timeout 600 python tools/perf_deep.py 1000 20000 2>&1 | tee gpurun_out/d4_perf_deep.log
