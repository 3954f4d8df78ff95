timeout 900 python -m pytest tests/test_umi_assign.py tests/test_ref_vectors.py -x -q -m gpu -k "deep or myclust or assign" > gpurun_out/deep_tests.log 2>&1; tail -3 gpurun_out/deep_tests.log
timeout 600 python tools/perf_deep.py 301 1000 4001 20000 2>&1 | tee gpurun_out/deep_perf_deep.log
timeout 600 python tools/perf_assign.py 4000000 4 2000 5 2>&1 | tail -1 | cut -c1-400
