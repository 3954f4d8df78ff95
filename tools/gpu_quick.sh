timeout 600 python -m pytest tests/test_umi_assign.py tests/test_ref_vectors.py -x -q -m gpu -k "deep or myclust" > gpurun_out/q_tests.log 2>&1; tail -2 gpurun_out/q_tests.log
timeout 600 python tools/perf_deep.py 1000 4097 20000 2>&1 | grep -v "dbg" | tee gpurun_out/q_perf_deep.log
