timeout 600 python -m pytest tests/test_umi_assign.py tests/test_ref_vectors.py tests/test_multi_gpu.py -x -q -m gpu -k "deep or myclust or multi or assign" > gpurun_out/q_tests.log 2>&1; tail -2 gpurun_out/q_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-120
timeout 300 python tools/perf_deep.py 20000 2>&1 | tail -2 | cut -c1-250
