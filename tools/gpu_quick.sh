timeout 900 python -m pytest tests/test_umi_assign.py tests/test_ref_vectors.py tests/test_abi.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/q_tests.log 2>&1; tail -3 gpurun_out/q_tests.log
timeout 600 python tools/perf_assign.py 4000000 4 2000 7 | cut -c1-400
