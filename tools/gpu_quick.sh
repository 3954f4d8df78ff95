timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_umi_assign.py -x -q -m gpu > gpurun_out/q_tests.log 2>&1; tail -4 gpurun_out/q_tests.log
