timeout 900 python -m pytest tests/test_umi_assign.py tests/test_ref_vectors.py -x -q -m gpu -k "deep or myclust or assign" > gpurun_out/d1_tests.log 2>&1; echo "rc=$?" >> gpurun_out/d1_tests.log
tail -40 gpurun_out/d1_tests.log
