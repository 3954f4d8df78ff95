set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c1_tests.log
timeout 900 python tools/make_profile.py gpurun_out/r2_kernel_profile.json > gpurun_out/c1_profile.log 2>&1
cp gpurun_out/r2_kernel_profile.json profiles/r2_kernel_profile.json
timeout 600 python bench.py > gpurun_out/c1_bench_bc3m.json 2> gpurun_out/c1_bench_bc3m.err
timeout 600 python tools/perf_assign.py 4000000 4 2000 5 > gpurun_out/c1_perf_assign.log 2>&1
timeout 600 python tools/bench_umi_cluster.py > gpurun_out/c1_bench_umi_cluster.json 2> gpurun_out/c1_bench_umi_cluster.err
timeout 600 python tools/bench_guided.py --flavour umi --ed 2 > gpurun_out/c1_bench_guided_umi.json 2> gpurun_out/c1_bench_guided_umi.err
tail -3 gpurun_out/c1_tests.log; cat gpurun_out/c1_bench_bc3m.json; tail -5 gpurun_out/c1_perf_assign.log
