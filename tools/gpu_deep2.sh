timeout 600 python tools/perf_assign.py 2000000 60 2000 3 > gpurun_out/d2_perf_assign_m60.log 2>&1
timeout 600 python tools/perf_assign.py 2000000 300 2000 3 > gpurun_out/d2_perf_assign_m300.log 2>&1
timeout 900 python bench.py --workload umi5kx2k > gpurun_out/d2_bench_umi.json 2> gpurun_out/d2_bench_umi.err
tail -2 gpurun_out/d2_perf_assign_m60.log gpurun_out/d2_perf_assign_m300.log; cat gpurun_out/d2_bench_umi.json | cut -c1-3000; tail -5 gpurun_out/d2_bench_umi.err
