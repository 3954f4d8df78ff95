set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_smoke.log
timeout 1800 python tools/make_profile.py gpurun_out/r2_kernel_profile.json > gpurun_out/f_profile.log 2>&1
cp gpurun_out/r2_kernel_profile.json profiles/r2_kernel_profile.json
timeout 600 python bench.py > gpurun_out/r2_bench_line_1gpu.json 2> gpurun_out/f_bench.err
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_line_reference_arm.json 2> gpurun_out/f_bench_ref.err
timeout 900 python bench.py --workload umi5kx2k > gpurun_out/r2_bench_line_umi5kx2k.json 2> gpurun_out/f_bench_umi.err
timeout 600 python bench.py --workload bc737k_ed1 > gpurun_out/r2_bench_line_bc737k_ed1.json 2> gpurun_out/f_bench_737.err
timeout 600 python tools/bench_umi_cluster.py > gpurun_out/r2_umi_cluster_bench_line.json 2> gpurun_out/f_bench_ucl.err
timeout 600 python tools/bench_guided.py --flavour umi --ed 2 > gpurun_out/r2_guided_bench_line_umi_ed2.json 2> gpurun_out/f_bench_g.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --reads 2000000 --no-cpu-baseline > gpurun_out/f_ncu_bench.log 2>&1
tail -3 gpurun_out/f_tests.log; tail -2 gpurun_out/f_smoke.log; cut -c1-600 gpurun_out/r2_bench_line_1gpu.json
