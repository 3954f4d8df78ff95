mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_bench_line_1gpu.json 2> gpurun_out/l_bench.err
timeout 900 python bench.py --workload umi5kx2k > gpurun_out/r2_bench_line_umi5kx2k.json 2> gpurun_out/l_bench_umi.err
timeout 600 python bench.py --workload bc737k_ed1 > gpurun_out/r2_bench_line_bc737k_ed1.json 2> gpurun_out/l_bench_737.err
for f in 1gpu umi5kx2k bc737k_ed1; do python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_bench_line_$f.json').read().strip().split('\n')[-1]); print('$f', round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), d['parity_all_ranks'], d['parity_rank0'])"; done
