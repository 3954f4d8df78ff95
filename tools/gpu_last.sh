mkdir -p gpurun_out
timeout 900 python bench.py --workload umi5kx2k > gpurun_out/r2_bench_line_umi5kx2k.json 2> gpurun_out/l_bench_umi.err
timeout 900 python tools/perf_deep.py 301 1000 4097 5001 7000 20000 2>&1 | grep -v "phases\|dbg" | tee gpurun_out/d7_perf_deep.log
timeout 600 python tools/perf_assign.py 2000000 60 2000 3 | cut -c1-600 | tee gpurun_out/d7_perf_assign_m60.log
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_line_umi5kx2k.json').read().strip().split('\n')[-1]); print('umi5kx2k', round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), d['parity_all_ranks'], d['cpu_baseline'])"
