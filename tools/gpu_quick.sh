timeout 175 python bench.py --workload umi5kx2k > gpurun_out/r2_bench_line_umi5kx2k.json 2> gpurun_out/l_bench_umi.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_line_umi5kx2k.json').read().strip().split('\n')[-1]); print('umi5kx2k', round(d['value']/1e6,1), round(d['ms_per_step'],2), d['legs_ms'], round(d['e2e']['value']/1e6,1), d['parity_all_ranks'], d['parity_rank0'], d['cpu_baseline']['value'])"
