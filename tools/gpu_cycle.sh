#!/bin/bash
# one GPU iteration of the kernel work: parity tests, kernel timings, bench line, launch list, ncu --set full captures (tag = $1)
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
python tools/perf_bc.py 3000000 3000000 2 2000000 5 | tee gpurun_out/${tag}_perf.log
for v in sicelore-2.1_b200/libslr_var_*.so; do [ -f "$v" ] && SLR_LIB_GPU=$PWD/$v python tools/perf_bc.py 3000000 3000000 2 2000000 5 | tee -a gpurun_out/${tag}_perf.log; done
python tools/perf_bc.py 737280 737 1 10000000 5 | tee -a gpurun_out/${tag}_perf.log
python tools/perf_umi.py 10000000 | tee -a gpurun_out/${tag}_perf.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --reads 2000000 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bc_assign -c 1 -o gpurun_out/${tag}_bc_full python tools/prof_bc.py 3000000 3000000 2 1000000 1 > gpurun_out/${tag}_ncu.log 2>&1; tail -1 gpurun_out/${tag}_ncu.log
ncu --set full --clock-control none --import-source on -k regex:umi_pairs -c 1 -o gpurun_out/${tag}_umi_full python tools/perf_umi.py 2000000 4 2000 1 > gpurun_out/${tag}_ncu_umi.log 2>&1; tail -1 gpurun_out/${tag}_ncu_umi.log
