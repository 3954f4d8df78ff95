#!/bin/bash
# time every kernel build variant (sicelore-2.1_b200/libslr_var_*.so) against the default library: same inputs, crc of the results
tag=${1:-var}
mkdir -p gpurun_out
python tools/perf_bc.py 3000000 3000000 2 2000000 5 | tee gpurun_out/${tag}_perf.log
for v in sicelore-2.1_b200/libslr_var_*.so; do [ -f "$v" ] && SLR_LIB_GPU=$PWD/$v python tools/perf_bc.py 3000000 3000000 2 2000000 5 | tee -a gpurun_out/${tag}_perf.log; done
python tools/perf_bc.py 737280 737 1 10000000 5 | tee -a gpurun_out/${tag}_perf.log
