#!/bin/bash
# torchrun bench line and the single-process slr_multi_* line at N GPUs:  bash tools/gpu_scale.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 \
  > gpurun_out/r2_bench_line_${N}gpu.json 2> gpurun_out/s_bench_${N}gpu.err
timeout 900 python tools/bench_multi.py $N > gpurun_out/r2_bench_multi_${N}gpu.json 2> gpurun_out/s_multi_${N}gpu.err
cut -c1-700 gpurun_out/r2_bench_line_${N}gpu.json; cut -c1-500 gpurun_out/r2_bench_multi_${N}gpu.json; tail -n 3 gpurun_out/s_bench_${N}gpu.err; tail -n 3 gpurun_out/s_multi_${N}gpu.err
