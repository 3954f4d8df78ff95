mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bc_assign -c 1 -o gpurun_out/r2f_bc_full python tools/prof_bc.py 3000000 3000000 2 1000000 1 > gpurun_out/r2f_ncu.log 2>&1; tail -1 gpurun_out/r2f_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umi_assign_deep -c 3 -o gpurun_out/r2f_deep_full python tools/perf_deep.py 20000 > gpurun_out/r2f_ncu_deep.log 2>&1; tail -1 gpurun_out/r2f_ncu_deep.log
ls -la gpurun_out/r2f_*.ncu-rep
