mkdir -p gpurun_out
for v in "" sicelore-2.1_b200/libslr_var_nib0.so sicelore-2.1_b200/libslr_var_uani.so; do
  echo "== lib: ${v:-default}" | tee -a gpurun_out/v2_perf.log
  SLR_LIB_GPU=${v:+$PWD/$v} python tools/perf_bc.py 3000000 3000000 2 10000000 5 | tee -a gpurun_out/v2_perf.log
  SLR_LIB_GPU=${v:+$PWD/$v} python tools/perf_assign.py 4000000 4 2000 7 | cut -c1-330 | tee -a gpurun_out/v2_perf.log
done
