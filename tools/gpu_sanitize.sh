mkdir -p gpurun_out
for t in memcheck racecheck synccheck initcheck; do
  echo "==== $t" >> gpurun_out/r2_compute_sanitizer_smoke.txt
  timeout 1200 compute-sanitizer --tool $t --print-limit 20 python __graft_entry__.py smoke 2>&1 | grep -v "^$" | tail -12 >> gpurun_out/r2_compute_sanitizer_smoke.txt
done
echo "==== memcheck, deep jobs on a cluster of 8 CTAs and on the cooperative grid (tools/perf_deep.py 1500 4200)" >> gpurun_out/r2_compute_sanitizer_smoke.txt
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/perf_deep.py 1500 4200 2>&1 | grep -v "^$" | tail -8 >> gpurun_out/r2_compute_sanitizer_smoke.txt
echo "==== racecheck, the same" >> gpurun_out/r2_compute_sanitizer_smoke.txt
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/perf_deep.py 1500 2>&1 | grep -v "^$" | tail -8 >> gpurun_out/r2_compute_sanitizer_smoke.txt
cat gpurun_out/r2_compute_sanitizer_smoke.txt | cut -c1-250
