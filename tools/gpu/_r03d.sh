set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03d; mkdir -p $out
for T in 32 64 16; do
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:evalc_kernel<\(int\)$T, \(bool\)1" -s 3 -c 1 -o $out/ncu_evalc_T$T -f \
    python bench.py --workload lih_10q --steps 2 --warmup 3 --tune --no-others > $out/ncu_T$T.log 2>&1
  echo "T=$T rc=$?"
  python tools/ncu_summary.py $out/ncu_evalc_T$T.ncu-rep > $out/ncu_evalc_T$T.txt 2>&1
  python tools/ncu_srclines.py $out/ncu_evalc_T$T.ncu-rep 70 > $out/ncu_evalc_T${T}_lines.txt 2>&1
  grep -E "duration|grid_size|warps_active|issue_active|fp64_cycles" $out/ncu_evalc_T$T.txt
done
