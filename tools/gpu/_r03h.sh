set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03h; mkdir -p $out
for K in compile_tapes_kernel eval_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^$K" -s 3 -c 1 -o $out/ncu_h2o_$K -f \
    python bench.py --workload h2o_8q --steps 2 --warmup 3 --tune --no-others > $out/ncu_$K.log 2>&1
  echo "$K rc=$?"
  python tools/ncu_summary.py $out/ncu_h2o_$K.ncu-rep > $out/ncu_h2o_$K.txt 2>&1
  python tools/ncu_srclines.py $out/ncu_h2o_$K.ncu-rep 80 > $out/ncu_h2o_${K}_lines.txt 2>&1
  grep -E "kernel:|duration|warps_active|issue_active|inst_executed.sum|thread_inst" $out/ncu_h2o_$K.txt
done
