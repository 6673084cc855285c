set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03e; mkdir -p $out
for lib in old b200; do
  export QAS_B200_LIB=$PWD/qas_b200/lib/libqas_$lib.so
  for wl in h2o_8q qaoa_7q lih_10q; do
    QAS_B200_STAGE_TIMES=1 timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --tune --no-others > $out/stages_${lib}_$wl.json 2> $out/stages_${lib}_$wl.err
    echo "== $lib $wl"; grep "stage times" $out/stages_${lib}_$wl.err | grep "B=65536" | tail -3
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --tune --no-others > $out/bench_${lib}_$wl.json 2> $out/bench_${lib}_$wl.err
    python tools/bench_brief.py $out/bench_${lib}_$wl.json
  done
  QAS_B200_COMPRESS=0 QAS_B200_STAGE_TIMES=1 timeout 300 python bench.py --workload lih_10q --steps 3 --warmup 3 --tune --no-others > $out/stages_${lib}_lih_nocompress.json 2> $out/stages_${lib}_lih_nocompress.err
  echo "== $lib lih_10q COMPRESS=0"; grep "stage times" $out/stages_${lib}_lih_nocompress.err | grep "B=65536" | tail -2
done
unset QAS_B200_LIB
timeout 900 python -m pytest tests -x -q -m gpu -k "config_batch or qubit_count_sweep or deep_tapes or full_vocabulary or edge_cases or support_restricted or property" 2>&1 | tail -3
