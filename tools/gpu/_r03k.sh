set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03k; mkdir -p $out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload lih_10q --steps 40 --warmup 5 --tune --no-others > $out/bench_$name.json 2> $out/bench_$name.err
  echo -n "$name: "; python tools/bench_brief.py $out/bench_$name.json
}
run old QAS_B200_LIB=$PWD/qas_b200/lib/libqas_old.so
run new QAS_B200_LIB=$PWD/qas_b200/lib/libqas_b200.so
run new_blk64 QAS_B200_LIB=$PWD/qas_b200/lib/libqas_b200.so QAS_B200_C_BLK64=1
run old2 QAS_B200_LIB=$PWD/qas_b200/lib/libqas_old.so
run new2 QAS_B200_LIB=$PWD/qas_b200/lib/libqas_b200.so
for wl in lih_12q h2o_14q; do
  for lib in old b200; do
    echo -n "$wl $lib: "; QAS_B200_LIB=$PWD/qas_b200/lib/libqas_$lib.so timeout 300 python bench.py --workload $wl --batch 16384 --steps 10 --warmup 3 --tune --no-others > $out/bench_${wl}_$lib.json 2> $out/bench_${wl}_$lib.err; python tools/bench_brief.py $out/bench_${wl}_$lib.json
  done
done
timeout 300 python tools/phase_clocks.py --workload lih_10q > $out/phases.txt 2>&1; grep -A6 "slot 7" $out/phases.txt
timeout 900 python -m pytest tests -x -q -m gpu -k "compressed or property or config_batch or qubit_count_sweep or baseline_size or uint8" 2>&1 | tail -2
