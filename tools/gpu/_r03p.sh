set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03p; mkdir -p $out
run() { # name, workload, env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --tune --no-others > $out/bench_${name}_$wl.json 2> $out/bench_${name}_$wl.err
  echo -n "$name: "; python tools/bench_brief.py $out/bench_${name}_$wl.json
}
run old lih_10q QAS_B200_LIB=$PWD/qas_b200/lib/libqas_old.so
run new lih_10q A=0
run new_ag7 lih_10q QAS_B200_C_AG1=7
run new_dense lih_10q QAS_B200_SPARSE_REDUCE=0
run old2 lih_10q QAS_B200_LIB=$PWD/qas_b200/lib/libqas_old.so
run new2 lih_10q A=0
run new_ag7_2 lih_10q QAS_B200_C_AG1=7
for wl in lih_12q h2o_14q; do
  for v in 3 7; do echo -n "$wl AG1=$v: "; QAS_B200_C_AG1=$v timeout 300 python bench.py --workload $wl --batch 16384 --steps 10 --warmup 3 --tune --no-others > $out/bench_${wl}_$v.json 2> $out/bench_${wl}_$v.err; python tools/bench_brief.py $out/bench_${wl}_$v.json; done
done
QAS_B200_STAGE_TIMES=1 timeout 300 python bench.py --workload lih_10q --steps 3 --warmup 3 --tune --no-others 2>&1 >/dev/null | grep "B=65536" | tail -1
QAS_B200_C_AG1=7 timeout 900 python -m pytest tests -x -q -m gpu -k "compressed or property or config_batch or uint8 or qubit_count" 2>&1 | tail -2
