set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03n; mkdir -p $out
run() { # name, workload, env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --tune --no-others > $out/bench_${name}_$wl.json 2> $out/bench_${name}_$wl.err
  echo -n "$name: "; python tools/bench_brief.py $out/bench_${name}_$wl.json
}
for wl in lih_10q h2o_8q qaoa_7q h2_4q; do
  run old $wl QAS_B200_LIB=$PWD/qas_b200/lib/libqas_old.so
  run new $wl QAS_B200_LIB=$PWD/qas_b200/lib/libqas_b200.so
  run new_dense_reduce $wl QAS_B200_LIB=$PWD/qas_b200/lib/libqas_b200.so QAS_B200_SPARSE_REDUCE=0
done
QAS_B200_STAGE_TIMES=1 timeout 300 python bench.py --workload lih_10q --steps 3 --warmup 3 --tune --no-others 2>&1 >/dev/null | grep "B=65536" | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
