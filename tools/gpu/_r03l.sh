set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03l; mkdir -p $out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload lih_10q --steps 40 --warmup 5 --tune --no-others > $out/bench_$name.json 2> $out/bench_$name.err
  echo -n "$name: "; python tools/bench_brief.py $out/bench_$name.json
}
run base A=0
run ag1_s6_m5 QAS_B200_C_AG1=1
run ag1_s6_m6 QAS_B200_C_AG1=1 QAS_B200_C_AG1_MINB=6
run ag1_s8 QAS_B200_C_AG1=2
run ag1_s6s8_m5 QAS_B200_C_AG1=3
run ag1_s6s8_m6 QAS_B200_C_AG1=3 QAS_B200_C_AG1_MINB=6
run base2 A=0
QAS_B200_C_AG1=3 timeout 600 python -m pytest tests -x -q -m gpu -k "compressed or property or config_batch" 2>&1 | tail -2
