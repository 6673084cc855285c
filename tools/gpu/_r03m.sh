set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03m; mkdir -p $out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload lih_10q --steps 40 --warmup 5 --tune --no-others > $out/bench_$name.json 2> $out/bench_$name.err
  echo -n "$name: "; python tools/bench_brief.py $out/bench_$name.json
}
run gate_sep QAS_B200_GATE_IN_COMPILE=0
run default A=0
run chunk32 QAS_B200_REDUCE_CHUNK=32
run chunk128 QAS_B200_REDUCE_CHUNK=128
run chunk32_g2 QAS_B200_REDUCE_CHUNK=32 QAS_B200_REDUCE_GROUPS=2
run chunk128_g2 QAS_B200_REDUCE_CHUNK=128 QAS_B200_REDUCE_GROUPS=2
run gate_sep2 QAS_B200_GATE_IN_COMPILE=0
run default2 A=0
timeout 600 python -m pytest tests -x -q -m gpu -k "compressed or property or config_batch or tuning or resident" 2>&1 | tail -2
