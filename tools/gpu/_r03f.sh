set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03g; mkdir -p $out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $out/pytest.log
for wl in h2o_8q qaoa_7q lih_10q; do
  QAS_B200_STAGE_TIMES=1 timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --tune --no-others > $out/stages_$wl.json 2> $out/stages_$wl.err
  echo "== $wl"; grep "stage times" $out/stages_$wl.err | grep "B=65536" | tail -2
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --tune --no-others > $out/bench_$wl.json 2> $out/bench_$wl.err
  python tools/bench_brief.py $out/bench_$wl.json
done
for wl in lih_12q h2_4q; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --tune --no-others > $out/bench_$wl.json 2> $out/bench_$wl.err
  python tools/bench_brief.py $out/bench_$wl.json
done
