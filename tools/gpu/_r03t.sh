set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03t; mkdir -p $out
for wl in lih_10q h2o_8q; do
  echo "== $wl"
  QAS_B200_HOST_TIMES=1 timeout 300 python tools/e2e_breakdown.py --workload $wl --steps 100 > $out/e2e_$wl.txt 2> $out/e2e_$wl.err
  grep -v "^\[build\]" $out/e2e_$wl.txt
  grep "host times" $out/e2e_$wl.err | grep "B=65536" | sed -n '20,22p;150,152p'
done
for wl in lih_10q h2o_8q qaoa_7q h2_4q; do timeout 300 python bench.py --workload $wl --steps 100 --warmup 5 --tune --no-others > $out/bench_$wl.json 2> $out/bench_$wl.err; python tools/bench_brief.py $out/bench_$wl.json; python - $out/bench_$wl.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("   e2e_pageable %.2f M" % (d["e2e_pageable"]["value"]/1e6))
PY
done
timeout 600 python -m pytest tests -x -q -m gpu -k "edge_cases or uint8 or model_methods or golden or search_and_tuning or device_pointer" 2>&1 | tail -2
