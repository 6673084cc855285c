set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03j; mkdir -p $out
for rep in 1 2; do
for lib in old b200; do
  export QAS_B200_LIB=$PWD/qas_b200/lib/libqas_$lib.so
  timeout 300 python bench.py --workload lih_10q --steps 100 --warmup 5 --tune --no-others > $out/bench_${lib}_$rep.json 2> $out/bench_${lib}_$rep.err
  python - $out/bench_${lib}_$rep.json $lib <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(sys.argv[2], "value %.2f M  e2e %.2f M  e2e_pageable %.2f M" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6))
PY
done
done
unset QAS_B200_LIB
timeout 600 python -m pytest tests -x -q -m gpu -k "edge_cases or device_pointer or model_methods or golden or uint8 or search_and_tuning or config_batch" 2>&1 | tail -2
