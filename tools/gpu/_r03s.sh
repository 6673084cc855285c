set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03s; mkdir -p $out
for wl in lih_10q h2o_8q; do
  echo "== $wl"
  QAS_B200_HOST_TIMES=1 timeout 300 python tools/e2e_breakdown.py --workload $wl --steps 30 > $out/e2e_$wl.txt 2> $out/e2e_$wl.err
  grep -v "^\[build\]" $out/e2e_$wl.txt
  grep "host times" $out/e2e_$wl.err | grep "B=65536" | sed -n '8,12p;45,49p;85,89p'
done
