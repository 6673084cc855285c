set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03i; mkdir -p $out
for rep in 1 2; do
for lib in old b200; do
  export QAS_B200_LIB=$PWD/qas_b200/lib/libqas_$lib.so
  echo "== $lib (rep $rep)"
  timeout 300 python tools/e2e_breakdown.py --workload lih_10q --steps 300 2>&1 | grep -v "^\[build\]" | tee -a $out/e2e_$lib.txt
done
done
export QAS_B200_LIB=$PWD/qas_b200/lib/libqas_b200.so
timeout 300 python tools/e2e_breakdown.py --workload h2o_8q --steps 100 2>&1 | grep -v "^\[build\]" | tee $out/e2e_h2o.txt
unset QAS_B200_LIB
timeout 600 python -m pytest tests -x -q -m gpu -k "edge_cases or device_pointer or model_methods or golden or uint8 or search_and_tuning" 2>&1 | tail -2
