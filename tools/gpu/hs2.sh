tag=${1:-hs2}
mkdir -p gpurun_out
for w in lih_10q h2o_8q; do for hs in 1; do
  f=gpurun_out/${tag}_${w}_hs${hs}.json
  QAS_B200_HSPARSE=$hs python bench.py --tune --steps 5 --workload $w 2>&1 | tail -1 > $f
  python - $f "$w hs=$hs" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); print(sys.argv[2], "%.3g evals/s"%d['value'], "step %.3f ms"%d['ms_per_step'], "kernel %.3f ms"%d['roofline']['kernel_ms'], "frac %.3f"%d['roofline']['frac'], "e2e %.3g"%d['e2e']['value'], "reward-only %.3g"%d['reward_only']['value'])
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
done; done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
