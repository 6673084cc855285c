tag=${1:-t}
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); print(sys.argv[2], d['config']['workload'], "%.3g evals/s"%d['value'], "step %.3f ms"%d['ms_per_step'], "kernel %.3f ms"%d['roofline']['kernel_ms'], "frac %.3f"%d['roofline']['frac'], "reward-only %.3g"%d['reward_only']['value'])
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
}
for w in lih_10q h2o_8q; do
for cfg in "" "2,4" "2,1" "3,2" "1,2"; do
  f=gpurun_out/${tag}_${w}_h${cfg/,/_}.json
  QAS_B200_HCFG=$cfg python bench.py --tune --steps 5 --workload $w 2>&1 | tail -1 > $f; show $f "HCFG=$cfg"
done
done
