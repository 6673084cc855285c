tag=${1:-t}
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); print(sys.argv[2], d['config']['workload'], "B",d['config']['circuits_per_gpu_per_step'], "%.3g evals/s"%d['value'], "step %.3f ms"%d['ms_per_step'], "kernel %.3f ms"%d['roofline']['kernel_ms'], "frac %.3f"%d['roofline']['frac'], "reward-only %.3g"%d['reward_only']['value'])
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
}
for cfg in "kagome_16q 1024"; do set -- $cfg
  f=gpurun_out/${tag}_$1.json
  timeout 600 python bench.py --tune --steps 3 --workload $1 --batch $2 2>&1 | tail -1 > $f; show $f ""
done
