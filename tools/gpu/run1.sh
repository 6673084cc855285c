# usage: bash tools/gpu/run1.sh [tag] -- gpu parity tests + tuning benches, logs under gpurun_out/
tag=${1:-c}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log
for w in lih_10q h2o_8q qaoa_7q h2_4q; do python bench.py --tune --steps 5 --workload $w 2>&1 | tail -1 > gpurun_out/${tag}_$w.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_$w.json").read()); print(d['config']['workload'], "%.3g evals/s"%d['value'], "step %.3f ms"%d['ms_per_step'], "kernel %.3f ms"%d['roofline']['kernel_ms'], "frac %.3f"%d['roofline']['frac'], "e2e %.3g"%d['e2e']['value'])
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/${tag}_$w.json").read()[-2000:])
PY
done
