# large-state mode: parity tests + 28/30-qubit benches with and without in-pass CNOT folding
tag=${1:-sv3}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_largestate.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
for cfg in "28 4 1" "28 4 0" "30 2 1"; do set -- $cfg
  QAS_B200_SV_FOLD=$3 python tools/bench_largestate.py --qubits $1 --layers $2 --reps 2 2>&1 | tail -1 > gpurun_out/${tag}_$1q_L$2_f$3.json
  python - gpurun_out/${tag}_$1q_L$2_f$3.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], "%.2f evals/s"%d['value'], "passes",d['passes'],"gate ms %.1f"%d['gate_kernel_ms'], "total %.1f ms"%(1e3*d['total_s']), "HBM frac %.3f"%d['roofline']['frac'], "E",d['energy'])
except Exception as e:
    print("failed",e,open(sys.argv[1]).read()[-800:])
PY
done
