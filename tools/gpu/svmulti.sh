# sharded large-state runs on N GPUs of one box: bash tools/gpu/svmulti.sh <N> <qubits> <layers> <tag>
N=${1:-2}; Q=${2:-29}; L=${3:-4}; tag=${4:-svm}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/bench_largestate.py --qubits $Q --layers $L --reps 2 > gpurun_out/${tag}_${Q}q_${N}gpu.log 2>&1
tail -1 gpurun_out/${tag}_${Q}q_${N}gpu.log > gpurun_out/${tag}_${Q}q_${N}gpu.json
python - gpurun_out/${tag}_${Q}q_${N}gpu.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], "%.2f evals/s"%d['value'], "gpus",d['n_gpus'],"passes",d['passes'],"swaps",d['swaps'],"gate ms %.1f"%d['gate_kernel_ms'], "total %.1f ms"%(1e3*d['total_s']), "HBM frac %.3f"%d['roofline']['frac'], "E",d['energy'])
except Exception as e:
    print("failed",e,open(sys.argv[1]).read()[-1500:])
PY
