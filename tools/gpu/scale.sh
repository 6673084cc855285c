# N-GPU scaling of the default bench on one box: bash tools/gpu/scale.sh <N> <tag>
N=${1:-2}; tag=${2:-scale}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --tune > gpurun_out/${tag}_n$N.json 2> gpurun_out/${tag}_n$N.err
tail -1 gpurun_out/${tag}_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('n_gpus',d['n_gpus'],'value %.4g'%d['value'],'ms %.3f'%d['ms_per_step'],'e2e %.4g'%d['e2e']['value'])" || tail -5 gpurun_out/${tag}_n$N.err
