set -u
out=gpurun_out/r03r; mkdir -p $out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests -x -q -m gpu -k "two_gpu" 2>&1 | tail -3 | tee $out/pytest_two_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 200 --warmup 5 > $out/scale_n2.json 2> $out/scale_n2.err; echo "rc=$?"
python tools/bench_brief.py $out/scale_n2.json; python - $out/scale_n2.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print({k:d.get(k) for k in ("parity_check","comm_us_per_step","n_gpus","scaling")})
PY
tail -3 $out/scale_n2.err
