#!/bin/bash
# Two independent one-GPU bench processes side by side on a multi-GPU box against one alone: separates what the
# box costs a rank (host cores, power, PCIe) from what the exchange step costs.
#   tools/gpu/contention.sh <tag> [<extra env>]
export QAS_B200_NO_REBUILD=1
out=gpurun_out/$1; mkdir -p "$out"
args="--workload lih_10q --steps 20 --warmup 5 --tune --no-others"
nproc > "$out/nproc.txt"; nvidia-smi --query-gpu=index,clocks.sm,power.draw,power.limit --format=csv >> "$out/nproc.txt"
CUDA_VISIBLE_DEVICES=0 python bench.py $args > "$out/solo.json" 2> "$out/solo.err"
CUDA_VISIBLE_DEVICES=0 python bench.py $args > "$out/pair0.json" 2> "$out/pair0.err" &
p0=$!
CUDA_VISIBLE_DEVICES=1 python bench.py $args > "$out/pair1.json" 2> "$out/pair1.err"
wait $p0
CUDA_VISIBLE_DEVICES=0 taskset -c 0-3 python bench.py $args > "$out/pin0.json" 2> "$out/pin0.err" &
p0=$!
CUDA_VISIBLE_DEVICES=1 taskset -c 8-11 python bench.py $args > "$out/pin1.json" 2> "$out/pin1.err"
wait $p0
cat "$out/nproc.txt"
for f in solo pair0 pair1 pin0 pin1; do echo -n "$f: "; python tools/bench_brief.py "$out/$f.json"; done
