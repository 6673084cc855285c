# ncu --set full capture of the evaluation kernel on one workload: bash tools/gpu/prof.sh <workload> <tag>
w=${1:-lih_10q}; tag=${2:-prof}
mkdir -p gpurun_out
python bench.py --tune --steps 1 --warmup 3 --workload $w > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 3 -c 1 -f -o gpurun_out/${tag} \
    python bench.py --tune --steps 1 --warmup 3 --workload $w > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_plain.log; tail -3 gpurun_out/${tag}_ncu.log
