# ncu --set full capture of one kernel (regex) of one bench workload: bash tools/gpu/prof_k.sh <workload> <kernel-regex> <tag>
w=${1:-lih_10q}; k=${2:-compile_tapes}; tag=${3:-profk}
mkdir -p gpurun_out
python bench.py --tune --steps 1 --warmup 3 --workload $w > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${tag} \
    python bench.py --tune --steps 1 --warmup 3 --workload $w > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
