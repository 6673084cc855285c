# ncu --set full capture of the large-state tile kernel: bash tools/gpu/svprof.sh <tag>
tag=${1:-svp}
mkdir -p gpurun_out
python tools/bench_largestate.py --qubits 27 --layers 4 --reps 1 > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sv_block -s 3 -c 3 -f -o gpurun_out/${tag} \
    python tools/bench_largestate.py --qubits 27 --layers 4 --reps 1 > gpurun_out/${tag}_ncu.log 2>&1
tail -1 gpurun_out/${tag}_plain.log | cut -c1-600; tail -2 gpurun_out/${tag}_ncu.log
