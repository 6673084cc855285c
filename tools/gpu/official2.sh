# bench lines of record + launch list + ncu capture (the short form of official.sh): bash tools/gpu/official2.sh <tag>
tag=${1:-r01h}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench_lih_10q.json 2> gpurun_out/${tag}_bench_lih_10q.err; tail -c 300 gpurun_out/${tag}_bench_lih_10q.json
for w in h2o_8q h2_4q qaoa_7q; do python bench.py --workload $w --steps 5 > gpurun_out/${tag}_bench_$w.json 2>/dev/null; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_reference_lih_10q.json 2>/dev/null
bash tools/gpu/launches.sh lih_10q ${tag}_launches_lih_10q > /dev/null
bash tools/gpu/launches.sh h2o_8q ${tag}_launches_h2o_8q > /dev/null
bash tools/gpu/prof.sh lih_10q ${tag}_prof_lih > /dev/null
echo done
