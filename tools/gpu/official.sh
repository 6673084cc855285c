# the bench lines, launch list and ncu capture that go under profiles/ : bash tools/gpu/official.sh <tag>
tag=${1:-r01h}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; tail -1 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench_lih_10q.json 2> gpurun_out/${tag}_bench_lih_10q.err; tail -c 300 gpurun_out/${tag}_bench_lih_10q.json
for w in h2o_8q h2_4q qaoa_7q; do python bench.py --workload $w --steps 5 > gpurun_out/${tag}_bench_$w.json 2>/dev/null; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_reference_lih_10q.json 2>/dev/null
for w in lih_10q h2o_8q h2_4q; do python bench.py --search-epoch --workload $w 2>/dev/null | tail -1 > gpurun_out/${tag}_search_epoch_$w.json; done
python tools/phase_clocks.py --workload lih_10q 2>/dev/null | tail -6 > gpurun_out/${tag}_phase_clocks_lih_10q.txt
bash tools/gpu/launches.sh lih_10q ${tag}_launches_lih_10q > /dev/null
bash tools/gpu/prof.sh lih_10q ${tag}_prof_lih > /dev/null
echo done
