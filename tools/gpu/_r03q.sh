set -u
out=gpurun_out/r03q; mkdir -p $out
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$? $(( $(date +%s) - t0 )) s"; tail -2 $out/smoke.log
t0=$(date +%s)
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $out/pytest.log; echo "pytest $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"
python tools/bench_brief.py $out/bench_default.json
t0=$(date +%s)
timeout 900 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference rc=$? $(( $(date +%s) - t0 )) s"; tail -c 600 $out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_lih_10q.csv python bench.py --workload lih_10q --steps 2 --warmup 3 --tune --no-others > $out/launches.log 2>&1
python tools/launch_table.py $out/launches_lih_10q.csv | tail -16
