# per-launch device times of one bench step (cold-cache, serialised: compare shares): bash tools/gpu/launches.sh <workload> <tag>
w=${1:-lih_10q}; tag=${2:-launches}
mkdir -p gpurun_out
python bench.py --tune --steps 1 --warmup 3 --workload $w > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 8 --csv --log-file gpurun_out/${tag}.csv \
    python bench.py --tune --steps 1 --warmup 3 --workload $w > gpurun_out/${tag}_ncu.log 2>&1
python - gpurun_out/${tag}.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; iK=hdr.index("Kernel Name"); iV=hdr.index("Metric Value")
for r in rows[1:]: print("%-60s %s us"%(r[iK][:60], r[iV]))
PY
