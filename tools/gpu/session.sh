#!/bin/bash
# One gpurun session: a list of steps, each under its own timeout, logs under gpurun_out/<tag>/.
#   tools/gpu/session.sh <tag> <step> [<step> ...]
# steps: tests | tests:<pytest -k expr> | bench:<workload>[:<env assignments,comma separated>] | stages:<workload> | phases:<workload> | svbench:<workload>[:envs]
#        | launches:<workload> | ncu:<workload>:<kernel regex> | sanity
set -u
export QAS_B200_NO_REBUILD=1
tag=$1; shift
out=gpurun_out/$tag
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$out/smi.csv" 2>&1
for step in "$@"; do
  IFS=: read -r kind a b <<< "$step"
  name=$(echo "$step" | tr ':/ =,*|()' '_________')
  echo "== $step" | tee -a "$out/steps.log"
  envs=""
  case $kind in
    tests)
      if [ -n "${a:-}" ]; then timeout 1500 python -m pytest tests -x -q -m gpu -k "$a" > "$out/$name.log" 2>&1
      else timeout 2400 python -m pytest tests -x -q -m gpu > "$out/$name.log" 2>&1; fi
      echo "rc=$?" | tee -a "$out/steps.log"; tail -5 "$out/$name.log" ;;
    bench)
      envs=$(echo "${b:-}" | tr ',' ' ')
      env $envs timeout 600 python bench.py --workload "$a" --steps 20 --warmup 5 --tune > "$out/$name.json" 2> "$out/$name.err"
      echo "rc=$?" | tee -a "$out/steps.log"; python tools/bench_brief.py "$out/$name.json" ;;
    fullbench)
      envs=$(echo "${b:-}" | tr ',' ' ')
      env $envs timeout 900 python bench.py --workload "$a" --steps 20 --warmup 5 > "$out/$name.json" 2> "$out/$name.err"
      echo "rc=$?" | tee -a "$out/steps.log"; python tools/bench_brief.py "$out/$name.json" ;;
    stages)
      envs=$(echo "${b:-}" | tr ',' ' ')
      env $envs QAS_B200_STAGE_TIMES=1 timeout 300 python bench.py --workload "$a" --steps 3 --warmup 3 --tune > "$out/$name.json" 2> "$out/$name.err"
      echo "rc=$?" | tee -a "$out/steps.log"; grep "stage times" "$out/$name.err" | sort | uniq -c | sort -rn | head -4 ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/$name.csv" \
        python bench.py --workload "$a" --steps 2 --warmup 3 --tune > "$out/$name.log" 2>&1
      echo "rc=$?" | tee -a "$out/steps.log"; python tools/launch_table.py "$out/$name.csv" | tail -25 ;;
    ncu)     # ncu:<workload>:<kernel base name>:<launches to skip>
      IFS=: read -r _k _w kname skip <<< "$step"
      timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$kname" -s "${skip:-6}" -c 1 -o "$out/$name" -f \
        python bench.py --workload "$_w" --steps 2 --warmup 3 --tune --no-others > "$out/$name.log" 2>&1
      echo "rc=$?" | tee -a "$out/steps.log"; python tools/ncu_summary.py "$out/$name.ncu-rep" > "$out/$name.txt" 2>&1; head -60 "$out/$name.txt" ;;
    phases)
      timeout 300 python tools/phase_clocks.py --workload "$a" > "$out/$name.txt" 2>&1
      echo "rc=$?" | tee -a "$out/steps.log"; tail -40 "$out/$name.txt" ;;
    svbench)
      envs=$(echo "${b:-}" | tr ',' ' ')
      env $envs timeout 600 python bench.py --workload "$a" --steps 5 --warmup 3 --tune > "$out/$name.json" 2> "$out/$name.err"
      echo "rc=$?" | tee -a "$out/steps.log"; tail -c 1500 "$out/$name.json"; tail -3 "$out/$name.err" ;;
    mgpu)    # mgpu:<N>:<workload>[:envs]  -- bench.py under torchrun on N GPUs of the box
      IFS=: read -r _k ng wl envs <<< "$step"
      envs=$(echo "${envs:-}" | tr ',' ' ')
      env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$ng" --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus "$ng" --workload "$wl" --steps 20 --warmup 5 --tune > "$out/$name.json" 2> "$out/$name.err"
      echo "rc=$?" | tee -a "$out/steps.log"; python tools/bench_brief.py "$out/$name.json"; tail -c 1200 "$out/$name.json"; tail -5 "$out/$name.err" ;;
    gputests2)   # the multi-GPU pytest cases (need >= 2 GPUs)
      timeout 900 python -m pytest tests -x -q -m gpu -k "two_gpu" > "$out/$name.log" 2>&1
      echo "rc=$?" | tee -a "$out/steps.log"; tail -5 "$out/$name.log" ;;
    searchepoch)
      timeout 900 python bench.py --search-epoch --workload "$a" > "$out/$name.json" 2> "$out/$name.err"
      echo "rc=$?" | tee -a "$out/steps.log"; tail -c 1500 "$out/$name.json"; tail -3 "$out/$name.err" ;;
    leafstudy)   # leafstudy:<LiH|H2O>:<seeds>
      timeout 1500 python tools/leaf_parallel_study.py "$a" --seeds "${b:-20}" > "$out/$name.json" 2> "$out/$name.err"
      echo "rc=$?" | tee -a "$out/steps.log"; python - "$out/$name.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
for k,v in d.items():
    if isinstance(v,dict): print(k,{a:b for a,b in v.items() if a!="runs"})
    else: print(k,v)
PY
      tail -3 "$out/$name.err" ;;
    sanity)
      timeout 600 python tools/gpu/sanity.py > "$out/$name.log" 2>&1
      echo "rc=$?" | tee -a "$out/steps.log"; tail -3 "$out/$name.log" ;;
    *) echo "unknown step $step" ;;
  esac
done
