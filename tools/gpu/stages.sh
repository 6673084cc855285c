# per-stage stream times of qas_eval_batch (QAS_B200_STAGE_TIMES=1) for a workload: bash tools/gpu/stages.sh <workload>
w=${1:-lih_10q}
for ch in 32 64 128; do
  echo "chunk $ch"; QAS_B200_REDUCE_CHUNK=$ch QAS_B200_STAGE_TIMES=1 python bench.py --tune --steps 2 --warmup 3 --workload $w 2>&1 | grep -i "chunk_reduce\|stage" | tail -2
done
