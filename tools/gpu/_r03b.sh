export QAS_B200_NO_REBUILD=1
timeout 600 python -m pytest tests -x -q -m gpu -k "tiled or kagome" 2>&1 | tail -2
for v in "QAS_B200_TILED_CS=1" "QAS_B200_TILED_CS=4"; do echo "== $v"; env $v timeout 300 python tools/tiled_stats.py --gpu --batch 2048 --sample 1 2>&1 | grep -E "eval kernel|forward |adjoint "; done
