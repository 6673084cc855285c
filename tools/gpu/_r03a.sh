set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03a; mkdir -p $out
timeout 600 python -m pytest tests -x -q -m gpu -k "tiled or qubit_count_sweep or deep_tapes or kagome" 2>&1 | tail -2
for v in "QAS_B200_TILED_CS=1" "QAS_B200_TILED_CS=4"; do echo "== $v"; env $v timeout 300 python tools/tiled_stats.py --gpu --batch 2048 --sample 1 2>&1 | grep -v "^\[build\]" | tail -7; done | tee $out/stats.txt
for B in 25 50 100; do echo "== B=$B"; timeout 300 python tools/tiled_stats.py --gpu --batch $B --sample 1 2>&1 | grep "eval kernel"; done | tee -a $out/stats.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_tiled -s 1 -c 1 -o $out/ncu_tiled_kagome -f env QAS_B200_TILED_CS=4 python tools/tiled_stats.py --gpu --batch 1184 --sample 1 > $out/ncu.log 2>&1
python tools/ncu_summary.py $out/ncu_tiled_kagome.ncu-rep > $out/ncu_tiled_kagome.txt 2>&1; grep -E "duration|warps_active|issue_active|fp64_cycles|wavefronts_mem_shared|bank_conflicts|stalled_(short|barrier|wait|long|math)" $out/ncu_tiled_kagome.txt
