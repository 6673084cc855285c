set -u
export QAS_B200_NO_REBUILD=1
out=gpurun_out/r03o; mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_largestate.py 2>&1 | tail -80 > $out/pytest.log
tail -5 $out/pytest.log
