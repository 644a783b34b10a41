set -x
mkdir -p gpurun_out/r2c
python -m pytest tests/test_gpu_detector_golden.py -q -s > gpurun_out/r2c/golden.log 2>&1
grep -n "^\[\|passed\|failed\|Error" gpurun_out/r2c/golden.log | head -60
( time python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_detector_golden.py ) > gpurun_out/r2c/gpu_tests.log 2>&1
tail -5 gpurun_out/r2c/gpu_tests.log
