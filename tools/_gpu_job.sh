set -x
mkdir -p gpurun_out/r2d
( time python -m pytest tests -q -m gpu -x ) > gpurun_out/r2d/gpu_tests.log 2>&1
tail -40 gpurun_out/r2d/gpu_tests.log
