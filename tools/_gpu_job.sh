set -x
mkdir -p gpurun_out/r2b
python -m pytest tests/test_gpu_detector_golden.py -x -q -s > gpurun_out/r2b/golden.log 2>&1
tail -5 gpurun_out/r2b/golden.log
python tools/l2_probe.py --batch 1 2 4 8 > gpurun_out/r2b/l2_probe.log 2>&1
tail -6 gpurun_out/r2b/l2_probe.log
timeout 600 ncu --profile-from-start off --cache-control none --clock-control none --metrics dram__bytes_read.sum --csv --log-file gpurun_out/r2b/l2_once_b1.csv python tools/l2_probe.py --batch 1 --once > gpurun_out/r2b/ncu_once.out 2>&1
