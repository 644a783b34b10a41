set -x
O=gpurun_out/r2n; mkdir -p $O
python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -3 $O/smoke.log
run() { name=$1; shift; env "$@" python bench.py --steps 40 --warmup 5 --no-cpu-baseline $EXTRA > $O/bench_$name.json 2> $O/$name.err; python - <<PY
import json
try:
    d=json.load(open('$O/bench_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['conv_ms_per_step'])
except Exception as e: print('$name ERR', e)
PY
}
EXTRA="" run base A=1
EXTRA="" run side IOU_FPN_SIDE=1
EXTRA="" run base2 A=1
EXTRA="" run side2 IOU_FPN_SIDE=1
EXTRA="--no-pipeline" run base_np A=1
EXTRA="--no-pipeline" run side_np IOU_FPN_SIDE=1
IOU_FPN_SIDE=1 python -m pytest tests/test_gpu_detector_golden.py -q -x -k "r50_full_size_default or graph" 2>&1 | tail -2
