set -x
O=gpurun_out/r2s; mkdir -p $O
run() { name=$1; shift; env "$@" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --dump-ops $O/ops_$name.json > $O/bench_$name.json 2> $O/$name.err; python - <<PY
import json
try:
    d=json.load(open('$O/bench_$name.json')); o={x['op']:x['ms'] for x in json.load(open('$O/ops_$name.json'))}
    print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['conv_ms_per_step'], {k:v for k,v in o.items() if 'fpn_convs.3' in k})
except Exception as e: print('$name ERR', e); print(open('$O/$name.err').read()[-800:])
PY
}
run ks4_pair IOU_P6_KSPLIT=4 IOU_P6_PAIR=1
run ks8_pair IOU_P6_KSPLIT=8 IOU_P6_PAIR=1
run ks4_nopair IOU_P6_KSPLIT=4 IOU_P6_PAIR=0
run ks2_pair IOU_P6_KSPLIT=2 IOU_P6_PAIR=1
IOU_P6_KSPLIT=4 python -m pytest tests/test_gpu_detector_golden.py -q -x -k "r50_full_size_default" 2>&1 | tail -2
IOU_P6_KSPLIT=8 python -m pytest tests/test_gpu_detector_golden.py -q -x -k "r50_full_size_default" 2>&1 | tail -2
