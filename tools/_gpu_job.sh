set -x
O=gpurun_out/r2j; mkdir -p $O
run() { name=$1; shift; env "$@" python bench.py --steps 30 --warmup 5 --no-cpu-baseline $EXTRA --dump-ops $O/ops_$name.json > $O/bench_$name.json 2> $O/$name.err; python - <<PY
import json
try:
    d=json.load(open('$O/bench_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['conv_ms_per_step'])
except Exception as e: print('$name ERR', e)
PY
}
EXTRA="" run base A=1
EXTRA="" run pair48 IOU_PAIR_MIN_BN=48
EXTRA="" run noph3 IOU_FUSE_PH3=0
EXTRA="" run respf1 IOU_RES_PREFETCH=1
EXTRA="--plans 3" run plans3 A=1
EXTRA="" run base2 A=1
