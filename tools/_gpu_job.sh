set -x
O=gpurun_out/r2f; mkdir -p $O
( time python -m pytest tests -q -m gpu ) > $O/gpu_tests.log 2>&1
tail -5 $O/gpu_tests.log | cut -c1-200
python bench.py --steps 20 --warmup 5 --weights reference-init > $O/bench_line_r50_reference_init.json 2> $O/ri.err
python bench.py --impl reference --steps 3 --warmup 1 --weights reference-init --reference-budget-s 120 > $O/bench_line_reference_arm_reference_init.json 2> $O/ri_ref.err
for m in r101 x101_32x4d x101_64x4d; do
python bench.py --model $m --steps 20 --warmup 5 --no-cpu-baseline --dump-ops $O/ops_$m.json > $O/bench_line_$m.json 2> $O/$m.err
done
python tools/bench_postproc.py > $O/config5_postproc_microbench.json 2> $O/c5.err
for f in $O/bench_line_*.json; do python - <<PY
import json
try:
    d=json.load(open('$f')); print('$f', d['value'], d.get('ms_per_step'), d['e2e']['value'], d.get('conv_ms_per_step'), (d.get('roofline') or {}).get('frac'), d.get('postproc_ms_per_step'))
except Exception as e: print('$f', 'ERR', e)
PY
done
cat $O/config5_postproc_microbench.json | head -c 1500
tail -3 $O/*.err
