set -x
O=gpurun_out/r2o; mkdir -p $O
( time python -m pytest tests -q -m gpu ) > $O/gpu_tests.log 2>&1
grep -n "passed\|failed" $O/gpu_tests.log | tail -2
python bench.py --steps 20 --warmup 5 > $O/bench_line_r50.json 2> $O/bench.err
python -c "
import json; d=json.load(open('$O/bench_line_r50.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['frac_burst'], d['roofline']['traffic'], d['gpu_launches'])"
