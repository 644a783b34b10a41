set -x
O=gpurun_out/r2v; mkdir -p $O
( time python -m pytest tests -q -m gpu ) > $O/gpu_tests.log 2>&1
grep -n "passed\|failed" $O/gpu_tests.log | tail -2
python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_line_reference_arm.json 2> $O/ref.err
python bench.py --steps 20 --warmup 5 --dump-ops $O/ops_r50.json > $O/bench_line_r50.json 2> $O/bench.err
python -c "
import json; d=json.load(open('$O/bench_line_r50.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['frac_burst'], d['postproc_ms_per_step'], d['cpu_baseline']['value'])"
for m in r101 x101_64x4d; do python bench.py --model $m --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_line_$m.json 2> $O/$m.err; python -c "
import json; d=json.load(open('$O/bench_line_$m.json')); print('$m', d['value'], d['ms_per_step'], d['e2e']['value'])"; done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --weights reference-init > $O/bench_line_r50_reference_init.json 2> $O/ri.err; python -c "
import json; d=json.load(open('$O/bench_line_r50_reference_init.json')); print('ri', d['value'], d['ms_per_step'], d['e2e']['value'], d['postproc_ms_per_step'])"
python tools/bench_postproc.py > $O/config5_postproc_microbench.json 2>/dev/null; cut -c80-330 $O/config5_postproc_microbench.json
