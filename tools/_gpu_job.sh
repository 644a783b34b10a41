set -x
O=gpurun_out/r2p; mkdir -p $O
( time python -m pytest tests/test_gpu_postproc.py tests/test_gpu_detector_golden.py tests/test_gpu_model.py -q -x ) > $O/gpu_tests.log 2>&1
grep -n "passed\|failed" $O/gpu_tests.log | tail -2
python tools/bench_postproc.py > $O/config5.json 2> $O/c5.err; cat $O/config5.json | cut -c1-420
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $O/config5_launches.csv python tools/bench_postproc.py > $O/c5.out 2>&1
grep -v "^==" $O/config5_launches.csv | python -c "
import csv,sys,collections
r=csv.DictReader(sys.stdin); agg=collections.OrderedDict()
for row in r:
    agg.setdefault(row['Kernel Name'][:40],[]).append(float(row['Metric Value'].replace(',','')))
for k,v in agg.items(): print(k, round(sum(v)/len(v)/1e3,1), len(v))
"
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['conv_ms_per_step'], d['postproc_ms_per_step'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k "regex:max_score|topk_hist|gather_decode|class_nms|final_select" --csv --log-file $O/post_bs8.csv python bench.py --no-graph --ncu-range --no-cpu-baseline --steps 1 --warmup 3 > $O/post.out 2>&1
grep -v "^==" $O/post_bs8.csv | cut -d, -f5,15 | cut -c1-80
