set -x
O=gpurun_out/r2k; mkdir -p $O
( time python -m pytest tests -q -m gpu ) > $O/gpu_tests.log 2>&1
tail -4 $O/gpu_tests.log | cut -c1-200
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 40 --csv --log-file $O/config5_launches.csv python tools/bench_postproc.py > $O/c5.out 2>&1
grep -v "^==" $O/config5_launches.csv | python -c "
import csv,sys,collections
r=csv.DictReader(sys.stdin); agg=collections.OrderedDict()
for row in r:
    k=row['Kernel Name'][:50]; agg.setdefault(k,collections.defaultdict(list))[row['Metric Name']].append(float(row['Metric Value'].replace(',','')))
for k,v in agg.items(): print(k, {m:(round(sum(x)/len(x)/1e3,1) if 'time' in m else round(sum(x)/len(x)/1e6,1)) for m,x in v.items()}, len(list(v.values())[0]))
"
