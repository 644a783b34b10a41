set -x
O=gpurun_out/r2q; mkdir -p $O
for t in 512 384 256; do
IOU_NMS_THREADS=$t timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k "regex:class_nms|final_select" --csv --log-file $O/post_bs8_$t.csv python bench.py --no-graph --ncu-range --no-cpu-baseline --steps 1 --warmup 3 > $O/post_$t.out 2>&1
python - <<PY
import csv
rows=[l for l in open('$O/post_bs8_$t.csv') if l.startswith('"')]
for r in csv.DictReader(rows): print($t, r['Kernel Name'][:24], r['Block Size'], r['Metric Value'])
PY
IOU_NMS_THREADS=$t python tools/bench_postproc.py | cut -c80-220
done
IOU_NMS_THREADS=384 python -m pytest tests/test_gpu_postproc.py -q -x 2>&1 | tail -2
python -m pytest tests/test_gpu_postproc.py -q -x 2>&1 | tail -2
