set -x
O=gpurun_out/r2m; mkdir -p $O
timeout 600 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file $O/launches_ncu_dram.csv python bench.py --no-graph --ncu-range --no-cpu-baseline --steps 1 --warmup 3 > $O/ncu_launches.out 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:max_score|topk_hist|gather_decode|class_nms|final_select" -c 5 -f -o $O/ncu_postproc python bench.py --no-graph --ncu-range --no-cpu-baseline --steps 1 --warmup 3 > $O/ncu_postproc.out 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off --launch-skip 65 --launch-count 1 -f -o $O/ncu_head_tower python bench.py --no-graph --ncu-range --no-cpu-baseline --steps 1 --warmup 3 > $O/ncu_head.out 2>&1
python bench.py --steps 20 --warmup 5 --dump-ops $O/ops_r50.json > $O/bench_line_r50.json 2> $O/bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_line_reference_arm.json 2> $O/ref.err
ls -la $O
