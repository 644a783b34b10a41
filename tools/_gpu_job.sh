set -x
O=gpurun_out/r2h; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --model x101_64x4d > $O/bench_line_x101_64x4d_8gpu.json 2> $O/x8.err
tail -c 900 $O/bench_line_x101_64x4d_8gpu.json | head -c 400; python -c "
import json; d=json.load(open('$O/bench_line_x101_64x4d_8gpu.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_line_r50_8gpu.json 2> $O/r8.err
python -c "
import json; d=json.load(open('$O/bench_line_r50_8gpu.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_line_r50_1gpu_same_box.json 2> $O/r1.err
python -c "
import json; d=json.load(open('$O/bench_line_r50_1gpu_same_box.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
