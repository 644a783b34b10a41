set -x
O=gpurun_out/r2u; mkdir -p $O
for rep in 1 2 3; do for p in 2 3 4; do
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --plans $p > $O/bench_p${p}_$rep.json 2> $O/p${p}_$rep.err
python -c "
import json; d=json.load(open('$O/bench_p${p}_$rep.json')); print('plans', $p, 'rep', $rep, d['value'], d['ms_per_step'])"
done; done
