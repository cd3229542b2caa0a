set -x
o=gpurun_out/r2z
mkdir -p $o
for g in 0 1; do
python bench.py --workload table1 --steps 50 --warmup 5 --no-cpu-baseline --graph $g > $o/bench_table1_g$g.json 2> $o/bench_table1_g$g.err
python - <<PY
import json
d=json.load(open('$o/bench_table1_g$g.json')); print('table1 graph $g', d['ms_per_step'], d['cuda_graph'], d['stage_source'], d['e2e']['ms_per_step'], d['gpu_launches'])
PY
done
for g in 0 1; do
python bench.py --workload grid512 --steps 10 --warmup 3 --no-cpu-baseline --graph $g > $o/bench_grid512_g$g.json 2> $o/bench_grid512_g$g.err
python - <<PY
import json
d=json.load(open('$o/bench_grid512_g$g.json')); print('grid512 graph $g', d['ms_per_step'], d['cuda_graph'], d['stage_source'], d['e2e']['ms_per_step'], d['gpu_launches'])
PY
done
