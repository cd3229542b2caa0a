set -x
o=gpurun_out/r2_final4
mkdir -p $o
python -c "import __graft_entry__ as g; g.smoke()" > $o/smoke.log 2>&1; echo smoke rc=$?; tail -2 $o/smoke.log
python bench.py --steps 10 --warmup 3 > $o/bench_n1.json 2> $o/bench_n1.err; echo bench rc=$?
python bench.py --workload table1 --steps 50 --warmup 5 --no-cpu-baseline > $o/bench_table1.json 2> /dev/null
python bench.py --workload approach --steps 20 --warmup 5 --no-cpu-baseline > $o/bench_approach_n1.json 2> /dev/null
python bench.py --workload grid512 --steps 5 --warmup 3 --no-cpu-baseline > $o/bench_grid512.json 2> /dev/null
python - <<'PY'
import json
o='gpurun_out/r2_final4/'
d=json.load(open(o+'bench_n1.json')); e=d['e2e']
print('n1', d['value'], d['ms_per_step'], 'e2e', e['value'], e['ms_per_step'], e['frac_of_h2d_ceiling'], 'pageable', e['pageable_ms_per_step'], 'roof', d['roofline']['frac'], d['roofline']['traffic'], d['stage_ms_per_step'])
for f in ('bench_table1','bench_approach_n1','bench_grid512'):
    d=json.load(open(o+f+'.json')); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d.get('cuda_graph'))
PY
python -m pytest tests -m gpu -q --durations=5 > $o/tests.log 2>&1; echo tests rc=$?
tail -4 $o/tests.log
