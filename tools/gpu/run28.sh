set -x
o=gpurun_out/r2aa
mkdir -p $o
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "several_streams or cuda_graph" > $o/tests.log 2>&1; echo tests rc=$?
tail -5 $o/tests.log
python bench.py --workload table1 --steps 50 --warmup 5 --no-cpu-baseline > $o/bench_table1.json 2> $o/bench_table1.err
python bench.py --workload approach --steps 20 --warmup 5 --no-cpu-baseline > $o/bench_approach.json 2> $o/bench_approach.err
python bench.py --workload approach --steps 20 --warmup 5 --no-cpu-baseline --graph 0 > $o/bench_approach_g0.json 2> $o/bench_approach_g0.err
python - <<'PY'
import json
for f in ('bench_table1','bench_approach','bench_approach_g0'):
    d=json.load(open('gpurun_out/r2aa/%s.json'%f)); print(f, d['ms_per_step'], d.get('cuda_graph'), d['e2e']['ms_per_step'])
PY
