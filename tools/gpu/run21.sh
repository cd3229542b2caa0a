set -x
mkdir -p gpurun_out/r2u
o=gpurun_out/r2u
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cuda_graph" > $o/tests_graph.log 2>&1; echo graph tests rc=$?
tail -15 $o/tests_graph.log
for ng in 0 1; do
HAF_NO_GRAPH=$ng timeout -k 10 200 python bench.py --workload table1 --steps 50 --warmup 5 --no-cpu-baseline > $o/bench_table1_nograph$ng.json 2> /dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/r2u/bench_table1_nograph$ng.json')); print('table1 HAF_NO_GRAPH=$ng', d['ms_per_step'], d['wall_ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])
PY
done
timeout -k 10 200 python bench.py --workload approach --steps 30 --warmup 5 > $o/bench_approach_n1.json 2> /dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2u/bench_approach_n1.json')); print('approach n1', d['ms_per_step'], d['e2e']['ms_per_step'])"
timeout -k 10 1500 python -m pytest tests -m gpu -q -x --durations=5 > $o/tests.log 2>&1; echo tests rc=$?
tail -8 $o/tests.log
