set -x
mkdir -p gpurun_out/r2t
o=gpurun_out/r2t
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_n1.json 2> $o/bench_n1.err; echo bench rc=$?
timeout -k 10 200 python bench.py --workload table1 --steps 20 --warmup 5 --no-cpu-baseline > $o/bench_table1.json 2> /dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2t/bench_n1.json')); print(d['ms_per_step'], d['stage_ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['stage_ms_per_step'], d['gpu_launches'])
d=json.load(open('gpurun_out/r2t/bench_table1.json')); print('table1', d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])
PY
timeout -k 10 1200 python -m pytest tests/test_gpu_parity.py tests/test_svm_frontends.py tests/test_probability.py -m gpu -q -x -k "not synth_models_and_batch" > $o/tests.log 2>&1; echo tests rc=$?
tail -4 $o/tests.log
