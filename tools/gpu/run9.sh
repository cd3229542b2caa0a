set -x
mkdir -p gpurun_out/r2i
o=gpurun_out/r2i
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_n1.json 2> $o/bench_n1.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i/bench_n1.json')); print(d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['wall_ms_per_step'])
PY
timeout -k 10 1500 python -m pytest tests -m gpu -q --durations=5 -x > $o/tests.log 2>&1; echo tests rc=$?
tail -30 $o/tests.log
