set -x
mkdir -p gpurun_out/r2s
o=gpurun_out/r2s
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_n1.json 2> $o/bench_n1.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s/bench_n1.json')); print(d['ms_per_step'], d['stage_ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
PY
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor or bundled or fast_tier or large_grid" > $o/tests.log 2>&1; echo tests rc=$?
tail -4 $o/tests.log
