set -x
mkdir -p gpurun_out/r2q
o=gpurun_out/r2q
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_n1.json 2> $o/bench_n1.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2q/bench_n1.json')); print(d['ms_per_step'], d['stage_ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
PY
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x -k "cell_indices or bundled or batch or approach or large_grid or full_size or padded or config4 or request" > $o/tests.log 2>&1; echo tests rc=$?
tail -4 $o/tests.log
