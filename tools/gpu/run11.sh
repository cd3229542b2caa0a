set -x
mkdir -p gpurun_out/r2k
o=gpurun_out/r2k
timeout -k 10 900 python -m pytest tests/test_pcd_device.py tests/test_gpu_parity.py -m gpu -q --durations=5 -k "pcd or guard or tier or cli" > $o/tests.log 2>&1; echo tests rc=$?
tail -30 $o/tests.log
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_n1.json 2> $o/bench_n1.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k/bench_n1.json')); print(d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['wall_ms_per_step'], d['audit'])
PY
timeout -k 10 600 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "bench_batch or far_from or audit" > $o/tests2.log 2>&1; echo tests2 rc=$?
tail -5 $o/tests2.log
