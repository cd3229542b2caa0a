set -x
mkdir -p gpurun_out/r2l
o=gpurun_out/r2l
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_n1.json 2> $o/bench_n1.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l/bench_n1.json')); print(d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['wall_ms_per_step'], d['audit'], d['e2e'].get('h2d_only_ms_per_step'), d['e2e'].get('pageable_ms_per_step'))
PY
timeout -k 10 300 python tools/dec_error_probe.py > $o/dec_error_probe.txt 2>&1; cat $o/dec_error_probe.txt
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x -k "tensor or bench_batch or far_from or audit or calibrated" > $o/tests.log 2>&1; echo tests rc=$?
tail -5 $o/tests.log
