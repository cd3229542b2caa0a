set -x
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -x -q -k "synth_models_and_batch or calibrated_per_model or streaming_pair" > gpurun_out/r2d_quick.log 2>&1; echo quick rc=$?
tail -5 gpurun_out/r2d_quick.log
timeout -k 10 200 python tools/tc_cycle_probe.py 32 > gpurun_out/r2d_cycles.log 2>&1; tail -17 gpurun_out/r2d_cycles.log
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench.json')); print(d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['ms_per_step'])
PY
