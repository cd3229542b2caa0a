set -x
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -x -q -k "synth_models_and_batch or calibrated_per_model or streaming_pair or fast_tier" > gpurun_out/r2c_quick.log 2>&1; echo quick rc=$?
tail -5 gpurun_out/r2c_quick.log
timeout -k 10 600 python tools/tc_pipeline_probe.py > gpurun_out/r2c_probe.log 2>&1; echo probe rc=$?
grep "^(" gpurun_out/r2c_probe.log
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --tc-variant 2 > gpurun_out/r2c_bench_tc2.json 2> gpurun_out/r2c_bench_tc2.err; echo bench2 rc=$?
python -c "
import json
for f in ('tc2',):
    d=json.load(open('gpurun_out/r2c_bench_%s.json'%f)); print(f, d['ms_per_step'], d['stage_ms_per_step'])
"
