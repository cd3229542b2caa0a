set -x
o=gpurun_out/r2ac
mkdir -p $o
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast_tier or test_tensor_core_mode or bundled_pcd_trained or several_streams or full_size or calibrated" --deselect "tests/test_gpu_parity.py::test_tensor_core_mode_synth_models_and_batch" > $o/tests.log 2>&1; echo tests rc=$?
tail -5 $o/tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_n1.json 2> $o/bench_n1.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ac/bench_n1.json')); e=d['e2e']
print('n1', d['value'], d['ms_per_step'], 'e2e', e['ms_per_step'], 'pageable', e['pageable_ms_per_step'], d['stage_ms_per_step'], d['audit'])
PY
