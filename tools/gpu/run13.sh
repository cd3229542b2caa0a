set -x
mkdir -p gpurun_out/r2m
o=gpurun_out/r2m
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_n1.json 2> $o/bench_n1.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m/bench_n1.json')); print(d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['audit']['max_rel_error_of_E'], d['guard_windows_per_step'])
PY
HAF_X_BUDGET_GIB=8 timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_n1_onechunk.json 2> /dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m/bench_n1_onechunk.json')); print('one chunk', d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['ms_per_step'])
PY
for gk in 0 1; do timeout -k 10 200 python bench.py --workload table1 --steps 20 --warmup 5 --no-cpu-baseline --guard-kernel $gk > $o/bench_table1_gk$gk.json 2> /dev/null; python -c "
import json; d=json.load(open('$o/bench_table1_gk$gk.json')); print('table1 gk$gk', d['ms_per_step'], d['stage_ms_per_step'])"; done
timeout -k 10 300 python tools/dec_error_probe.py > $o/dec_error_probe.txt 2>&1; cat $o/dec_error_probe.txt
timeout -k 10 1500 python -m pytest tests -m gpu -q --durations=5 > $o/tests.log 2>&1; echo tests rc=$?
tail -12 $o/tests.log
