set -x
mkdir -p gpurun_out/r2p
o=gpurun_out/r2p
for f in 0 64; do
HAF_TC_DEBUG=$f timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_dbg$f.json 2> $o/bench_dbg$f.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2p/bench_dbg$f.json')); print('flags $f', d['ms_per_step'], d['stage_ms_per_step']['svm'], d['roofline']['kernel_ms'], d['roofline']['frac'])
PY
done
