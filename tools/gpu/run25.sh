set -x
o=gpurun_out/r2y
mkdir -p $o
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "several_streams or cuda_graph or batch_matches_single" > $o/tests.log 2>&1; echo tests rc=$?
tail -5 $o/tests.log
run() { tag=$1; shift; env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_$tag.json 2> $o/bench_$tag.err; echo $tag rc=$?; python - <<PY
import json
try:
    d=json.load(open('$o/bench_$tag.json')); e=d['e2e']; print('$tag', 'dev', round(d['ms_per_step'],3), 'e2e', round(e['ms_per_step'],3), 'h2d', round(e['h2d_only_ms_per_step'],3), 'chunks', e['chunks_per_step'], e['stage_note'][:60], 'traffic', d['roofline']['traffic'], d['roofline']['frac'])
except Exception as ex: print('$tag', 'failed', ex)
PY
}
run default HAF_X=0
run s2 HAF_DUAL_STREAM=2
run s4 HAF_DUAL_STREAM=4
run s1 HAF_DUAL_STREAM=0
run s3_64 HAF_STAGE_SCHED=16,24,32,48,64
run s3_taper HAF_STAGE_SCHED=16,24,32,48,56,56,56,56,56,48,32,24,8
