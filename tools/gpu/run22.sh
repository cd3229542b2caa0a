set -x
o=gpurun_out/r2v
mkdir -p $o
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_streams or cuda_graph or batch_matches_single or full_size_batch" > $o/tests.log 2>&1; echo tests rc=$?
tail -15 $o/tests.log
run() { tag=$1; shift; env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_$tag.json 2> $o/bench_$tag.err; echo $tag rc=$?; python - <<PY
import json
try:
    d=json.load(open('$o/bench_$tag.json')); e=d['e2e']; print('$tag', 'dev', round(d['ms_per_step'],3), 'e2e', round(e['ms_per_step'],3), 'h2d', round(e['h2d_only_ms_per_step'],3), 'chunks', e['chunks_per_step'], 'stages', {k: round(v,2) for k,v in e['stage_ms_per_step'].items()})
except Exception as ex: print('$tag', 'failed', ex)
PY
}
run dual HAF_X=0
run single HAF_DUAL_STREAM=0
run dual_taper HAF_STAGE_SCHED=16,24,36,54,64,64,64,64,48,40,24,14
run dual_big HAF_STAGE_SCHED=16,32,48,64,96,96,96,40,24
run dual_small HAF_STAGE_SCHED=16,16,24,32,32,40,48,48,48,48,48,48,48,16
run dual_first8 HAF_STAGE_SCHED=8,16,24,36,54,64,64,64,64,64,32,22
