set -x
o=gpurun_out/r2x
mkdir -p $o
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_streams or cuda_graph" > $o/tests.log 2>&1; echo tests rc=$?
tail -5 $o/tests.log
run() { tag=$1; shift; env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_$tag.json 2> $o/bench_$tag.err; echo $tag rc=$?; python - <<PY
import json
try:
    d=json.load(open('$o/bench_$tag.json')); e=d['e2e']; print('$tag', 'dev', round(d['ms_per_step'],3), {k: round(v,2) for k,v in d['stage_ms_per_step'].items()}, 'e2e', round(e['ms_per_step'],3), 'h2d', round(e['h2d_only_ms_per_step'],3), 'chunks', e['chunks_per_step'], 'stages', {k: round(v,2) for k,v in e['stage_ms_per_step'].items()})
except Exception as ex: print('$tag', 'failed', ex)
PY
}
run s3 HAF_DUAL_STREAM=3
run s3_small HAF_DUAL_STREAM=3 HAF_STAGE_SCHED=16,16,24,32,32,40,48,48,48,48,48,48,48,16
run s3_mid HAF_DUAL_STREAM=3 HAF_STAGE_SCHED=16,24,32,48,56,56,56,56,56,56,56
run s2_small HAF_DUAL_STREAM=2 HAF_STAGE_SCHED=16,16,24,32,32,40,48,48,48,48,48,48,48,16
run s3_small2 HAF_DUAL_STREAM=3 HAF_STAGE_SCHED=16,24,36,48,48,48,48,48,48,48,48,32,20
run s3_small_r2 HAF_DUAL_STREAM=3 HAF_STAGE_SCHED=16,16,24,32,32,40,48,48,48,48,48,48,48,16 HAF_RESIDENT_SPLIT=2
run s3_small_r3 HAF_DUAL_STREAM=3 HAF_STAGE_SCHED=16,16,24,32,32,40,48,48,48,48,48,48,48,16 HAF_RESIDENT_SPLIT=3
run s2_r4 HAF_DUAL_STREAM=2 HAF_RESIDENT_SPLIT=4
