set -x
o=gpurun_out/r2w
mkdir -p $o
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_streams or cuda_graph" > $o/tests.log 2>&1; echo tests rc=$?
tail -5 $o/tests.log
run() { tag=$1; shift; env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_$tag.json 2> $o/bench_$tag.err; echo $tag rc=$?; python - <<PY
import json
try:
    d=json.load(open('$o/bench_$tag.json')); e=d['e2e']; print('$tag', 'dev', round(d['ms_per_step'],3), 'e2e', round(e['ms_per_step'],3), 'h2d', round(e['h2d_only_ms_per_step'],3), 'chunks', e['chunks_per_step'], 'stages', {k: round(v,2) for k,v in e['stage_ms_per_step'].items()})
except Exception as ex: print('$tag', 'failed', ex)
PY
}
run s2 HAF_DUAL_STREAM=2
run s3 HAF_DUAL_STREAM=3
run s3_small HAF_DUAL_STREAM=3 HAF_STAGE_SCHED=16,16,24,32,32,40,48,48,48,48,48,48,48,16
run s2_40 HAF_DUAL_STREAM=2 HAF_STAGE_SCHED=16,16,24,32,40,40,40,40,40,40,40,40,40,40,8
run s2_32 HAF_DUAL_STREAM=2 HAF_STAGE_SCHED=16,16,24,32
run s4_small HAF_DUAL_STREAM=4 HAF_STAGE_SCHED=16,16,24,32,32,40,48,48,48,48,48,48,48,16
run s3_32 HAF_DUAL_STREAM=3 HAF_STAGE_SCHED=16,16,24,32
HAF_DUAL_STREAM=0 timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $o/launches_e2e.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $o/ncu_e2e.log 2>&1
