set -x
o=gpurun_out/r2ab
mkdir -p $o
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "several_streams or batch_matches_single or full_size_batch or cuda_graph" > $o/tests.log 2>&1; echo tests rc=$?
tail -5 $o/tests.log
run() { tag=$1; shift; env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_$tag.json 2> $o/bench_$tag.err; echo $tag rc=$?; python - <<PY
import json
try:
    d=json.load(open('$o/bench_$tag.json')); e=d['e2e']; print('$tag', 'dev', round(d['ms_per_step'],3), 'e2e', round(e['ms_per_step'],3), 'pageable', round(e['pageable_ms_per_step'],3))
except Exception as ex: print('$tag', 'failed', ex)
PY
}
run t4 HAF_STAGE_THREADS=4
run t8 HAF_STAGE_THREADS=8
run t2 HAF_STAGE_THREADS=2
run t12 HAF_STAGE_THREADS=12
run t0 HAF_STAGE_THREADS=0
nproc
