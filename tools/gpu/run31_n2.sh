set -x
o=gpurun_out/r2n3
mkdir -p $o
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload approach --steps 20 --warmup 5 > $o/bench_approach_n2.json 2> $o/bench_approach_n2.err; echo approach n2 rc=$?
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --workload approach --steps 20 --warmup 5 --graph 0 > $o/bench_approach_n2_g0.json 2> $o/bench_approach_n2_g0.err; echo approach n2 g0 rc=$?
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $o/bench_n2.json 2> $o/bench_n2.err; echo n2 rc=$?
python - <<'PY'
import json
o='gpurun_out/r2n3/'
for f in ('bench_approach_n2','bench_approach_n2_g0','bench_n2'):
    try:
        d=json.load(open(o+f+'.json')); e=d['e2e']; print(f, d['n_gpus'], d['value'], d['ms_per_step'], e['ms_per_step'], e.get('pageable_ms_per_step'), d.get('cuda_graph'))
    except Exception as ex: print(f,'ERR',ex)
PY
