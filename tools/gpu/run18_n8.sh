set -x
N=${1:-8}
mkdir -p gpurun_out/r2r
o=gpurun_out/r2r
nvidia-smi -L | wc -l
nvidia-smi topo -m > $o/topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 10 300 $TR --master-port 29521 bench.py --gpus $N --workload approach --steps 20 --warmup 5 > $o/bench_approach_n$N.json 2> $o/bench_approach_n$N.err; echo approach rc=$?
timeout -k 10 300 $TR --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > $o/bench_n$N.json 2> $o/bench_n$N.err; echo batch rc=$?
timeout -k 10 300 $TR --master-port 29523 bench.py --gpus $N --workload grid512 --steps 10 --warmup 3 > $o/bench_grid512_n$N.json 2> $o/bench_grid512_n$N.err; echo grid rc=$?
timeout -k 10 300 python bench.py --group $N --steps 5 --warmup 3 --no-cpu-baseline > $o/bench_group$N.json 2> $o/bench_group$N.err; echo group rc=$?
python - <<PY
import json
for f in ('bench_approach_n$N','bench_n$N','bench_grid512_n$N','bench_group$N'):
    try:
        d=json.load(open('gpurun_out/r2r/%s.json'%f)); print(f, d['n_gpus'], '%.4g'%d['value'], 'ms %.4g'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'e2e ms %.4g'%d['e2e']['ms_per_step'], d['e2e'].get('h2d_only_ms_per_step'), d['config'].get('units_per_rank'))
    except Exception as e: print(f, 'ERR', e)
PY
for f in $o/*n$N.err $o/*group$N.err; do echo == $f; tail -n 3 $f; done
