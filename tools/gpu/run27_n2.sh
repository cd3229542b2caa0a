set -x
mkdir -p gpurun_out/r2n2
o=gpurun_out/r2n2
nvidia-smi -L
timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=5 > $o/tests_multi.log 2>&1; echo tests rc=$?
tail -15 $o/tests_multi.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $o/bench_n2.json 2> $o/bench_n2.err; echo n2 rc=$?
timeout -k 10 300 python bench.py --group 2 --steps 10 --warmup 3 --no-cpu-baseline > $o/bench_group2.json 2> $o/bench_group2.err; echo g2 rc=$?
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload grid512 --steps 10 --warmup 3 > $o/bench_grid512_n2.json 2> $o/bench_grid512_n2.err; echo grid n2 rc=$?
python - <<'PY'
import json
for f in ('bench_n2','bench_group2','bench_grid512_n2'):
    try:
        d=json.load(open('gpurun_out/r2n2/%s.json'%f)); print(f, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'])
    except Exception as e: print(f, 'ERR', e)
PY
for f in $o/*.err; do echo == $f; tail -n 5 $f; done
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload approach --steps 20 --warmup 5 > $o/bench_approach_n2.json 2> $o/bench_approach_n2.err; echo approach n2 rc=$?
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $o/bench_reference_n2.json 2> $o/bench_reference_n2.err; echo ref n2 rc=$?
cat $o/bench_approach_n2.json | cut -c1-400
