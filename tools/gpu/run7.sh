set -x
mkdir -p gpurun_out/r2g
o=gpurun_out/r2g
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout -k 10 1500 python -m pytest tests -m gpu -q --durations=12 > $o/tests.log 2>&1; echo tests rc=$?
tail -25 $o/tests.log
timeout -k 10 400 python bench.py --steps 10 --warmup 3 > $o/bench_n1.json 2> $o/bench_n1.err; echo bench rc=$?
timeout -k 10 200 python bench.py --workload table1 --steps 20 --warmup 5 --no-cpu-baseline > $o/bench_table1.json 2> /dev/null
timeout -k 10 300 python bench.py --workload grid512 --steps 5 --warmup 3 --no-cpu-baseline > $o/bench_grid512.json 2> /dev/null
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 200 --csv --log-file $o/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/launches.log 2>&1; echo ncu3 rc=$?
timeout -k 10 900 ncu --set full --import-source on --clock-control none -k regex:"svm_rbf_tc|features_tc|bin_maxz_cloud|guard_fma|guard_inputs" -s 30 -c 10 \
    -o $o/prof_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $o/prof_full.log 2>&1; echo ncu1 rc=$?
timeout -k 10 400 python tools/dec_error_probe.py > $o/dec_error_probe.txt 2>&1; echo probe rc=$?
cat $o/bench_n1.json
ls -la $o
