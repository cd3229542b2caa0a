#!/bin/bash
# One gpurun call: the measurement set that backs profiles/ (bench lines, ncu launch list, full captures, probes).
# usage (from the repo root, on the GPU box):  bash tools/final_measure.sh <tag>
tag=${1:-r2}
out=gpurun_out/$tag
mkdir -p $out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/smi.txt
python bench.py --steps 10 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --workload table1 --steps 20 --warmup 5 > $out/bench_table1.json 2> /dev/null
python bench.py --workload table1 --steps 50 --warmup 5 --no-cpu-baseline --graph 0 > $out/bench_table1_nograph.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 36 --csv --log-file $out/launches_table1.csv python bench.py --workload table1 --steps 2 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
python bench.py --workload grid512 --steps 5 --warmup 3 > $out/bench_grid512.json 2> /dev/null
python bench.py --workload approach --steps 20 --warmup 5 > $out/bench_approach_n1.json 2> /dev/null
python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_reference.json 2> /dev/null
# launch list of the two timed steps of a --steps 2 --warmup 3 run: every kernel of the timed region, nothing else
# (3 warm-up steps of 18 launches each are skipped; 2 x 18 are kept -- counted with `gpu_launches` of the bench line)
ncu --metrics gpu__time_duration.sum --clock-control none -s 54 -c 36 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'svm_rbf_tc|features_tc_kernel|bin_maxz_cloud_kernel|guard_dmma_kernel|guard_inputs_kernel' -s 15 -c 5 -o $out/prof_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_full.log 2>&1
# every launch of a whole bench run incl. the host-staged (chunked) steps: per-chunk kernel durations
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_e2e.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/dec_error_probe.py > $out/dec_error_probe.txt 2>&1
python tools/svm_cli_bench.py > $out/svm_cli_bench.json 2> $out/svm_cli_bench.err
compute-sanitizer --tool memcheck python tools/sanitizer_probe.py > $out/sanitizer_memcheck.log 2>&1
tail -3 $out/sanitizer_memcheck.log
python -m pytest tests -m gpu -q --durations=8 > $out/tests.log 2>&1; echo tests rc=$?
tail -6 $out/tests.log
cat $out/bench_n1.json
