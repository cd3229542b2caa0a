set -x
mkdir -p gpurun_out
# full captures of the four top kernels, one bench step (2 chunks) after the warm-up launches; default = X-resident tc3
timeout -k 10 900 ncu --set full --import-source on --clock-control none -k regex:"svm_rbf_tc|features_tc|bin_maxz_cloud|guard_fma" -s 24 -c 8 \
    -o gpurun_out/r2_prof_tc3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_prof_tc3.log 2>&1; echo ncu1 rc=$?
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:"svm_rbf_tc" -s 6 -c 2 \
    -o gpurun_out/r2_prof_tc2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --tc-variant 2 > gpurun_out/r2_prof_tc2.log 2>&1; echo ncu2 rc=$?
# launch list of one timed step of the default bench
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 200 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_launches.log 2>&1; echo ncu3 rc=$?
ls -la gpurun_out
