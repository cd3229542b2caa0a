set -x
mkdir -p gpurun_out/r2j
o=gpurun_out/r2j
timeout -k 10 900 ncu --set full --import-source on --clock-control none -k regex:"features_tc|guard_fma" -s 12 -c 4 \
    -o $o/prof_fg -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $o/prof_fg.log 2>&1; echo ncu1 rc=$?
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "guard or tier or large_grid or approach or batch_matches or request or padded or roll_begin or full_size" > $o/tests.log 2>&1; echo tests rc=$?
tail -5 $o/tests.log
