set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout -k 10 240 python -m pytest tests/test_gpu_parity.py -x -q -k "synth_models_and_batch or calibrated_per_model" > gpurun_out/r2a_quick.log 2>&1; echo quick rc=$?
tail -5 gpurun_out/r2a_quick.log
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo smoke rc=$?
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_tc3.json 2> gpurun_out/r2a_bench_tc3.err; echo bench rc=$?
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --tc-variant 2 > gpurun_out/r2a_bench_tc2.json 2> gpurun_out/r2a_bench_tc2.err; echo bench2 rc=$?
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --bin-variant 2 > gpurun_out/r2a_bench_bin1.json 2> gpurun_out/r2a_bench_bin1.err; echo bench3 rc=$?
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo tests rc=$?
tail -15 gpurun_out/r2a_tests.log
