set -x
mkdir -p gpurun_out
timeout -k 10 2400 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r2e_tests.log 2>&1; echo tests rc=$?
tail -25 gpurun_out/r2e_tests.log
