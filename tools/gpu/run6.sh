set -x
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests/test_gpu_scale.py tests/test_svm_frontends.py -m gpu -q -x --durations=8 -k "not config4" > gpurun_out/r2f_tests.log 2>&1; echo tests rc=$?
tail -25 gpurun_out/r2f_tests.log
