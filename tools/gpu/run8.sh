set -x
mkdir -p gpurun_out/r2h
o=gpurun_out/r2h
timeout -k 10 900 python -m pytest tests/test_probability.py tests/test_gpu_scale.py tests/test_svm_frontends.py -m gpu -q --durations=8 -k "probability or far_from or audit or chunking or b1" > $o/tests.log 2>&1; echo tests rc=$?
tail -40 $o/tests.log
