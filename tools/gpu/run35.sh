o=gpurun_out/r2ad
mkdir -p $o
timeout -k 5 170 compute-sanitizer --tool memcheck python tools/sanitizer_probe.py staged > $o/sanitizer_staged.log 2>&1; echo rc=$?
grep -v "^=========$" $o/sanitizer_staged.log | tail -8
