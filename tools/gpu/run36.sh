o=gpurun_out/r2ae
mkdir -p $o
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 5 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "several_streams or cuda_graph or batch_matches_single or (bundled_pcd_trained and pcd2)" 2>&1 | tail -3
