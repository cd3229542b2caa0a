#!/usr/bin/env python
"""Where do the role threads of svm_rbf_tc3_kernel spend their cycles?  (HAF_TC_DEBUG=32: clock64 around every wait.)
Runs one 256-cloud bench-shaped batch and prints per-role averages over the CTAs.  GPU box only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 32
os.environ["HAF_TC_DEBUG"] = str(flags)
import torch  # noqa: E402

import bench  # noqa: E402
import haf_grasping_b200 as h  # noqa: E402
from haf_grasping_b200 import synth  # noqa: E402

model = bench.model_path(2048)
n = 256
clouds = [synth.synth_cloud(1234 + i, 100000) for i in range(n)]
off = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int64)
dev = torch.from_numpy(np.concatenate(clouds)).cuda()
gs = h.GraspSearch(bench.FEATURES, bench.RANGE, model)
gs.set_profiling(True)
for _ in range(3):
    gs.search_batch_packed(dev, off)
t = gs.timing()
p = gs.debug_tc_probe(148).astype(np.float64)
lead, peer = p[0::2], p[1::2]
names = ["prod wait x_empty", "prod wait s_empty", "prod issue", "prod loads", "mma wait t_empty", "mma wait x_full", "mma wait s_full",
         "mma issue+commit", "mma k-blocks", "epi(w2) wait t_full", "epi(w2) work", "epi tiles", "cta cycles"]
print("flags", flags, "svm ms", t.ms_svm, "windows", t.n_windows)
for k, nm in enumerate(names):
    print("%-22s leader %12.0f   peer %12.0f" % (nm, lead[:, k].mean(), peer[:, k].mean()))
kb = lead[:, 8].mean()
print("per k-block (leader): wait s_full %.0f  issue+commit %.0f   per tile: wait t_empty %.0f   cta cycles per k-block %.0f" % (
    lead[:, 6].mean() / kb, lead[:, 7].mean() / kb, lead[:, 4].mean() / (kb / 6), lead[:, 12].mean() / kb))
