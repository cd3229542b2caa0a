#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel of the library runs at least once.
  compute-sanitizer --tool memcheck python tools/sanitizer_probe.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import haf_grasping_b200 as h  # noqa: E402
from haf_grasping_b200 import synth  # noqa: E402

F = os.path.join(ROOT, "tests", "golden", "refdata", "Features.txt")
R = os.path.join(ROOT, "tests", "golden", "refdata", "range21062012_allfeatures")


def main():
    tmp = tempfile.mkdtemp()
    model = synth.write_synth_model(os.path.join(tmp, "s.model"), 300, rho=-0.2972253)
    clouds = [synth.synth_cloud(77 + i, 6000 + 500 * i) for i in range(3)]
    for mode in (h.HAF_SVM_TENSOR_GUARD, h.HAF_SVM_FP32_GUARD, h.HAF_SVM_FP64_EXACT):
        for tier2 in ((0, 2) if mode != h.HAF_SVM_FP64_EXACT else (0,)):
            g = h.GraspSearch(F, R, model, svm_mode=mode, guard_rel=1e-3 if mode != h.HAF_SVM_FP64_EXACT else 0.0, guard_tier2=tier2)
            res = g.search(clouds[0])
            best = g.search_batch(clouds)
            t = g.timing()
            print("mode", mode, "tier2", tier2, "best", res["best"].astuple(), [b.astuple() for b in best], "guard", t.n_guard, "exact", t.n_exact)
            g.close()
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, size=(700, 330)) * (rng.random((700, 330)) < 0.8)
    for mode in (0, 2, 1):
        p = h.SvmPredictor(model, svm_mode=mode, min_dims=330, guard_rel=1e-3 if mode != 1 else 0.0)
        lab, dec = p.predict(x)
        print("svm mode", mode, "labels+", int((lab > 0).sum()), "guard", p.timing().n_guard)
        p.close()
    fmin, fmax = h.scale_minmax(x, 330)
    out = h.scale_apply(x, 330, fmin, fmax)
    print("scale", float(out.min()), float(out.max()))


if __name__ == "__main__":
    main()
