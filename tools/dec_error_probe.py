#!/usr/bin/env python
"""Measures the decision-value error of the FP32 SIMT and tensor-core SVM paths against the FP64 exact-order path
on the GPU (same windows), relative to sum_i |coef_i| -- the evidence behind the guard-band widths."""
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


def run(model, xyz, mode):
    g = h.GraspSearch(F, R, model, svm_mode=mode)
    res = g.search(xyz)
    win = g.debug_windows()
    dec, lab, guard = g.debug_decisions()
    order = np.lexsort((win[:, 1], win[:, 0]))
    t = g.timing()
    g.close()
    return dec[order], lab[order], guard[order], res["best"].astuple(), t


def main():
    import gzip
    tmp = tempfile.mkdtemp()
    models = {"synth2048": synth.write_synth_model(os.path.join(tmp, "s2048.model"), 2048)}
    tm = os.path.join(tmp, "trained.model")
    with gzip.open(os.path.join(ROOT, "tests", "golden", "substitute_trained.model.gz"), "rb") as src, open(tm, "wb") as dst:
        dst.write(src.read())
    models["trained"] = tm
    clouds = {"synth100k": synth.synth_cloud(1234, 100000), "table1": np.load(os.path.join(ROOT, "tests", "golden", "clouds.npz"))["table1"]}
    for mn, mp in models.items():
        coefs = []
        with open(mp) as fh:
            sv = False
            for ln in fh:
                if sv and ln.strip():
                    coefs.append(abs(float(ln.split()[0])))
                elif ln.startswith("SV"):
                    sv = True
        scale = sum(coefs)
        for cn, xyz in clouds.items():
            d64, l64, _, b64, _ = run(mp, xyz, h.HAF_SVM_FP64_EXACT)
            for mode, name in ((h.HAF_SVM_FP32_GUARD, "simt"), (h.HAF_SVM_TENSOR_GUARD, "tensor")):
                d, lab, guard, b, t = run(mp, xyz, mode)
                ng = ~guard.astype(bool)
                err = np.abs(d - d64)
                print("%-10s %-10s %-7s W=%d max|err|=%.3e (%.2e of sum|coef|=%.1f) rms=%.2e guard=%d labels_equal=%s best_equal=%s min|dec| outside guard=%.3e device_ms=%.3f"
                      % (mn, cn, name, len(d), err[ng].max(), err[ng].max() / scale, scale, np.sqrt((err[ng] ** 2).mean()), int(guard.sum()),
                         bool((lab == l64).all()), b == b64, np.abs(d[ng]).min(), t.ms_total))


if __name__ == "__main__":
    main()
