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


def run(model, xyz, mode, want_scaled=False):
    g = h.GraspSearch(F, R, model, svm_mode=mode)
    res = g.search(xyz)
    win = g.debug_windows()
    dec, lab, guard = g.debug_decisions()
    order = np.lexsort((win[:, 1], win[:, 0]))
    t = g.timing()
    scaled = g.debug_features(raw=False)[1][order] if want_scaled else None
    g.close()
    return dec[order], lab[order], guard[order], res["best"].astuple(), t, scaled


def load_model(path):
    """libsvm text model -> (gamma, rho, coef [S], sv [S][D]) dense float64"""
    gamma = rho = 0.0
    coefs, rows, dmax = [], [], 0
    with open(path) as fh:
        sv = False
        for ln in fh:
            if sv:
                t = ln.split()
                if not t:
                    continue
                coefs.append(float(t[0]))
                r = [(int(a.split(":")[0]), float(a.split(":")[1])) for a in t[1:]]
                rows.append(r)
                dmax = max([dmax] + [i for i, _ in r])
            elif ln.startswith("gamma"):
                gamma = float(ln.split()[1])
            elif ln.startswith("rho"):
                rho = float(ln.split()[1])
            elif ln.startswith("SV"):
                sv = True
    m = np.zeros((len(rows), dmax))
    for k, r in enumerate(rows):
        for i, v in r:
            m[k, i - 1] = v
    return gamma, rho, np.array(coefs), m


def guard_scale(model, scaled):
    """E + |rho| per window in float64, E = sum_i |coef_i| K_i (1 + gamma log2(e) (|x|^2 + |sv_i|^2)): the quantity the
    guard band is a fraction of (svm_tc.cuh, GUARD SCALE)"""
    gamma, rho, coef, sv = model
    D = max(sv.shape[1], scaled.shape[1])
    x = np.zeros((len(scaled), D)); x[:, :scaled.shape[1]] = scaled
    s = np.zeros((len(sv), D)); s[:, :sv.shape[1]] = sv
    base = (x * x).sum(1)[:, None] + (s * s).sum(1)[None, :]
    d2 = base - 2.0 * x @ s.T
    return (np.exp(-gamma * np.maximum(d2, 0.0)) * (1.0 + gamma * 1.4426950408889634 * base)) @ np.abs(coef) + abs(rho)


def main():
    import gzip
    tmp = tempfile.mkdtemp()
    models = {"synth2048": synth.write_synth_model(os.path.join(tmp, "s2048.model"), 2048)}
    tm = os.path.join(tmp, "trained.model")
    with gzip.open(os.path.join(ROOT, "tests", "golden", "substitute_trained.model.gz"), "rb") as src, open(tm, "wb") as dst:
        dst.write(src.read())
    models["trained"] = tm
    clouds = {"synth100k": synth.synth_cloud(1234, 100000), "table1": np.load(os.path.join(ROOT, "tests", "golden", "clouds.npz"))["table1"]}
    only = sys.argv[1:]   # optional model names to restrict the run to
    for mn, mp in models.items():
        if only and mn not in only:
            continue
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
            d64, l64, _, b64, _, scaled = run(mp, xyz, h.HAF_SVM_FP64_EXACT, want_scaled=True)
            gs = guard_scale(load_model(mp), scaled)
            for mode, name in ((h.HAF_SVM_FP32_GUARD, "simt"), (h.HAF_SVM_TENSOR_GUARD, "tensor")):
                d, lab, guard, b, t, _ = run(mp, xyz, mode)
                ng = ~guard.astype(bool)
                err = np.abs(d - d64)
                print("%-10s %-10s %-7s W=%d max|err|=%.3e (%.2e of sum|coef|=%.1f; max err/(E+|rho|)=%.2e, rms %.2e) guard=%d labels_equal=%s best_equal=%s min|dec| outside guard=%.3e device_ms=%.3f"
                      % (mn, cn, name, len(d), err[ng].max(), err[ng].max() / scale, scale, (err[ng] / gs[ng]).max(), np.sqrt(((err[ng] / gs[ng]) ** 2).mean()),
                         int(guard.sum()), bool((lab == l64).all()), b == b64, np.abs(d[ng]).min(), t.ms_total))
                # distributions over the windows OUTSIDE the guard band (the ones that keep the contraction's value):
                # err / (E + |rho|), the unit the band is stated in, and err / |dec|, the plain relative error
                def hist(v, edges):
                    c, _ = np.histogram(v, bins=edges)
                    return "  ".join("<%g:%d" % (e, n) for e, n in zip(edges[1:], c))
                rel_e = err[ng] / gs[ng]
                rel_d = err[ng] / np.maximum(np.abs(d64[ng]), 1e-300)
                print("    err/(E+|rho|): " + hist(rel_e, [0, 1e-9, 1e-8, 1e-7, 2e-7, 5e-7, 1e-6, 1e-5, 1]))
                print("    err/|dec|    : " + hist(rel_d, [0, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e9]) + "   max %.2e" % rel_d.max())


if __name__ == "__main__":
    main()
