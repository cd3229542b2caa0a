"""Deterministic synthetic workloads for tests and bench.py.

* ``synth_cloud``  -- counter-based (splitmix64) point clouds: a floor at z = 0.02 m plus 8 Gaussian
  "objects", uniform in x,y over +-r.  Same numbers on every box for a given seed.
* ``write_synth_model`` -- a substitute SVM model in libsvm 3.12's text format (the trained model
  ``data/all_features.txt.scale.model`` is missing from the reference checkout, see
  /root/reference/.MISSING_LARGE_BLOBS).  Format follows the writer at libsvm-3.12/svm.cpp:2599-2691.

Host-side numpy only; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import os

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed, counter):
    """splitmix64 finaliser of (seed * golden + counter); vectorised over ``counter``."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.asarray(counter, dtype=np.uint64)
             + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _u01(seed, counter):
    """uniform [0,1) float32 from the top 24 bits."""
    return (splitmix64(seed, counter) >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))


def synth_cloud(seed: int, n_points: int, r: float = 0.28) -> np.ndarray:
    """float32 [n_points, 3]; x,y uniform in (-r, r), z = floor + 8 Gaussian bumps + 2 mm noise."""
    k = np.arange(n_points, dtype=np.uint64)
    u1 = _u01(seed, 3 * k)
    u2 = _u01(seed, 3 * k + np.uint64(1))
    u3 = _u01(seed, 3 * k + np.uint64(2))
    r32 = np.float32(r)
    x = (np.float32(2) * u1 - np.float32(1)) * r32
    y = (np.float32(2) * u2 - np.float32(1)) * r32
    z = np.full(n_points, 0.02, np.float32)
    stream = np.uint64(1) << np.uint64(32)
    for o in range(8):
        p = _u01(seed, stream + np.arange(4 * o, 4 * o + 4, dtype=np.uint64))
        a = (np.float32(2) * p[0] - np.float32(1)) * np.float32(0.6) * r32
        b = (np.float32(2) * p[1] - np.float32(1)) * np.float32(0.6) * r32
        h = np.float32(0.03) + np.float32(0.17) * p[2]
        s = np.float32(0.02) + np.float32(0.06) * p[3]
        d2 = (x - a) ** 2 + (y - b) ** 2
        z = z + h * np.exp(-d2 / (np.float32(2) * s * s)).astype(np.float32)
    z = z + np.float32(0.002) * (u3 - np.float32(0.5))
    return np.stack([x, y, z.astype(np.float32)], axis=1).astype(np.float32)


def synth_model_arrays(n_sv: int, dim: int = 323, seed: int = 7):
    """coef [n_sv] (first half > 0, second half < 0, zero-sum), SV [n_sv, dim] with 6 significant digits."""
    rng = np.random.default_rng(seed)
    sv = rng.uniform(-1.0, 1.0, size=(n_sv, dim))
    sv = np.array([[float("%.6g" % v) for v in row] for row in sv])
    n0 = n_sv // 2
    coef = rng.uniform(0.05, 1.0, size=n_sv)
    coef[n0:] *= -1.0
    coef[:n0] *= (-coef[n0:].sum()) / coef[:n0].sum()  # zero-sum -> decision values straddle 0
    return coef, sv, n0


def write_synth_model(path: str, n_sv: int = 2048, gamma: float = 1.0 / 323.0, seed: int = 7, dim: int = 323,
                      rho: float = 0.0, labels=(1, -1)) -> str:
    """Write a dense 2-class RBF C-SVC model in libsvm text format; returns ``path``."""
    coef, sv, n0 = synth_model_arrays(n_sv, dim, seed)
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "w") as fh:
        fh.write("svm_type c_svc\nkernel_type rbf\n")
        fh.write("gamma %g\n" % gamma)
        fh.write("nr_class 2\ntotal_sv %d\n" % n_sv)
        fh.write("rho %g\n" % rho)
        fh.write("label %d %d\n" % (labels[0], labels[1]))
        fh.write("nr_sv %d %d\nSV\n" % (n0, n_sv - n0))
        for i in range(n_sv):
            fh.write("%.16g " % coef[i])
            fh.write(" ".join("%d:%.8g" % (d + 1, sv[i, d]) for d in range(dim) if sv[i, d] != 0))
            fh.write(" \n")
    os.replace(tmp, path)
    return path
