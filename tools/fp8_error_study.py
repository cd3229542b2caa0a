#!/usr/bin/env python
"""CPU study (no GPU): how large would the decision-value error of the tensor path be if the two cross terms of the split
contraction (x_hi.sv_lo + x_lo.sv_hi) were computed with FP8 (e4m3) operands instead of fp16 -- 2 bf16-equivalent MMA
passes instead of 3.  Operand rounding is emulated with torch dtypes, accumulation in float64, so
the numbers isolate the OPERAND error of each scheme; they are reported as a fraction of the guard scale
E = sum_i |coef_i| K_i (1 + gamma log2(e) (|x|^2 + |sv_i|^2)) + |rho|, like tools/dec_error_probe.py.

  python tools/fp8_error_study.py          (needs oracle/libhaf_oracle.so; ~1 minute)
"""
import gzip
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from haf_grasping_b200 import synth  # noqa: E402
from oracle import orc  # noqa: E402
from tools.dec_error_probe import load_model  # noqa: E402

F = os.path.join(ROOT, "tests", "golden", "refdata", "Features.txt")
R = os.path.join(ROOT, "tests", "golden", "refdata", "range21062012_allfeatures")


def q(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dtype).to(torch.float64).numpy()


def study(name, model_path, scaled):
    gamma, rho, coef, sv = load_model(model_path)
    D = max(sv.shape[1], scaled.shape[1])
    x = np.zeros((len(scaled), D)); x[:, :scaled.shape[1]] = scaled
    s = np.zeros((len(sv), D)); s[:, :sv.shape[1]] = sv
    xn, sn = (x * x).sum(1), (s * s).sum(1)
    base = xn[:, None] + sn[None, :]

    def dec_from_dot(dot):
        d2 = np.maximum(base - 2.0 * dot, 0.0)
        return np.exp(-gamma * d2) @ coef - rho

    K = np.exp(-gamma * np.maximum(base - 2.0 * x @ s.T, 0.0))
    E = (K * (1.0 + gamma * 1.4426950408889634 * base)) @ np.abs(coef) + abs(rho)
    ref = dec_from_dot(x @ s.T)
    xh, sh = q(x, torch.float16), q(s, torch.float16)
    xl, sl = q(x - xh, torch.float16), q(s - sh, torch.float16)
    schemes = {
        "fp16 x3 (current)": xh @ sh.T + xh @ sl.T + xl @ sh.T,
        "bf16 x3 (first version)": None,
        "fp16 main + e4m3 cross (2 passes)": xh @ sh.T + (q(x, torch.float8_e4m3fn) @ q((s - sh) * 4096.0, torch.float8_e4m3fn).T
                                                         + q((x - xh) * 4096.0, torch.float8_e4m3fn) @ q(s, torch.float8_e4m3fn).T) / 4096.0,
        "fp16 x1 (1 pass)": xh @ sh.T,
    }
    bh, bsh = q(x, torch.bfloat16), q(s, torch.bfloat16)
    bl, bsl = q(x - bh, torch.bfloat16), q(s - bsh, torch.bfloat16)
    schemes["bf16 x3 (first version)"] = bh @ bsh.T + bh @ bsl.T + bl @ bsh.T
    for sname, dot in schemes.items():
        err = np.abs(dec_from_dot(dot) - ref) / E
        print("%-10s %-36s max err/E = %.2e   rms = %.2e" % (name, sname, err.max(), np.sqrt((err ** 2).mean())))


def main():
    orc.build(ref=False)
    tmp = tempfile.mkdtemp()
    trained = os.path.join(tmp, "trained.model")
    with gzip.open(os.path.join(ROOT, "tests", "golden", "substitute_trained.model.gz"), "rb") as src, open(trained, "wb") as dst:
        dst.write(src.read())
    models = {"trained": trained, "synth2048": synth.write_synth_model(os.path.join(tmp, "s2048.model"), 2048)}
    xyz = np.load(os.path.join(ROOT, "tests", "golden", "clouds.npz"))["table1"]
    o = orc.Oracle(F, R, trained)
    ores = o.search(xyz, orc.make_request())
    rows = []
    for roll in range(0, 12, 3):
        feats, _ = o.calc_featurevectors(ores["integral"][roll], ores["mask"][roll])
        rows.append(o.scale(feats))
    scaled = np.concatenate(rows)
    print("windows", len(scaled))
    for name, mp in models.items():
        study(name, mp, scaled)


if __name__ == "__main__":
    main()
