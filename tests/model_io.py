"""libsvm text model -> numpy arrays (test helper; the product's own loader is csrc/haf_host.hpp)."""
import numpy as np


def load_model_arrays(path, dim=None):
    """returns dict(gamma, rho, coef [S], sv [S, dim]) of a 2-class RBF C-SVC model in libsvm 3.12 text format"""
    gamma, rho, rows, coefs, in_sv = None, 0.0, [], [], False
    with open(path) as fh:
        for ln in fh:
            if in_sv:
                t = ln.split()
                if not t:
                    continue
                coefs.append(float(t[0]))
                rows.append([(int(a.split(":")[0]), float(a.split(":")[1])) for a in t[1:]])
            elif ln.startswith("gamma"):
                gamma = float(ln.split()[1])
            elif ln.startswith("rho"):
                rho = float(ln.split()[1])
            elif ln.startswith("SV"):
                in_sv = True
    width = max([i for r in rows for i, _ in r] + [dim or 1])
    sv = np.zeros((len(rows), width))
    for k, r in enumerate(rows):
        for i, v in r:
            sv[k, i - 1] = v
    return dict(gamma=gamma, rho=rho, coef=np.array(coefs), sv=sv)


def coefK_sum(model, scaled, block=2048):
    """sum_i |coef_i| K_i per row of `scaled` (float64): the natural scale of a decision value's rounding error"""
    sv, coef, g = model["sv"], np.abs(model["coef"]), model["gamma"]
    D = max(sv.shape[1], scaled.shape[1])
    svp = np.zeros((len(sv), D)); svp[:, :sv.shape[1]] = sv
    sn = (svp ** 2).sum(1)
    out = np.zeros(len(scaled))
    for b in range(0, len(scaled), block):
        x = np.zeros((min(block, len(scaled) - b), D)); x[:, :scaled.shape[1]] = scaled[b:b + block]
        d2 = np.maximum((x ** 2).sum(1)[:, None] + sn[None, :] - 2.0 * (x @ svp.T), 0.0)
        out[b:b + block] = np.exp(-g * d2) @ coef
    return out
