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


def coefK_sum(model, scaled, block=2048, with_E=False):
    """sum_i |coef_i| K_i per row of `scaled` (float64): the natural scale of a decision value's rounding error.
    with_E: also E = sum_i |coef_i| K_i (1 + gamma log2(e) (|x|^2 + |sv_i|^2)), the guard band's scale (svm_tc.cuh): an
    exponential evaluated in FP32 is off by ~2^-24 times the MAGNITUDE of its argument, relative -- whatever the
    implementation -- so each term's error budget grows with the numbers its argument is assembled from."""
    sv, coef, g = model["sv"], np.abs(model["coef"]), model["gamma"]
    g2 = g * np.log2(np.e)
    D = max(sv.shape[1], scaled.shape[1])
    svp = np.zeros((len(sv), D)); svp[:, :sv.shape[1]] = sv
    sn = (svp ** 2).sum(1)
    out = np.zeros(len(scaled))
    outE = np.zeros(len(scaled))
    for b in range(0, len(scaled), block):
        x = np.zeros((min(block, len(scaled) - b), D)); x[:, :scaled.shape[1]] = scaled[b:b + block]
        xn = (x ** 2).sum(1)
        base = xn[:, None] + sn[None, :]
        K = np.exp(-g * np.maximum(base - 2.0 * (x @ svp.T), 0.0))
        out[b:b + block] = K @ coef
        if with_E:
            outE[b:b + block] = (K * (1.0 + g2 * base)) @ coef
    return (out, outE) if with_E else out


def check_decision_tolerance(model, scaled, dec_gpu, dec_ref, rtol_E=2e-6, rtol_plain=1e-5):
    """The stated tolerance of the FP32 / tensor decision values (north_star: <= 1e-5 relative in FP32), per window:
         |dec_gpu - dec_ref| <= rtol_E * E(window)            E as above; measured <= 5e-7 (profiles/r2_final_dec_error_probe.txt);
       and wherever the exponent arguments are moderate (gamma log2e (|x|^2 + max|sv|^2) <= 4) the plain statement
         |dec_gpu - dec_ref| <= 1e-5 * sum_i |coef_i| K_i(window).
    Returns (max err / E, max err / sum|coef|K over the moderate windows)."""
    plain, E = coefK_sum(model, scaled, with_E=True)
    err = np.abs(dec_gpu - dec_ref)
    tiny = 1e-300
    rE = err / np.maximum(E, tiny)
    assert (err <= rtol_E * E + tiny).all(), ("decision values off by more than %g E" % rtol_E, float(rE.max()), int((err > rtol_E * E + tiny).sum()))
    g2 = model["gamma"] * np.log2(np.e)
    D = max(model["sv"].shape[1], scaled.shape[1])
    snmax = (model["sv"] ** 2).sum(1).max()
    moderate = g2 * ((scaled ** 2).sum(1) + snmax) <= 4.0
    rP = err[moderate] / np.maximum(plain[moderate], tiny)
    assert (rP <= rtol_plain).all(), ("decision values off by more than %g sum|coef|K" % rtol_plain, float(rP.max()))
    return float(rE.max(initial=0.0)), float(rP.max(initial=0.0))
