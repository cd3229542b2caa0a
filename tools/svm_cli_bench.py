#!/usr/bin/env python
"""Times the libsvm front ends next to the reference's own programs on the same files.

  features text (unscaled, 324 values per row, what CIntImage_to_Featurevec writes)  --svm-scale-->  scaled text
  scaled text  --svm-predict-->  labels

Rows are real SVM inputs: windows of synthetic clouds, pushed through the GPU path's bit-exact feature / scaling stages
and printed the way the reference prints them ("%.4g" / "%g").  The reference programs (oracle/_ref, libsvm-3.12 compiled
unmodified) run on a bounded sample of the same rows; output files are compared byte for byte on that sample.
Prints one JSON line; run on the GPU box:  python tools/svm_cli_bench.py [--rows 60000] [--ref-rows 3000]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import haf_grasping_b200 as h  # noqa: E402
from haf_grasping_b200 import build, synth  # noqa: E402
from oracle import orc  # noqa: E402

F = os.path.join(ROOT, "tests", "golden", "refdata", "Features.txt")
R = os.path.join(ROOT, "tests", "golden", "refdata", "range21062012_allfeatures")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=60000)
    ap.add_argument("--ref-rows", type=int, default=3000)
    ap.add_argument("--n-sv", type=int, default=2048)
    a = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="svmcli_")
    model = synth.write_synth_model(os.path.join(tmp, "synth.model"), a.n_sv)
    tools = build.build_svm_tools()
    g = h.GraspSearch(F, R, model)
    raw_rows = []
    seed = 1234
    while sum(len(r) for r in raw_rows) < a.rows:
        g.search(synth.synth_cloud(seed, 100000))
        raw, _ = g.debug_features(scaled=False)
        raw_rows.append(raw)
        seed += 1
    g.close()
    raw = np.concatenate(raw_rows)[:a.rows]
    f_feat = os.path.join(tmp, "features.txt")
    with open(f_feat, "w") as fh:       # write_featurevector, II2FV.cpp:122-137: "+1 k:%.4g ..."
        for row in raw:
            fh.write("+1" + "".join(" %d:%.4g" % (k + 1, v) for k, v in enumerate(row)) + "\n")
    f_feat_s = os.path.join(tmp, "features_sample.txt")
    with open(f_feat) as src, open(f_feat_s, "w") as dst:
        for k, ln in enumerate(src):
            if k >= a.ref_rows:
                break
            dst.write(ln)

    def timed(cmd, stdout=None):
        t0 = time.perf_counter()
        subprocess.run(cmd, stdout=stdout, stderr=subprocess.DEVNULL, check=True)
        return time.perf_counter() - t0

    out = {"rows": int(len(raw)), "ref_rows": a.ref_rows, "n_sv": a.n_sv, "features_text_MB": os.path.getsize(f_feat) / 1e6}
    # svm-scale
    f_sc, f_sc_s, f_sc_ref = (os.path.join(tmp, n) for n in ("scaled.txt", "scaled_sample.txt", "scaled_sample_ref.txt"))
    with open(f_sc, "w") as so:
        out["scale_b200_s"] = timed([tools[1], "-r", R, f_feat], so)
    with open(f_sc_s, "w") as so:
        out["scale_b200_sample_s"] = timed([tools[1], "-r", R, f_feat_s], so)
    with open(f_sc_ref, "w") as so:
        out["scale_ref_sample_s"] = timed([os.path.join(orc.REF_DIR, "svm-scale"), "-r", R, f_feat_s], so)
    out["scale_sample_identical"] = open(f_sc_s, "rb").read() == open(f_sc_ref, "rb").read()
    # svm-predict
    f_o, f_o_s, f_o_ref = (os.path.join(tmp, n) for n in ("out.txt", "out_sample.txt", "out_sample_ref.txt"))
    out["predict_b200_s"] = timed([tools[0], f_sc, model, f_o], subprocess.DEVNULL)
    out["predict_b200_sample_s"] = timed([tools[0], f_sc_s, model, f_o_s], subprocess.DEVNULL)
    out["predict_ref_sample_s"] = timed([os.path.join(orc.REF_DIR, "svm-predict"), f_sc_ref, model, f_o_ref], subprocess.DEVNULL)
    out["predict_sample_identical"] = open(f_o_s, "rb").read() == open(f_o_ref, "rb").read()
    out["predict_rows_per_s_b200_cli"] = len(raw) / out["predict_b200_s"]
    out["predict_rows_per_s_ref_cli"] = a.ref_rows / out["predict_ref_sample_s"]
    # C ABI alone (rows already parsed): device time and wall time of haf_svm_predict
    rows = []
    with open(f_sc) as fh:
        for ln in fh:
            x = np.zeros(324)
            for t in ln.split()[1:]:
                i, v = t.split(":")
                x[int(i) - 1] = float(v)
            rows.append(x)
    x = np.array(rows)
    for mode, name in ((0, "tensor"), (2, "fp32"), (1, "fp64_exact")):
        if mode == 1:
            xs = x[:a.ref_rows * 2]
        else:
            xs = x
        p = h.SvmPredictor(model, svm_mode=mode, min_dims=324)
        p.predict(xs)
        t0 = time.perf_counter()
        p.predict(xs)
        wall = time.perf_counter() - t0
        t = p.timing()
        out["abi_%s" % name] = {"rows": len(xs), "device_ms": t.ms_total, "wall_ms": wall * 1e3, "rows_per_s_device": len(xs) / (t.ms_total * 1e-3),
                                "guard_rows": int(t.n_guard), "exact_rows": int(t.n_exact)}
        p.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
