"""Regenerates tests/golden/svm_cli/: inputs and outputs of the REFERENCE'S OWN svm-scale / svm-predict programs
(oracle/_ref = libsvm-3.12 compiled unmodified from /root/reference), so that the GPU tests of the libsvm front ends
(tests/test_svm_frontends.py) have reference outputs even where oracle/_ref is not present.
Run in the BUILD container only:   python tests/golden/make_svm_cli_golden.py

  features.txt          20 rows of /tmp/features.txt of roll 3 of data/pcd2.pcd, written by the reference's own
                        CIntImage_to_Featurevec::write_featurevector (324 "%.4g" values per row)
  scaled_ref.txt        svm-scale -r data/range21062012_allfeatures features.txt
  sparse.txt            features.txt with |v| < 0.02 dropped and labels -1 / 0 / 1 (absent entries matter)
  fit_<k>_ref.txt, fit_<k>_ref.range, fit_<k>_ref.stderr      svm-scale <args k> -s range sparse.txt  (no -r: fitted)
  out_trained_ref.txt, out_synth_ref.txt, *_ref.stdout        svm-predict scaled_ref.txt <model> out
                        (trained = tests/golden/substitute_trained.model.gz, synth = synth.write_synth_model(256))
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from haf_grasping_b200 import synth  # noqa: E402
from oracle import orc  # noqa: E402

FEATURES = os.path.join(HERE, "refdata", "Features.txt")
RANGE = os.path.join(HERE, "refdata", "range21062012_allfeatures")
OUT = os.path.join(HERE, "svm_cli")
FIT_ARGS = [[], ["-l", "0", "-u", "1"], ["-y", "-3", "5"], ["-l", "-2.5", "-u", "0.75", "-y", "0", "1"]]


def main():
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp()
    trained = os.path.join(tmp, "trained.model")
    with gzip.open(os.path.join(HERE, "substitute_trained.model.gz"), "rb") as src, open(trained, "wb") as dst:
        dst.write(src.read())
    synth_model = synth.write_synth_model(os.path.join(tmp, "synth256.model"), 256)
    clouds = np.load(os.path.join(HERE, "clouds.npz"))
    o = orc.Oracle(FEATURES, RANGE, trained)
    ores = o.search(clouds["pcd2"], orc.make_request())
    ref = orc.Ref(FEATURES, RANGE, trained)
    ref.roll_file_exact(ores["integral"][3], ores["mask"][3], workdir=tmp)
    rows = open(os.path.join(tmp, "features.txt")).read().split("\n")[:20]
    f_feat = os.path.join(OUT, "features.txt")
    open(f_feat, "w").write("\n".join(rows) + "\n")
    scale, predict = os.path.join(orc.REF_DIR, "svm-scale"), os.path.join(orc.REF_DIR, "svm-predict")
    with open(os.path.join(OUT, "scaled_ref.txt"), "w") as so:
        subprocess.run([scale, "-r", RANGE, f_feat], stdout=so, stderr=subprocess.DEVNULL, check=True)
    sparse = []
    for k, ln in enumerate(rows):
        t = ln.split()
        sparse.append(" ".join([str(k % 3 - 1)] + [a for a in t[1:] if abs(float(a.split(":")[1])) >= 0.02]))
    f_sparse = os.path.join(OUT, "sparse.txt")
    open(f_sparse, "w").write("\n".join(sparse) + "\n")
    for k, args in enumerate(FIT_ARGS):
        r = subprocess.run([scale] + args + ["-s", os.path.join(OUT, "fit_%d_ref.range" % k), f_sparse], capture_output=True, check=True)
        open(os.path.join(OUT, "fit_%d_ref.txt" % k), "wb").write(r.stdout)
        open(os.path.join(OUT, "fit_%d_ref.stderr" % k), "wb").write(r.stderr)
    for name, model in (("trained", trained), ("synth", synth_model)):
        r = subprocess.run([predict, os.path.join(OUT, "scaled_ref.txt"), model, os.path.join(OUT, "out_%s_ref.txt" % name)],
                           capture_output=True, check=True)
        open(os.path.join(OUT, "out_%s_ref.stdout" % name), "wb").write(r.stdout)
    json.dump({"fit_args": FIT_ARGS}, open(os.path.join(OUT, "manifest.json"), "w"))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
