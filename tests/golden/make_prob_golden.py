"""Probability-model fixture.  Run in the BUILD container only (needs /root/reference and oracle/_ref):
    python tests/golden/make_prob_golden.py

The trained substitute model (substitute_trained.model.gz, see make_golden.py) is re-trained with the REFERENCE'S OWN
svm-train and `-b 1` on the same training file: libsvm then appends the sigmoid of svm_binary_svc_probability
(svm.cpp:1893-1978: 5-fold cross-validated decision values -> sigmoid_train) as the lines `probA` / `probB`; the support
vectors are those of the -b 0 model (same final training run).  Only those two lines are committed
(substitute_trained.prob.json); tests splice them into the committed model.  Also committed: the reference's own
`svm-predict -b 1` output for the 20-row file of tests/golden/svm_cli (prob_out_trained_ref.txt / .stdout).
"""
import gzip
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

import make_golden as mg  # noqa: E402
from oracle import orc  # noqa: E402


def splice(model_text: str, probA: str, probB: str) -> str:
    """libsvm's writer puts probA / probB between `label` and `nr_sv` (svm.cpp:2640-2655)"""
    head, tail = model_text.split("nr_sv", 1)
    return head + "probA %s\nprobB %s\nnr_sv" % (probA, probB) + tail


def main():
    orc.build(ref=True)
    clouds = {k: v for k, v in np.load(os.path.join(HERE, "clouds.npz")).items()}
    o = orc.Oracle(mg.FEATURES, mg.RANGE)
    # same training file as make_golden.make_trained_model (deterministic); it leaves /tmp/haf_substitute_train.txt behind
    real_gz = os.path.join(HERE, "substitute_trained.model.gz")
    keep = open(real_gz, "rb").read()
    try:
        mg.make_trained_model(clouds, o)
    finally:
        open(real_gz, "wb").write(keep)   # make_trained_model rewrites the .gz: keep the committed bytes
    train = "/tmp/haf_substitute_train.txt"
    pm = "/tmp/haf_substitute_trained.prob.model"
    subprocess.run([os.path.join(orc.REF_DIR, "svm-train"), "-q", "-b", "1", "-g", "0.02", "-c", "8", train, pm], check=True)
    text = open(pm).read()
    pa = [ln for ln in text.split("\n") if ln.startswith("probA ")][0].split(" ", 1)[1]
    pb = [ln for ln in text.split("\n") if ln.startswith("probB ")][0].split(" ", 1)[1]
    base = gzip.open(real_gz, "rb").read().decode()
    assert splice(base, pa, pb) == text, "the -b 1 model differs from the committed model in more than probA / probB"
    with open(os.path.join(HERE, "substitute_trained.prob.json"), "w") as fh:
        json.dump({"probA": pa, "probB": pb, "note": "svm-train -q -b 1 -g 0.02 -c 8 on make_golden's training file"}, fh, indent=1)
    # the reference program's own -b 1 output on the committed 20-row scaled file
    scaled = os.path.join(HERE, "svm_cli", "scaled_ref.txt")
    out = os.path.join(HERE, "svm_cli", "prob_out_trained_ref.txt")
    res = subprocess.run([os.path.join(orc.REF_DIR, "svm-predict"), "-b", "1", scaled, pm, out], capture_output=True, text=True, check=True)
    open(os.path.join(HERE, "svm_cli", "prob_out_trained_ref.stdout"), "w").write(res.stdout)
    print("probA", pa, "probB", pb, "|", res.stdout.strip())


if __name__ == "__main__":
    main()
