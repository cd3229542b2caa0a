"""Regenerates the committed fixtures under tests/golden/ .  Run in the BUILD container only
(needs /root/reference and oracle/_ref):   python tests/golden/make_golden.py

  clouds.npz                     float32 xyz of every data/*.pcd of the reference, decoded by
                                 haf_grasping_b200.pcd.read_pcd (objects_N are symlinks to tableN: skipped)
  substitute_trained.model.gz    a 2-class RBF C-SVC model trained with the REFERENCE'S OWN svm-train
                                 (oracle/_ref/svm-train) on features the oracle harvests from the bundled
                                 clouds, labelled by a deterministic geometric rule (the real model is
                                 missing from the reference checkout)
  expected.json                  ORACLE-GENERATED regression pins (best grasp tuple, window counts,
                                 per-roll tops, CRC32 of grids/masks/evals) for every cloud x model.
                                 The reference ships no golden outputs; these pin oracle == CUDA path.
"""
import glob
import gzip
import json
import os
import subprocess
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from haf_grasping_b200 import synth  # noqa: E402
from haf_grasping_b200.pcd import read_pcd  # noqa: E402
from oracle import orc  # noqa: E402

REFDATA = "/root/reference/data"
FEATURES = os.path.join(HERE, "refdata", "Features.txt")
RANGE = os.path.join(HERE, "refdata", "range21062012_allfeatures")


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def make_clouds():
    clouds = {}
    for p in sorted(glob.glob(os.path.join(REFDATA, "*.pcd"))):
        if os.path.islink(p):
            continue
        name = os.path.basename(p)[:-4]
        if name.startswith("table"):
            name = name.split("_")[0]
        clouds[name] = read_pcd(p)
    np.savez_compressed(os.path.join(HERE, "clouds.npz"), **clouds)
    return clouds


def geometric_label(patch):
    """+1 when the middle strip of the 14x14 window is >= 1.5 cm higher than both outer strips."""
    def rect(x1, x2, y1, y2):  # inclusive cell ranges on the 15x15 integral patch
        return patch[x2 + 1, y2 + 1] - patch[x1, y2 + 1] - patch[x2 + 1, y1] + patch[x1, y1]
    mid = rect(5, 8, 0, 13) / (4 * 14)
    lo = rect(0, 3, 0, 13) / (4 * 14)
    hi = rect(10, 13, 0, 13) / (4 * 14)
    return 1 if (mid - max(lo, hi)) > 0.015 else -1


def make_trained_model(clouds, o):
    rng = np.random.default_rng(20151015)
    lines = []
    for name, xyz in clouds.items():
        av = o.normalize_approach((0, 0, 1))
        for roll in range(0, 12, 3):
            M = o.build_transform((0, 0, 0), av, 1, roll)
            integral = o.calc_intimage(o.generate_grid(xyz, M))
            mask = o.pnt_in_box(integral, roll)
            feats, rc = o.calc_featurevectors(integral, mask)
            if len(feats) == 0:
                continue
            scaled = o.scale(feats)
            pick = rng.choice(len(feats), size=min(len(feats), 14), replace=False)
            for w in sorted(pick):
                r, c = rc[w]
                lab = geometric_label(integral[r - 7:r + 8, c - 7:c + 8].astype(np.float64))
                lines.append("%+d " % lab + " ".join("%d:%.6g" % (k + 1, v) for k, v in enumerate(scaled[w]) if v != 0))
    train = "/tmp/haf_substitute_train.txt"
    with open(train, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    model = "/tmp/haf_substitute_trained.model"
    subprocess.run([os.path.join(orc.REF_DIR, "svm-train"), "-q", "-g", "0.02", "-c", "8", train, model], check=True)
    with open(model, "rb") as fh, gzip.GzipFile(os.path.join(HERE, "substitute_trained.model.gz"), "wb", mtime=0) as gz:
        gz.write(fh.read())
    npos = sum(1 for ln in lines if ln.startswith("+1"))
    print("trained on", len(lines), "samples (", npos, "positive ); model", os.path.getsize(model), "bytes")
    return model


def main():
    orc.build(ref=True)
    clouds = make_clouds()
    o = orc.Oracle(FEATURES, RANGE)
    trained = make_trained_model(clouds, o)
    synth_path = synth.write_synth_model("/tmp/haf_synth256.model", n_sv=256, seed=7)
    expected = {}
    for mname, mpath in (("trained", trained), ("synth256", synth_path)):
        o.load_model(mpath)
        for name, xyz in clouds.items():
            res = o.search(xyz, orc.make_request())
            b = res["best"]
            expected["%s/%s" % (mname, name)] = {
                "best": list(b.astuple()), "eval": b.eval, "n_windows": int(b.n_windows),
                "per_roll_top": res["per_roll_top"].tolist(),
                "mask_popcount": res["mask"].reshape(len(res["mask"]), -1).sum(1).tolist(),
                "crc_heights": crc(res["heights"]), "crc_integral": crc(res["integral"]),
                "crc_mask": crc(res["mask"]), "crc_graspseval": crc(res["graspseval"]),
                "n_positive": int((res["dec"] > 0).sum()),
            }
            print(mname, name, expected["%s/%s" % (mname, name)]["best"], b.n_windows)
    with open(os.path.join(HERE, "expected.json"), "w") as fh:
        json.dump(expected, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
