"""The reference's bundled point clouds as FILE BYTES (input fixtures for the device PCD ingest).
Run in the BUILD container only:  python tests/golden/make_pcd_fixture.py  ->  tests/golden/pcd_files.npz
Keys as in clouds.npz (objects_N are symlinks to tableN: skipped); each entry is the file's bytes as uint8."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDATA = "/root/reference/data"

files = {}
for p in sorted(glob.glob(os.path.join(REFDATA, "*.pcd"))):
    if os.path.islink(p):
        continue
    name = os.path.basename(p)[:-4]
    if name.startswith("table"):
        name = name.split("_")[0]
    files[name] = np.frombuffer(open(p, "rb").read(), np.uint8)
np.savez_compressed(os.path.join(HERE, "pcd_files.npz"), **files)
print({k: len(v) for k, v in files.items()})
