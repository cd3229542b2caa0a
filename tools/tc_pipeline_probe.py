#!/usr/bin/env python
"""Timing experiments on the X-resident tensor kernel's pipeline (HAF_TC_DEBUG knobs of svm_rbf_tc3_kernel): which of
MMA issue, TMEM loads and epilogue arithmetic bounds a tile, and whether they overlap.  Results of a debug run are garbage
by construction; only the SVM stage time is read.  Usage: python tools/tc_pipeline_probe.py  (on the GPU box)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for flags, what in ((0, "full kernel"), (16, "full kernel, clusters walk the SV tiles in step (no skew)"), (19, "skeleton, no skew"), (2, "no MMAs (epilogue alone, TMA still streaming)"),
                    (3, "neither (barrier / TMA skeleton)"), (4, "MMAs + TMEM loads, no arithmetic"), (8, "MMAs + arithmetic, no TMEM loads")):
    env = dict(os.environ, HAF_TC_DEBUG=str(flags))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu-baseline"], env=env,
                         capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        rows.append((flags, what, d["stage_ms_per_step"]["svm"], d["roofline"]["kernel"]))
    except Exception as exc:
        rows.append((flags, what, None, "failed: %s %s" % (exc, out.stderr[-300:])))
    print(rows[-1], flush=True)
print(json.dumps(rows))
