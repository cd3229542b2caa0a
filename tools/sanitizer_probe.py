#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel of the library runs at least once.
  compute-sanitizer --tool memcheck python tools/sanitizer_probe.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import haf_grasping_b200 as h  # noqa: E402
from haf_grasping_b200 import synth  # noqa: E402

F = os.path.join(ROOT, "tests", "golden", "refdata", "Features.txt")
R = os.path.join(ROOT, "tests", "golden", "refdata", "range21062012_allfeatures")


def staged():
    """round 2, end: host-staged batches -- chunks alternating between three streams / buffer sets, the counter merge, the
    library's own staging of pageable memory (helper + copy threads, pinned ring); CUDA graph capture / replay of single goals"""
    import torch
    tmp = tempfile.mkdtemp()
    model = synth.write_synth_model(os.path.join(tmp, "s.model"), 300, rho=-0.2972253)
    cl = [synth.synth_cloud(500 + i, 9000 + 100 * (i % 5)) for i in range(50)]       # 16 + 24 + 10 clouds: three chunks, three streams
    off = np.concatenate([[0], np.cumsum([len(c) for c in cl])])
    xyz = np.concatenate(cl)                                                          # 5.5 MB pageable: the library stages it itself
    pinned = torch.empty((len(xyz), 3), dtype=torch.float32, pin_memory=True)
    pinned.numpy()[:] = xyz
    g = h.GraspSearch(F, R, model, guard_rel=1e-3, use_graph=True)
    want = [b.astuple() for b in g.search_batch_packed(torch.from_numpy(xyz).cuda(), off)]
    for name, buf in (("pageable", xyz), ("pinned", pinned.numpy())):
        got = [b.astuple() for b in g.search_batch_packed(buf, off)]
        print("staged", name, "chunks", g.timing().n_chunks, "equal", got == want, "guard", g.timing().n_guard)
    for _ in range(4):
        r = g.search(cl[0])
    print("graph replays", g.timing().graph_replays, r["best"].astuple() == want[0])
    g.close()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "staged":
        return staged()
    tmp = tempfile.mkdtemp()
    model = synth.write_synth_model(os.path.join(tmp, "s.model"), 300, rho=-0.2972253)
    clouds = [synth.synth_cloud(77 + i, 6000 + 500 * i) for i in range(3)]
    for mode in (h.HAF_SVM_TENSOR_GUARD, h.HAF_SVM_FP32_GUARD, h.HAF_SVM_FP64_EXACT):
        for tier2 in ((0, 2) if mode != h.HAF_SVM_FP64_EXACT else (0,)):
            g = h.GraspSearch(F, R, model, svm_mode=mode, guard_rel=1e-3 if mode != h.HAF_SVM_FP64_EXACT else 0.0, guard_tier2=tier2)
            res = g.search(clouds[0])
            best = g.search_batch(clouds)
            t = g.timing()
            print("mode", mode, "tier2", tier2, "best", res["best"].astuple(), [b.astuple() for b in best], "guard", t.n_guard, "exact", t.n_exact)
            g.close()
    # round 2: DFMA guard kernel (the DMMA one is the default above), probability mode, device PCD / PointCloud2 ingest
    g = h.GraspSearch(F, R, model, guard_rel=1e-3, guard_kernel=2)
    print("dfma guard", g.search(clouds[0])["best"].astuple(), g.timing().n_guard)
    g.close()
    pm = os.path.join(tmp, "p.model")
    with open(model) as fh:
        head, tail = fh.read().split("nr_sv", 1)
    with open(pm, "w") as fh:
        fh.write(head + "probA -2.5\nprobB 0.1\nnr_sv" + tail)
    g = h.GraspSearch(F, R, pm)
    print("probability", g.search(clouds[1], [h.make_request(svm_with_probability=1)])["best"].astuple())
    xyz = clouds[2]
    hdr = ("# .PCD v0.7\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n")
    ascii_pcd = (hdr % (len(xyz), len(xyz), "ascii")).encode() + "".join("%.9g %.9g %.9g\n" % tuple(p) for p in xyz).encode()
    bin_pcd = (hdr % (len(xyz), len(xyz), "binary")).encode() + xyz.tobytes()
    for name, raw in (("ascii", ascii_pcd), ("binary", bin_pcd)):
        got = g.pcd_decode_to_host(raw)
        print("pcd", name, bool((got == xyz).all()), g.search_pcd(raw)["best"].astuple())
    z = np.load(os.path.join(ROOT, "tests", "golden", "pcd_files.npz"))
    print("pcd lzf", g.pcd_decode(z["table3"].tobytes())[1])
    buf = np.zeros((len(xyz), 16), np.uint8)
    buf[:, :12] = xyz.view(np.uint8).reshape(len(xyz), 12)
    g.pointcloud2_to_xyz(buf, len(xyz), 16)
    print("pc2", bool((g.debug_pcd_xyz(len(xyz)) == xyz).all()))
    g.close()
    p = h.SvmPredictor(pm, min_dims=330)
    rngp = np.random.default_rng(2)
    print("svm prob", p.predict_probability(rngp.uniform(-1, 1, size=(300, 330)))[1][:2].tolist())
    p.close()
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, size=(700, 330)) * (rng.random((700, 330)) < 0.8)
    for mode in (0, 2, 1):
        p = h.SvmPredictor(model, svm_mode=mode, min_dims=330, guard_rel=1e-3 if mode != 1 else 0.0)
        lab, dec = p.predict(x)
        print("svm mode", mode, "labels+", int((lab > 0).sum()), "guard", p.timing().n_guard)
        p.close()
    fmin, fmax = h.scale_minmax(x, 330)
    out = h.scale_apply(x, 330, fmin, fmax)
    print("scale", float(out.min()), float(out.max()))


if __name__ == "__main__":
    main()
