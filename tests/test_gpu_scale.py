"""GPU parity at BASELINE.json's own sizes and at the bench's operating point ("Next round" 1 and 3):

* configs[3]: synthetic 1 M-point cloud, G = 512, area 362 x 362, all 12 rolls -- every stage against the oracle;
* the bench batch (512 clouds x 100 k points, 2048 support vectors, one tensor-core product, gamma = 1/323): every one of
  its ~3.4 M labels against the FP64 exact-order mode (itself oracle-checked), and 8 clouds label by label against the oracle;
* the audit that backs the calibrated number of tensor-core products, and models far from the calibration's assumptions.

The oracle legs run in a process pool (the oracle is a single-threaded C library; the text round trips dominate).
"""
import hashlib
import multiprocessing as mp
import os

import numpy as np
import pytest

from conftest import FEATURES, RANGE
from model_io import check_decision_tolerance, load_model_arrays

pytestmark = pytest.mark.gpu


def _pool():
    return mp.get_context("fork").Pool(max(1, min(os.cpu_count() or 1, 16)))


# ---- oracle workers (top level: picklable) ---------------------------------------------------------------------------
_XYZ = {}    # inherited by the forked workers


def _oracle_roll_slice(job):
    """one slice of one roll through the oracle's stage functions: (roll, k, n) -> decision values / labels of the k-th
    of n slices of the roll's windows; slice 0 also returns the roll's grids"""
    key, model, G, area, step, roll, k, n, want_hash = job
    from oracle import orc
    o = orc.Oracle(FEATURES, RANGE, model)
    xyz = _XYZ[key]
    M = o.build_transform((0, 0, 0), o.normalize_approach((0, 0, 1)), 1, roll, step)
    heights = o.generate_grid(xyz, M, G)
    integral = o.calc_intimage(heights)
    mask = o.pnt_in_box(integral, roll, (int(area[0]), int(area[1])), step)
    feats, rc = o.calc_featurevectors(integral, mask)
    W = len(feats)
    lo, hi = W * k // n, W * (k + 1) // n
    scaled = o.scale(feats[lo:hi])
    dec, lab = o.svm_decision(scaled)
    out = dict(roll=roll, k=k, dec=dec, lab=lab, W=W)
    if want_hash:
        out["raw_sha"] = hashlib.sha1(feats[lo:hi].tobytes()).hexdigest()
        out["scaled_sha"] = hashlib.sha1(scaled[:, :].tobytes()).hexdigest()
        out["scaled"] = scaled
    if k == 0:
        out.update(heights=heights, integral=integral, mask=mask, cells=rc[:, 0] * G + rc[:, 1])
    return out


def _oracle_rolls(key, model, G, area, rolls, slices, want_hash=False, step=15):
    jobs = [(key, model, G, area, step, roll, k, slices, want_hash) for roll in rolls for k in range(slices)]
    with _pool() as pool:
        parts = pool.map(_oracle_roll_slice, jobs, chunksize=1)
    res = {}
    for roll in rolls:
        ps = sorted([p for p in parts if p["roll"] == roll], key=lambda p: p["k"])
        r = dict(ps[0])
        r["dec"] = np.concatenate([p["dec"] for p in ps])
        r["lab"] = np.concatenate([p["lab"] for p in ps])
        if want_hash:
            r["scaled"] = np.concatenate([p["scaled"] for p in ps])
            r["parts"] = ps
        res[roll] = r
    return res


def _oracle_cloud(job):
    key, model = job
    from oracle import orc
    o = orc.Oracle(FEATURES, RANGE, model)
    res = o.search(_XYZ[key], orc.make_request(), full=True)
    return key, res["dec"], res["mask"], res["graspseval"], res["per_roll_top"], res["best"].astuple(), int(res["best"].n_windows)


# ---- configs[3] ---------------------------------------------------------------------------------------------------------
def test_config4_grid512_all_stages(hg, oracle_lib, tmp_models):
    """BASELINE configs[3] at its own shape: 1 M points, G = 512, area 362 x 362 (121 801 geometric windows per roll), 12 rolls.
    64-SV model so that the oracle finishes; heights / integral images / masks / window lists / labels / graspseval /
    per-roll tops / best grasp must equal the oracle's (the reference's loop rules are replayed on the oracle's per-roll tops)."""
    from haf_grasping_b200 import synth
    model = tmp_models(64, seed=11)
    xyz = synth.synth_cloud(1234, 1000000, r=2.56)
    _XYZ["g512"] = xyz
    G, area = 512, (362.0, 362.0)
    gpu = hg.GraspSearch(FEATURES, RANGE, model, grid=G)
    try:
        R = gpu.R
        assert R == 12
        ores = _oracle_rolls("g512", model, G, area, list(range(R)), slices=4)
        gres = gpu.search(xyz, [hg.make_request(area=area)])
        integral = gpu.debug_integral(R)
        win = gpu.debug_windows()
        dec, lab, guard = gpu.debug_decisions()
        order = np.lexsort((win[:, 1], win[:, 0]))
        win, dec, lab = win[order], dec[order], lab[order]
        mdl = load_model_arrays(model)
        best, pos = (-1000, -1, -1, -1), 0
        for roll in range(R):
            o = ores[roll]
            assert np.array_equal(gres["heights"][0][roll], o["heights"]), roll
            assert integral[roll].tobytes() == o["integral"].tobytes(), roll
            assert np.array_equal(gres["mask"][0][roll], o["mask"]), roll
            W = o["W"]
            sel = slice(pos, pos + W)
            assert (win[sel, 0] == roll).all() and np.array_equal(win[sel, 1], o["cells"]), roll
            o_lab = np.where(o["dec"] > 0, gpu.info.label0, gpu.info.label1)
            assert np.array_equal(lab[sel], o_lab), (roll, int((lab[sel] != o_lab).sum()))
            ev, top, _ = oracle_lib.Oracle(FEATURES, RANGE, model).show_predicted_gps(o["lab"], o["mask"])
            assert np.array_equal(gres["graspseval"][0][roll], ev), roll
            assert tuple(gres["per_roll_top"][0][roll]) == top, (roll, tuple(gres["per_roll_top"][0][roll]), top)
            if top[2] > best[0]:   # server.cpp:953 strict >
                best = (top[2], top[0], top[1], roll)
            pos += W
        assert pos == len(win) == gres["best"].n_windows_scored
        b = gres["best"]
        assert (b.topval, b.row, b.col, b.roll) == best
        assert pos > 1000000          # the config's size: ~1.4 M windows
    finally:
        gpu.close()
        _XYZ.pop("g512", None)


def test_config4_grid512_one_roll_features_and_2048_sv(hg, oracle_lib, tmp_models):
    """Roll 0 of configs[3] in depth: raw features and scaled SVM inputs bit-exact for all 121 801 windows (compared by
    SHA-1 over the oracle's slices), and the 2048-SV bench model's decision values / labels at this grid size."""
    from haf_grasping_b200 import synth
    model = tmp_models(2048)
    xyz = synth.synth_cloud(1234, 1000000, r=2.56)
    _XYZ["g512"] = xyz
    G, area = 512, (362.0, 362.0)
    gpu = hg.GraspSearch(FEATURES, RANGE, model, grid=G)
    try:
        o = _oracle_rolls("g512", model, G, area, [0], slices=16, want_hash=True)[0]
        gpu.search(xyz, [hg.make_request(area=area, roll_limit=1)], outputs=False)
        win = gpu.debug_windows()
        raw, scaled = gpu.debug_features()
        dec, lab, guard = gpu.debug_decisions()
        order = np.argsort(win[:, 1], kind="stable")
        assert np.array_equal(win[order, 1], o["cells"])
        raw, scaled, dec, lab = raw[order], scaled[order], dec[order], lab[order]
        W = o["W"]
        assert W == 121801 == len(win)
        for p in o["parts"]:
            lo, hi = W * p["k"] // 16, W * (p["k"] + 1) // 16
            assert hashlib.sha1(raw[lo:hi].tobytes()).hexdigest() == p["raw_sha"], p["k"]
            assert hashlib.sha1(np.ascontiguousarray(scaled[lo:hi, :]).tobytes()).hexdigest() == \
                hashlib.sha1(np.ascontiguousarray(p["scaled"][:, :gpu.D]).tobytes()).hexdigest(), p["k"]
        o_lab = np.where(o["dec"] > 0, gpu.info.label0, gpu.info.label1)
        assert np.array_equal(lab, o_lab), int((lab != o_lab).sum())
        check_decision_tolerance(load_model_arrays(model), o["scaled"][:, :gpu.D], dec, o["dec"])
    finally:
        gpu.close()
        _XYZ.pop("g512", None)


# ---- the bench's operating point ---------------------------------------------------------------------------------------
def _batch_labels(gpu, xyz_all, offsets):
    best = gpu.search_batch_packed(xyz_all, offsets)
    win = gpu.debug_windows()
    dec, lab, guard = gpu.debug_decisions()
    order = np.lexsort((win[:, 1], win[:, 0]))
    return [b.astuple() for b in best], win[order], dec[order], lab[order], guard[order], gpu.timing()


def test_bench_batch_labels_at_scale(hg, oracle_lib, tmp_models):
    """One full bench batch -- 512 synthetic clouds x 100 k points, the 2048-SV model, the DEFAULT tensor path (one
    product per k-slice is what the calibration picks for gamma = 1/323) -- against svm_mode FP64_EXACT (libsvm's own
    summation order, oracle-checked to 1e-13 elsewhere): every window's label and every cloud's best grasp must be
    identical; 8 of the clouds are also checked label by label against the oracle itself.  Also the audit's bookkeeping."""
    from haf_grasping_b200 import synth
    model = tmp_models(2048)
    n_clouds, half = 512, 256
    clouds = [synth.synth_cloud(1234 + i, 100000) for i in range(n_clouds)]
    fast = hg.GraspSearch(FEATURES, RANGE, model)
    exact = hg.GraspSearch(FEATURES, RANGE, model, svm_mode=hg.HAF_SVM_FP64_EXACT)
    mdl = load_model_arrays(model)
    check = [0, 1, 100, 255, 256, 300, 400, 511]
    for c in check:
        _XYZ[c] = clouds[c]
    try:
        assert fast.info.reserved[0] == 1, "the bench model is expected to calibrate to one tensor-core product"
        fast.set_debug(True)
        exact.set_debug(True)
        with _pool() as pool:
            orc_async = pool.map_async(_oracle_cloud, [(c, model) for c in check], chunksize=1)
            total = 0
            per_half = []
            for h0 in (0, half):
                cl = clouds[h0:h0 + half]
                off = np.concatenate([[0], np.cumsum([len(c) for c in cl])])
                import torch
                xyz_all = torch.from_numpy(np.concatenate(cl)).cuda()   # device-resident: one pass over the stage sequence
                bf, wf, df, lf, gf, tf = _batch_labels(fast, xyz_all, off)
                be, we, de, le, ge, te = _batch_labels(exact, xyz_all, off)
                assert np.array_equal(wf, we)
                assert np.array_equal(lf, le), "labels differ from the FP64 exact-order mode in %d windows" % int((lf != le).sum())
                assert bf == be
                # decision values outside the guard band: within the stated tolerance of the exact ones (per-window scale)
                assert tf.tc_passes == 1 and tf.escalations == 0
                assert tf.n_audit > 0 and 0 < tf.audit_max_rel <= 0.25 * fast.info.reserved[1] * 1e-9
                total += len(wf)
                per_half.append((wf, df, lf, de))
            assert total > 3300000
            for key, o_dec, o_mask, o_ev, o_top, o_best, o_nw in orc_async.get():
                wf, df, lf, de = per_half[key // half]
                c = key % half
                sel = (wf[:, 0] >= c * 12) & (wf[:, 0] < (c + 1) * 12)
                assert sel.sum() == o_nw == len(o_dec)
                o_lab = np.where(o_dec > 0, fast.info.label0, fast.info.label1)
                assert np.array_equal(lf[sel], o_lab), key
                assert np.abs(de[sel] - o_dec).max() <= 1e-13 * np.abs(mdl["coef"]).sum()
    finally:
        fast.close()
        exact.close()
        for c in check:
            _XYZ.pop(c, None)


def test_audit_escalates_the_number_of_products(hg, oracle_lib, trained_model_path, clouds_npz):
    """gamma = 0.02 (the trained substitute) with ONE product forced and a guard band that is too narrow for it: the
    audit measures the contraction's error on the call's own windows, the call is repeated with more products, and the
    labels equal the oracle's.  With a band too narrow even for three products the call fails loudly."""
    xyz = clouds_npz["table1"]
    # one product on this model is off by ~4e-6 E on table1's windows (measured); a band of 8e-6 E leaves it less than 4x
    gpu = hg.GraspSearch(FEATURES, RANGE, trained_model_path, tc_passes=1, guard_rel=8e-6, audit_every=8)
    o = oracle_lib.Oracle(FEATURES, RANGE, trained_model_path)
    try:
        res = gpu.search(xyz)
        t = gpu.timing()
        assert t.escalations >= 1 and t.tc_passes >= 2, (t.escalations, t.tc_passes, t.audit_max_rel)
        assert t.audit_max_rel <= 0.25 * 8e-6
        ores = o.search(xyz, oracle_lib.make_request())
        dec, lab, _ = gpu.debug_decisions()
        win = gpu.debug_windows()
        order = np.lexsort((win[:, 1], win[:, 0]))
        assert np.array_equal(lab[order], np.where(ores["dec"] > 0, gpu.info.label0, gpu.info.label1))
        assert res["best"].astuple() == ores["best"].astuple()
        assert np.array_equal(res["graspseval"][0], ores["graspseval"])
    finally:
        gpu.close()
    gpu = hg.GraspSearch(FEATURES, RANGE, trained_model_path, guard_rel=1e-9, audit_every=8)
    try:
        with pytest.raises(hg.HafError) as e:
            gpu.search(xyz)
        assert e.value.code == -5 and "audit" in str(e.value)
    finally:
        gpu.close()


@pytest.mark.parametrize("gamma,nsv", [(0.05, 256), (0.2, 300), (1.0 / 323.0, 1000)])
def test_models_far_from_the_calibration_probes(hg, oracle_lib, tmp_models, clouds_npz, gamma, nsv):
    """the calibration probes are the model's own support vectors; real windows lie far from every SV of a
    synthetic model (uniform random SVs), and a large gamma multiplies every operand error.  Labels, scores and the best
    grasp must still equal the oracle's, decision values stay inside the stated tolerance, and the audit keeps its margin."""
    model = tmp_models(nsv, gamma=gamma)
    gpu = hg.GraspSearch(FEATURES, RANGE, model, audit_every=16)
    o = oracle_lib.Oracle(FEATURES, RANGE, model)
    mdl = load_model_arrays(model)
    try:
        for name in ("table2", "pcd7"):
            xyz = clouds_npz[name]
            res = gpu.search(xyz)
            ores = o.search(xyz, oracle_lib.make_request())
            dec, lab, guard = gpu.debug_decisions()
            win = gpu.debug_windows()
            _, scaled = gpu.debug_features(raw=False)
            order = np.lexsort((win[:, 1], win[:, 0]))
            dec, lab, scaled = dec[order], lab[order], scaled[order]
            assert np.array_equal(lab, np.where(ores["dec"] > 0, gpu.info.label0, gpu.info.label1))
            assert np.array_equal(res["graspseval"][0], ores["graspseval"])
            assert res["best"].astuple() == ores["best"].astuple()
            check_decision_tolerance(mdl, scaled, dec, ores["dec"])
            t = gpu.timing()
            assert t.audit_max_rel <= 0.25 * gpu.info.reserved[1] * 1e-9
    finally:
        gpu.close()
