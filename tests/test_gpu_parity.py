"""GPU parity: the CUDA path, called through the C ABI (libhafgpu.so), against the CPU oracle on the same inputs.

Bars (north_star): bit-exact heights / cell indices / integral images / masks / raw features / scaled SVM
inputs / graspseval / per-roll tops / best grasp; decision values within a stated tolerance; labels identical.
"""
import gzip
import json
import os

import numpy as np
import pytest

from conftest import FEATURES, GOLDEN, RANGE
from model_io import check_decision_tolerance, coefK_sum, load_model_arrays

pytestmark = pytest.mark.gpu

# north_star: decision values within <= 1e-5 relative in FP32.  A decision value is a SUM of S signed terms coef_i K_i
# that may cancel to anything, so "relative" is stated against what the rounding errors scale with, PER WINDOW (computed
# here in float64 from the oracle's scaled inputs and the model file -- not a flat model-wide constant):
#     |dec_gpu - dec_ref| <= DEC_RTOL_E * E(window),   E = sum_i |coef_i| K_i (1 + gamma log2e (|x|^2 + |sv_i|^2))
# and, for windows whose exponent arguments are moderate, |dec_gpu - dec_ref| <= 1e-5 * sum_i |coef_i| K_i
# (model_io.check_decision_tolerance).  E is the guard band's own scale; the band is 4e-6 E wide or wider.
DEC_RTOL = 1e-5      # plain form
DEC_RTOL_E = 2e-6    # E form; measured <= 5e-7 on every FP32 / tensor path (profiles/r2_final_dec_error_probe.txt)
# FP64 exact-order path: only exp() implementation differences (glibc vs CUDA, <= 1 ulp each term)
DEC64_RTOL = 1e-13


@pytest.fixture(scope="module")
def clouds():
    return np.load(os.path.join(GOLDEN, "clouds.npz"))


@pytest.fixture(scope="module")
def expected():
    with open(os.path.join(GOLDEN, "expected.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def trained_model(tmp_path_factory):
    p = str(tmp_path_factory.mktemp("trained") / "substitute_trained.model")
    with gzip.open(os.path.join(GOLDEN, "substitute_trained.model.gz"), "rb") as src, open(p, "wb") as dst:
        dst.write(src.read())
    return p


@pytest.fixture(scope="module")
def hg():
    import haf_grasping_b200 as h
    return h


class Pair:
    """GPU context + oracle on the same three files."""

    def __init__(self, hg, orc, model, range_path=RANGE, **kw):
        self.gpu = hg.GraspSearch(FEATURES, range_path, model, **kw)
        self.orc = orc.Oracle(FEATURES, range_path, model)
        self.G = kw.get("grid", 56)
        self.step = kw.get("roll_step_deg", 15)
        self.rmax = kw.get("roll_max_deg", 190)

    def close(self):
        self.gpu.close()


def mk_requests(hg, orc, **kw):
    return hg.make_request(**kw), orc.make_request(**kw)


def check_search(pair, xyz, hg, orc, model_path, dec_rtol=DEC_RTOL, full=True, **rqkw):
    grq, orq = mk_requests(hg, orc, **rqkw)
    G = pair.G
    ores = pair.orc.search(xyz, orq, G=G, roll_step_deg=pair.step, roll_max_deg=pair.rmax)
    gres = pair.gpu.search(xyz, [grq])
    ob, gb = ores["best"], gres["best"]
    nroll = ob.rolls_done
    # heights, masks, graspseval: bit-exact (== on floats; the sign of a zero height is the only freedom)
    lim = nroll  # rolls the reference loop evaluated (roll_limit / early exit); the GPU may have done more
    assert np.array_equal(gres["heights"][0][:lim], ores["heights"][:lim])
    assert np.array_equal(gres["mask"][0][:lim], ores["mask"][:lim])
    integral = pair.gpu.debug_integral(len(ores["integral"]))
    assert integral[:lim].tobytes() == ores["integral"][:lim].tobytes()
    # windows: same set; features / scaled inputs / labels per window
    win = pair.gpu.debug_windows()
    raw, scaled = pair.gpu.debug_features()
    dec, lab, guard = pair.gpu.debug_decisions()
    order = np.lexsort((win[:, 1], win[:, 0]))  # (unit, cell) row-major == the reference's file order per roll
    win, raw, scaled, dec, lab, guard = win[order], raw[order], scaled[order], dec[order], lab[order], guard[order]
    pos = 0
    scaled_rolls = []
    for roll in range(lim):
        feats_o, rc = pair.orc.calc_featurevectors(ores["integral"][roll], ores["mask"][roll])
        W = len(feats_o)
        sel = slice(pos, pos + W)
        assert (win[sel, 0] == roll).all()
        assert np.array_equal(win[sel, 1], rc[:, 0] * G + rc[:, 1])
        assert raw[sel].tobytes() == feats_o.tobytes(), "raw features differ (roll %d)" % roll
        scaled_o = pair.orc.scale(feats_o)[:, :pair.gpu.D]
        assert scaled[sel].tobytes() == scaled_o.tobytes(), "scaled SVM inputs differ (roll %d)" % roll
        scaled_rolls.append(scaled_o)
        pos += W
    if nroll == len(ores["heights"]):
        assert pos == len(win)
    if full:
        # decision values of all evaluated rolls (oracle concatenates rolls it evaluated; with early exit fewer)
        o_dec = ores["dec"]
        n = len(o_dec)
        assert n <= len(dec)
        err = np.abs(dec[:n] - o_dec)
        if n and dec_rtol >= DEC_RTOL:      # FP32 / tensor contraction
            check_decision_tolerance(load_model_arrays(model_path), np.concatenate(scaled_rolls)[:n], dec[:n], o_dec, DEC_RTOL_E, dec_rtol)
        elif n:                             # FP64 paths: against the plain per-window scale
            scale = coefK_sum(load_model_arrays(model_path), np.concatenate(scaled_rolls)[:n])
            bad = err > dec_rtol * scale + 1e-300
            assert not bad.any(), (int(bad.sum()), float((err / np.maximum(scale, 1e-300)).max()))
        o_lab = np.where(o_dec > 0, pair.gpu.info.label0, pair.gpu.info.label1)
        assert np.array_equal(lab[:n], o_lab), "labels differ"
        assert np.array_equal(gres["graspseval"][0][:nroll], ores["graspseval"][:nroll])
        assert np.array_equal(gres["per_roll_top"][0][:nroll], ores["per_roll_top"][:nroll])
    assert gb.astuple() == ob.astuple(), (gb.astuple(), ob.astuple())
    assert gb.eval == ob.eval and gb.rolls_done == ob.rolls_done and gb.n_windows_scored == ob.n_windows
    assert abs(gb.roll_rad - ob.roll_rad) == 0
    return gres, ores, guard


@pytest.fixture(scope="module")
def pair_trained(hg, oracle_lib, trained_model):
    p = Pair(hg, oracle_lib, trained_model)
    yield p
    p.close()


@pytest.fixture(scope="module")
def pair_synth(hg, oracle_lib, tmp_models):
    """synthetic 256-SV model on the FP32 SIMT mode (the bundled-PCD / trained-model tests run the default tensor mode)"""
    p = Pair(hg, oracle_lib, tmp_models(256), svm_mode=hg.HAF_SVM_FP32_GUARD)
    yield p
    p.close()


ALL_CLOUDS = ["pcd1", "pcd2", "pcd3", "pcd4", "pcd5", "pcd6", "pcd7", "pcd8", "pcd9", "pcd10", "pcd11", "pcd12",
              "plastic_mug2", "table1", "table2", "table3"]


@pytest.mark.parametrize("name", ALL_CLOUDS)
def test_bundled_pcd_trained_model(pair_trained, clouds, expected, hg, oracle_lib, trained_model, name):
    gres, ores, _ = check_search(pair_trained, clouds[name], hg, oracle_lib, trained_model)
    exp = expected["trained/" + name]
    assert list(gres["best"].astuple()) == exp["best"]
    assert gres["best"].n_windows_scored == exp["n_windows"]
    assert gres["per_roll_top"][0].tolist() == exp["per_roll_top"]


@pytest.mark.parametrize("name", ["pcd2", "pcd7", "plastic_mug2", "table1", "table3"])
def test_bundled_pcd_synth_model(pair_synth, clouds, expected, hg, oracle_lib, tmp_models, name):
    gres, _, _ = check_search(pair_synth, clouds[name], hg, oracle_lib, tmp_models(256))
    assert list(gres["best"].astuple()) == expected["synth256/" + name]["best"]


@pytest.mark.parametrize("name,roll", [("pcd2", 0), ("table1", 4), ("table2", 11), ("plastic_mug2", 7)])
def test_cell_indices_bit_exact(pair_synth, clouds, hg, oracle_lib, name, roll):
    xyz = clouds[name]
    grq, _ = mk_requests(hg, oracle_lib)
    cells = pair_synth.gpu.debug_cell_indices(xyz, grq, roll)
    o = pair_synth.orc
    M = o.build_transform((0, 0, 0), o.normalize_approach((0, 0, 1)), 1, roll)
    _, ocells, clamped = o.generate_grid(xyz, M, want_cells=True)
    assert clamped == 0
    assert np.array_equal(cells, ocells)


def test_device_text_roundtrips_match_glibc(pair_synth):
    rng = np.random.default_rng(3)
    f = np.concatenate([
        rng.integers(0, 2 ** 32, 200000, dtype=np.uint64).astype(np.uint32).view(np.float32),
        (rng.uniform(-1, 1, 200000) * 10.0 ** rng.integers(-8, 5, 200000)).astype(np.float32),
        np.array([0.0, -0.0, 1.0, 9999.5, 99995.0, 0.00012345, 1e-20, 3e38, 1e-45, 12345.0, 0.5, 1234.5, 1235.5], np.float32)])
    f = f[np.isfinite(f)]
    v = np.concatenate([rng.uniform(-1.5, 1.5, 300000), rng.uniform(-1, 1, 100000) * 10.0 ** rng.integers(-30, 30, 100000),
                        np.array([0.0, 1.0, -1.0, 0.1234565, 1234565.0, 0.9999995, 1e-7, 123456.5, 1e22])])
    o4, o6 = pair_synth.gpu.debug_text_roundtrip(f, v)
    r4 = np.array([float("%.4g" % float(x)) for x in f])
    r6 = np.array([float("%g" % float(x)) for x in v])
    assert o4.tobytes() == r4.tobytes()
    assert o6.tobytes() == r6.tobytes()


def test_fp64_exact_mode(hg, oracle_lib, trained_model, clouds):
    p = Pair(hg, oracle_lib, trained_model, svm_mode=hg.HAF_SVM_FP64_EXACT)
    try:
        check_search(p, clouds["pcd2"], hg, oracle_lib, trained_model, dec_rtol=DEC64_RTOL)
        check_search(p, clouds["table2"], hg, oracle_lib, trained_model, dec_rtol=DEC64_RTOL)
    finally:
        p.close()


@pytest.mark.parametrize("name", ["pcd2", "table3"])
def test_tensor_core_mode_streaming_pair_one_product(hg, oracle_lib, tmp_models, clouds, name):
    """tc_variant = 2 on the bench model (2048 SVs, gamma = 1/323: one product per k-slice): the streaming CTA-pair kernel on
    the path the X-resident kernel normally takes."""
    model = tmp_models(2048)
    p = Pair(hg, oracle_lib, model, svm_mode=hg.HAF_SVM_TENSOR_GUARD, tc_variant=2)
    try:
        assert p.gpu.info.reserved[0] == 1
        check_search(p, clouds[name], hg, oracle_lib, model)
    finally:
        p.close()


@pytest.mark.parametrize("name", ["pcd2", "pcd7", "table1", "plastic_mug2"])
def test_tensor_core_mode(hg, oracle_lib, trained_model, clouds, name):
    """HAF_SVM_TENSOR_GUARD: tcgen05 split-fp16 contraction + FP64 guard band; everything before and after the
    decision values is shared with the SIMT mode, labels / evals / best grasp must equal the oracle's."""
    p = Pair(hg, oracle_lib, trained_model, svm_mode=hg.HAF_SVM_TENSOR_GUARD)
    try:
        check_search(p, clouds[name], hg, oracle_lib, trained_model)
    finally:
        p.close()


def test_tensor_mode_fast_tier_inputs_track_the_exact_scaled_values(hg, oracle_lib, trained_model, clouds):
    """The tensor path's operands come from a fast tier (double-precision "%.4g" decision, float scaling, no "%g" step,
    fp16 hi + lo split = 22 bits).  They must sit within the skipped 6-digit rounding (<= 5e-6 relative) plus the float
    scaling error (a few 1e-7 absolute: svm-scale's -1 + 2 (v - min) / (max - min) cancels) of the exact scaled values;
    only a mis-decided decimal near-tie may differ by one 4th-digit step, and none is expected on this data."""
    gpu = hg.GraspSearch(FEATURES, RANGE, trained_model, svm_mode=hg.HAF_SVM_TENSOR_GUARD)
    try:
        gpu.search(clouds["table1"])
        x = gpu.debug_tensor_inputs().astype(np.float64)
        _, scaled = gpu.debug_features(raw=False)       # bit-exact emulation tier (checked against the oracle elsewhere)
        err = np.abs(x - scaled)
        tol = 5.5e-6 * np.abs(scaled) + 5e-7
        assert (err <= tol).mean() == 1.0, (err.max(), (err > tol).sum())
    finally:
        gpu.close()


@pytest.mark.parametrize("kw", [{}, {"sv_table_global": 1}, {"tc_variant": 2}, {"tc_variant": 2, "sv_table_global": 1},
                                {"tc_passes": 1}, {"tc_passes": 2}, {"tc_passes": 3},
                                {"tc_passes": 1, "tc_variant": 2}, {"tc_passes": 1, "sv_table_global": 1}])
def test_tensor_core_mode_synth_models_and_batch(hg, oracle_lib, tmp_models, kw):
    """X-resident CTA-pair (one product, tc_variant 0) and streaming CTA-pair (tc_variant 2 / more products) kernels, with the coef table staged in shared memory (default) or read from global memory (what models with > 4096
    support vectors get), and with 1, 2 or 3 tensor-core products per k-slice forced (the default calibrates the count
    per model and widens the guard band by the calibrated operand error)."""
    from haf_grasping_b200 import synth
    for nsv in (256, 300):   # 300: support-vector count not a multiple of the 256-wide tile
        model = tmp_models(nsv)
        p = Pair(hg, oracle_lib, model, svm_mode=hg.HAF_SVM_TENSOR_GUARD, **kw)
        try:
            cl = [synth.synth_cloud(4321 + i, 15000 + 211 * i) for i in range(5)]
            best = p.gpu.search_batch(cl)
            for i, c in enumerate(cl):
                ob = p.orc.search(c, oracle_lib.make_request(), full=False)["best"]
                assert best[i].astuple() == ob.astuple()
            check_search(p, cl[0], hg, oracle_lib, model)
        finally:
            p.close()


@pytest.mark.parametrize("gk", [0, 2])    # tier-2 kernel: FP64 tensor cores (DMMA, the default), DFMA register tiles
@pytest.mark.parametrize("tier2", [0, 1, 2])
@pytest.mark.parametrize("mode", ["tensor", "simt"])
def test_guard_band_catches_near_zero_decisions(hg, oracle_lib, tmp_models, clouds, mode, tier2, gk):
    """rho chosen so that many decision values sit next to 0: labels must still equal the oracle's.
    tier2 = 0: guard windows are settled by the FP64 FMA tier (none should need the exact-order kernels);
    1: FMA tier off, all of them go through the exact-order kernels; 2: both tiers run on every guard window."""
    model = tmp_models(256, rho=-0.2972253)  # a decision value of pcd2 / roll 0 with the synth model is -0.2972253033...
    if gk == 2 and tier2 == 1:
        pytest.skip("tier 2 off: no kernel to choose")
    p = Pair(hg, oracle_lib, model, guard_rel=1e-3, guard_tier2=tier2, guard_kernel=gk,
             svm_mode=hg.HAF_SVM_TENSOR_GUARD if mode == "tensor" else hg.HAF_SVM_FP32_GUARD)
    try:
        # tier 2 alone leaves FMA-order decision values (<= 1e-12 relative) on the guard windows; the exact-order
        # kernels leave libsvm's order (only exp() differs)
        _, _, guard = check_search(p, clouds["pcd2"], hg, oracle_lib, model)
        assert guard.sum() > 0
        t = p.gpu.timing()
        assert t.n_guard == guard.sum()
        assert t.n_exact == (0 if tier2 == 0 else t.n_guard)   # the audit sample's windows are measured, never rewritten or escalated
    finally:
        p.close()


@pytest.mark.parametrize("gk", [1, 2])    # FP64 tensor cores (DMMA: what batches use), DFMA register tiles (single goals)
def test_guard_tier2_decision_values_match_fp64_order(hg, oracle_lib, tmp_models, clouds, gk):
    """With a guard band so wide that EVERY window is re-evaluated by the FP64 contraction tier, all decision values must sit
    within 1e-12 * sum|coef| of the oracle's libsvm-order values (the bound tier 2's own escalation test relies on)."""
    model = tmp_models(256)
    p = Pair(hg, oracle_lib, model, guard_rel=1e6, svm_mode=hg.HAF_SVM_TENSOR_GUARD, guard_kernel=gk)
    try:
        _, _, guard = check_search(p, clouds["pcd3"], hg, oracle_lib, model, dec_rtol=1e-12)
        assert guard.all()
        assert p.gpu.timing().n_exact == 0
    finally:
        p.close()


def test_tensor_mode_inputs_beyond_fp16_range_take_the_exact_path(hg, oracle_lib, tmp_models, clouds, tmp_path):
    """A range file whose span for one feature is tiny scales that feature far beyond +-65504 (svm-scale does not clamp,
    svm-scale.c:339-346).  The tensor path clamps its fp16 operands and must send every such window to the FP64 exact
    path: labels, scores and the best grasp still equal the oracle's."""
    lines = open(RANGE).read().split("\n")
    idx, lo, hi = lines[2 + 4].split()          # 5th feature line ("idx min max")
    lines[2 + 4] = "%s %s %.12g" % (idx, lo, float(lo) + 1e-7)
    rp = str(tmp_path / "narrow.range")
    open(rp, "w").write("\n".join(lines))
    model = tmp_models(256)
    p = Pair(hg, oracle_lib, model, range_path=rp, svm_mode=hg.HAF_SVM_TENSOR_GUARD)
    try:
        _, _, guard = check_search(p, clouds["pcd2"], hg, oracle_lib, model)
        win = p.gpu.debug_windows()
        _, scaled = p.gpu.debug_features(raw=False)
        scaled = scaled[np.lexsort((win[:, 1], win[:, 0]))]     # same order as the guard flags check_search returns
        big = (np.abs(scaled) > 65000.0).any(axis=1)
        assert big.sum() > 0, "the narrowed range did not push any value beyond the fp16 range"
        assert guard[big].all()
    finally:
        p.close()


def test_tensor_passes_are_calibrated_per_model(hg, oracle_lib, tmp_models, trained_model, clouds):
    """gamma = 1/323 (libsvm's default) with 2048 support vectors, the bench model: the cross terms are below the FP32
    epilogue error, the calibration drops them and widens the guard band; the trained substitute (gamma = 0.02) and the
    small synthetic models keep all three products.  The one-product search still equals the oracle's."""
    seen = {}
    for name, model in (("synth2048", tmp_models(2048)), ("synth256", tmp_models(256)), ("trained", trained_model)):
        g = hg.GraspSearch(FEATURES, RANGE, model)
        seen[name] = (g.info.reserved[0], g.info.reserved[1] * 1e-9)
        g.close()
    assert seen["trained"][0] == 3 and abs(seen["trained"][1] - 4e-6) < 1e-9
    assert seen["synth256"][0] == 3
    assert seen["synth2048"][0] == 1 and 4e-6 < seen["synth2048"][1] <= 1.2e-5, seen
    p = Pair(hg, oracle_lib, tmp_models(2048))
    try:
        check_search(p, clouds["pcd5"], hg, oracle_lib, tmp_models(2048))
    finally:
        p.close()


def test_label_order_minus_plus(hg, oracle_lib, tmp_models, clouds):
    model = tmp_models(256, labels=(-1, 1), rho=0.01)
    p = Pair(hg, oracle_lib, model)
    try:
        check_search(p, clouds["pcd3"], hg, oracle_lib, model)
    finally:
        p.close()


def test_no_text_emulation_mode(hg, oracle_lib, tmp_models, clouds):
    model = tmp_models(256)
    gpu = hg.GraspSearch(FEATURES, RANGE, model, emulate_text_roundtrip=False)
    o = oracle_lib.Oracle(FEATURES, RANGE, model)
    try:
        res = gpu.search(clouds["pcd2"])
        ores = o.search(clouds["pcd2"], oracle_lib.make_request(), emulate_text=False)
        raw, scaled = gpu.debug_features()
        win = gpu.debug_windows()
        order = np.lexsort((win[:, 1], win[:, 0]))
        feats_o, _ = o.calc_featurevectors(ores["integral"][0], ores["mask"][0])
        sc_o = o.scale(feats_o, emulate_text=False)[:, :gpu.D]
        W0 = len(feats_o)
        assert scaled[order][:W0].tobytes() == sc_o.tobytes()
        assert res["best"].astuple() == ores["best"].astuple()
    finally:
        gpu.close()


def test_request_variants(pair_synth, clouds, hg, oracle_lib, tmp_models):
    m = tmp_models(256)
    xyz = clouds["table1"]
    check_search(pair_synth, xyz, hg, oracle_lib, m, center=(0.1, 0.25, 0.0), area=(30.0, 40.0), full=True)
    check_search(pair_synth, xyz, hg, oracle_lib, m, width=2)
    check_search(pair_synth, xyz, hg, oracle_lib, m, roll_limit=5)
    check_search(pair_synth, xyz, hg, oracle_lib, m, return_only_best=1, graspval_top=60)
    check_search(pair_synth, xyz, hg, oracle_lib, m, area=(14.9, 20.0))  # height_r = 0
    check_search(pair_synth, np.zeros((0, 3), np.float32), hg, oracle_lib, m)  # empty cloud


def test_padded_stride_and_device_pointer(pair_synth, clouds, hg):
    import torch
    xyz = clouds["pcd2"]
    ref = pair_synth.gpu.search(xyz)["best"].astuple()
    padded = np.zeros((len(xyz), 4), np.float32)  # pcl::PointXYZ layout (16-byte stride)
    padded[:, :3] = xyz
    padded[:, 3] = 1.0
    assert pair_synth.gpu.search(padded, stride_bytes=16)["best"].astuple() == ref
    dev = torch.from_numpy(xyz).cuda()
    assert pair_synth.gpu.search(dev)["best"].astuple() == ref


TILTED = [(0.0, 0.0, 1.0), (0.5, 0.0, 0.8660254), (-0.5, 0.0, 0.8660254), (0.0, 0.5, 0.8660254), (0.0, -0.5, 0.8660254)]


def test_approach_vectors(pair_synth, clouds, hg, oracle_lib):
    """config 3: five approach vectors in one call == five oracle goals; overall winner = strict >, earliest."""
    xyz = clouds["table2"]
    greqs = [hg.make_request(approach=a) for a in TILTED]
    gres = pair_synth.gpu.search(xyz, greqs)
    tops = []
    for i, a in enumerate(TILTED):
        ores = pair_synth.orc.search(xyz, oracle_lib.make_request(approach=a))
        assert gres["best_per_request"][i].astuple() == ores["best"].astuple()
        assert np.array_equal(gres["heights"][i], ores["heights"])
        assert np.array_equal(gres["mask"][i], ores["mask"])
        assert np.array_equal(gres["graspseval"][i], ores["graspseval"])
        assert np.array_equal(gres["per_roll_top"][i], ores["per_roll_top"])
        tops.append(ores["best"].topval)
    win = int(np.argmax(tops))  # argmax returns the first maximum
    assert gres["best"].approach_idx == win
    assert gres["best"].astuple() == gres["best_per_request"][win].astuple()


def test_batch_matches_single(pair_synth, hg, oracle_lib):
    from haf_grasping_b200 import synth
    cl = [synth.synth_cloud(1234 + i, 20000 + 137 * i) for i in range(6)]
    best = pair_synth.gpu.search_batch(cl)
    off = np.concatenate([[0], np.cumsum([len(c) for c in cl])])
    best2 = pair_synth.gpu.search_batch_packed(np.concatenate(cl), off)
    for i, c in enumerate(cl):
        ob = pair_synth.orc.search(c, oracle_lib.make_request(), full=False)["best"]
        assert best[i].astuple() == ob.astuple()
        assert best2[i].astuple() == ob.astuple()
        assert best[i].n_windows_scored == ob.n_windows


def test_host_staged_batch_in_chunks_on_several_streams(hg, oracle_lib, tmp_models, monkeypatch):
    """A batch handed over in HOST memory is staged piecewise and processed in chunks of 16, 24, 32, ... clouds that alternate
    between three streams, each with its own set of per-chunk buffers and counters (hafgpu.cu, haf_ctx::ChunkWs).  The results, the window / guard
    totals and the audit bookkeeping must equal the one-pass run on the same clouds resident in device memory and the
    one-stream chunked run (HAF_DUAL_STREAM=0); spot checks against the oracle.  Repeated: buffers are reused across calls."""
    import torch
    from haf_grasping_b200 import synth
    model = tmp_models(256)
    cl = [synth.synth_cloud(4321 + i, 9000 + 211 * (i % 7)) for i in range(90)]     # 16 + 24 + 32 + 18 clouds: four chunks
    off = np.concatenate([[0], np.cumsum([len(c) for c in cl])])
    xyz = np.concatenate(cl)
    gpu = hg.GraspSearch(FEATURES, RANGE, model)
    try:
        res = gpu.search_batch_packed(torch.from_numpy(xyz).cuda(), off)
        t_res = gpu.timing()
        want = [(b.astuple(), b.n_windows_scored, b.rolls_done) for b in res]
        assert t_res.n_chunks == 1
        for rep in range(3):
            got = gpu.search_batch_packed(xyz, off)
            t = gpu.timing()
            assert t.n_chunks == 4
            assert [(b.astuple(), b.n_windows_scored, b.rolls_done) for b in got] == want
            assert (t.n_windows, t.n_guard, t.n_exact) == (t_res.n_windows, t_res.n_guard, t_res.n_exact)
            assert t.n_audit > 0 and 0 < t.audit_max_rel <= 0.25 * gpu.info.reserved[1] * 1e-9
        monkeypatch.setenv("HAF_DUAL_STREAM", "0")
        one = gpu.search_batch_packed(xyz, off)
        t1 = gpu.timing()
        assert [(b.astuple(), b.n_windows_scored, b.rolls_done) for b in one] == want
        assert (t1.n_chunks, t1.n_windows, t1.n_guard) == (4, t_res.n_windows, t_res.n_guard)
        assert t1.n_audit > 0 and 0 < t1.audit_max_rel <= 0.25 * gpu.info.reserved[1] * 1e-9   # (the sample follows the compaction order: not fixed)
        assert t1.launches + 1 == t.launches        # the two-stream run adds the counter merge, nothing else
        monkeypatch.delenv("HAF_DUAL_STREAM")
        # (experiment knob) a device-resident batch split into two halves on two of the streams
        monkeypatch.setenv("HAF_RESIDENT_SPLIT", "2")
        halves = gpu.search_batch_packed(torch.from_numpy(xyz).cuda(), off)
        assert [(b.astuple(), b.n_windows_scored, b.rolls_done) for b in halves] == want
        assert (gpu.timing().n_chunks, gpu.timing().n_windows, gpu.timing().n_guard) == (2, t_res.n_windows, t_res.n_guard)
        monkeypatch.delenv("HAF_RESIDENT_SPLIT")
        # single goals in between use the first set of buffers only
        o = oracle_lib.Oracle(FEATURES, RANGE, model)
        for i in (0, 17, 41, 89):
            ob = o.search(cl[i], oracle_lib.make_request(), full=False)["best"]
            assert want[i][0] == ob.astuple() and want[i][1] == ob.n_windows
            assert gpu.search(cl[i], outputs=False)["best"].astuple() == ob.astuple()
        got = gpu.search_batch_packed(xyz, off)
        assert [(b.astuple(), b.n_windows_scored, b.rolls_done) for b in got] == want
    finally:
        gpu.close()


def test_large_grid_path(hg, oracle_lib, tmp_models):
    """G = 192 exercises the two-kernel integral image and multi-CTA mask compaction."""
    from haf_grasping_b200 import synth
    model = tmp_models(64, seed=11)
    p = Pair(hg, oracle_lib, model, grid=192, roll_step_deg=45, roll_max_deg=190)
    try:
        xyz = synth.synth_cloud(77, 150000, r=0.96)
        check_search(p, xyz, hg, oracle_lib, model, area=(60.0, 52.0))
    finally:
        p.close()


def test_create_errors(hg, tmp_models, tmp_path):
    with pytest.raises(hg.HafError) as e:
        hg.GraspSearch(FEATURES, RANGE, str(tmp_path / "missing.model"))
    assert e.value.code == -2
    bad = tmp_path / "lin.model"
    bad.write_text("svm_type c_svc\nkernel_type linear\nnr_class 2\ntotal_sv 1\nrho 0\nlabel 1 -1\nnr_sv 1 0\nSV\n1 1:0.5 \n")
    with pytest.raises(hg.HafError) as e:
        hg.GraspSearch(FEATURES, RANGE, str(bad))
    assert e.value.code == -5


def test_cpp_host_cli_matches_oracle_grasp_output(hg, oracle_lib, trained_model, clouds, tmp_path):
    """haf_cli = C++ host mirror (csrc/host/calc_grasppoints_b200.hpp: read_pc_cb -> loop_control ->
    transform_gp_in_wcs_and_publish) on the C ABI.  GraspOutput fields vs the oracle's restatement of
    server.cpp:1274-1401."""
    import json
    import subprocess
    from haf_grasping_b200 import build
    cli = build.build_cli()
    o = oracle_lib.Oracle(FEATURES, RANGE, trained_model)
    for name, kw in (("pcd2", {}), ("table1", {"approach": (0.5, 0.0, 0.8660254), "center": (0.05, 0.2, 0.0)}), ("pcd7", {})):
        xyz = clouds[name]
        f = tmp_path / (name + ".pcd")
        with open(f, "wb") as fh:
            fh.write(("# .PCD v0.7\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH %d\nHEIGHT 1\n"
                      "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (len(xyz), len(xyz))).encode())
            fh.write(np.ascontiguousarray(xyz, np.float32).tobytes())
        args = [cli, "--features", FEATURES, "--range", RANGE, "--model", trained_model, "--pcd", str(f), "--json"]
        if "approach" in kw:
            args += ["--approach"] + [repr(v) for v in kw["approach"]]
        if "center" in kw:
            args += ["--center"] + [repr(v) for v in kw["center"]]
        got = json.loads(subprocess.run(args, capture_output=True, text=True, check=True).stdout)
        orq = oracle_lib.make_request(**kw)
        ores = o.search(xyz, orq)
        b = ores["best"]
        assert (got["row"], got["col"], got["roll_index"], got["topval"], got["eval"]) == (b.row, b.col, b.roll, b.topval, b.eval)
        assert got["per_roll_top"] == ores["per_roll_top"].tolist()
        pose = o.transform_gp_in_wcs(orq, ores["heights"][max(b.roll, 0)], b.row, b.col, b.roll)
        mine = np.array(got["graspPoint1"] + got["graspPoint2"] + got["averagedGraspPoint"] + got["approachVector"] + [got["roll"]])
        assert np.allclose(mine, pose, rtol=0, atol=2e-6), (mine, pose)


def test_roll_begin_shards_one_goal(pair_synth, clouds, hg, oracle_lib):
    """rolls [5, 9) only: per-roll tops equal the full run's, the others read (-1, -1, -1000)."""
    xyz = clouds["table3"]
    full = pair_synth.gpu.search(xyz, [hg.make_request()])
    part = pair_synth.gpu.search(xyz, [hg.make_request(roll_begin=5, roll_limit=9)])
    assert np.array_equal(part["per_roll_top"][0][5:9], full["per_roll_top"][0][5:9])
    rest = np.concatenate([part["per_roll_top"][0][:5], part["per_roll_top"][0][9:]])
    assert (rest == np.array([-1, -1, -1000])).all()      # rolls not evaluated
    assert np.array_equal(part["graspseval"][0][5:9], full["graspseval"][0][5:9])
    assert part["best"].rolls_done == 4


def test_full_size_batch_properties(hg, oracle_lib, tmp_models):
    """BASELINE-size inputs (100k-point clouds, 2048 SVs), checked through size-independent properties:
    (1) a batch equals the same clouds searched one by one, (2) the result does not depend on the order of the points
    (max-z binning is order-free), (3) repeating a cloud inside the batch repeats its result, (4) a spot-checked
    cloud equals the oracle."""
    from haf_grasping_b200 import synth
    model = tmp_models(2048)
    gpu = hg.GraspSearch(FEATURES, RANGE, model, svm_mode=hg.HAF_SVM_TENSOR_GUARD)
    try:
        rng = np.random.default_rng(9)
        cl = [synth.synth_cloud(1234 + i, 100000) for i in range(12)]
        cl.append(cl[3].copy())                          # duplicate
        cl.append(cl[5][rng.permutation(len(cl[5]))])    # same points, shuffled
        best = gpu.search_batch(cl)
        tup = [b.astuple() for b in best]
        assert tup[12] == tup[3] and tup[13] == tup[5]
        for i in (0, 5, 11):
            assert gpu.search(cl[i], outputs=False)["best"].astuple() == tup[i]
        dev_windows = [b.n_windows_scored for b in best]
        assert dev_windows[12] == dev_windows[3] and dev_windows[13] == dev_windows[5]
        o = oracle_lib.Oracle(FEATURES, RANGE, model)
        ob = o.search(cl[0], oracle_lib.make_request(), full=False)["best"]
        assert tup[0] == ob.astuple() and dev_windows[0] == ob.n_windows
    finally:
        gpu.close()


def test_cuda_graph_replay_equals_plain_launches(hg, oracle_lib, tmp_models, clouds):
    """One goal = ~25 dependent stream operations; with cfg.reserved[1] bit 4 a call that repeats the previous call's shape is captured
    into a CUDA graph once and replayed (server.cpp:335-402 runs goal after goal with the same parameters).  Replays must return exactly what
    plain launches return -- for the same cloud, for other clouds of the same 64 k-point bucket, for other requests of the same
    shape -- and a change of shape must fall back and re-capture."""
    model = tmp_models(256)
    g = hg.GraspSearch(FEATURES, RANGE, model, use_graph=True)      # graphs on (opt-in; one table1 goal: 0.224 ms launched, 0.189 ms replayed)
    p = hg.GraspSearch(FEATURES, RANGE, model)                      # plain launches (the default)
    try:
        def same(a, b):
            assert a["best"].astuple() == b["best"].astuple()
            assert (a["best"].rolls_done, a["best"].n_windows_scored) == (b["best"].rolls_done, b["best"].n_windows_scored)
            for k in ("graspseval", "mask", "heights", "per_roll_top"):
                assert np.array_equal(a[k], b[k]), k
        seq = [("pcd2", {}), ("pcd2", {}), ("pcd2", {}), ("pcd7", {}), ("plastic_mug2", {}), ("pcd2", {"approach": (0.5, 0.0, 0.8660254)}),
               ("pcd3", {"center": (0.01, -0.02, 0.0)}), ("pcd2", {"return_only_best": 1, "graspval_top": 60})]
        for name, kw in seq:                                         # all: 12 units, same window bound, clouds < 65 536 points
            same(g.search(clouds[name], [hg.make_request(**kw)]), p.search(clouds[name], [hg.make_request(**kw)]))
        assert g.timing().graph_replays >= 1 and p.timing().graph_replays == 0   # (a bigger cloud re-allocates the staging buffer: re-capture)
        assert g.timing().launches == p.timing().launches
        n0 = g.timing().graph_replays
        # other shapes: a bigger cloud (another bucket), two requests, a narrower area, a roll range -- each captured anew after one plain call
        for name, reqs in (("table1", [hg.make_request()]), ("pcd2", [hg.make_request(), hg.make_request(approach=(0.0, 0.5, 0.8660254))]),
                           ("pcd2", [hg.make_request(area=(28.0, 36.0))]), ("pcd2", [hg.make_request(roll_begin=2, roll_limit=9)])):
            for _ in range(3):
                same(g.search(clouds[name], reqs), p.search(clouds[name], reqs))
        assert g.timing().graph_replays >= n0 + 4
        # several graphs are kept (least recently used out): a goal sharded over ranks alternates between a few shapes
        shapes = ([hg.make_request()], [hg.make_request(roll_begin=2, roll_limit=9)], [hg.make_request(roll_begin=0, roll_limit=6)])
        for _ in range(2):
            for reqs in shapes:
                same(g.search(clouds["pcd2"], reqs), p.search(clouds["pcd2"], reqs))
        n1 = g.timing().graph_replays
        for _ in range(3):
            for reqs in shapes:
                same(g.search(clouds["pcd2"], reqs), p.search(clouds["pcd2"], reqs))
        assert g.timing().graph_replays == n1 + 9
        # batches of one shape replay too; device-resident clouds and outputs=False (no per-roll copies)
        from haf_grasping_b200 import synth
        cl = [synth.synth_cloud(40 + i, 30000) for i in range(5)]
        for _ in range(3):
            assert [b.astuple() for b in g.search_batch(cl)] == [b.astuple() for b in p.search_batch(cl)]
        import torch
        dev = torch.from_numpy(clouds["table2"]).cuda()
        for _ in range(3):
            a = g.search(dev, [hg.make_request()], outputs=False)
            b = p.search(clouds["table2"], [hg.make_request()], outputs=False)
            assert a["best"].astuple() == b["best"].astuple() and np.array_equal(a["per_roll_top"], b["per_roll_top"])
        # the debug accessors read the state a replay left behind
        g.search(clouds["pcd2"]); g.search(clouds["pcd2"]); g.search(clouds["pcd2"])
        p.search(clouds["pcd2"])
        dg, lg, _ = g.debug_decisions()
        dp, lp, _ = p.debug_decisions()
        wg, wp = g.debug_windows(), p.debug_windows()
        og, op = np.lexsort((wg[:, 1], wg[:, 0])), np.lexsort((wp[:, 1], wp[:, 0]))   # (the compaction order across units is not fixed)
        assert np.array_equal(wg[og], wp[op]) and np.array_equal(lg[og], lp[op])
    finally:
        g.close()
        p.close()
