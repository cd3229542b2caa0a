"""Probability estimates: `svm-predict -b 1` numerics and the server's svm_with_probability branch.

Reference: libsvm-3.12/svm.cpp:1818-1826 (sigmoid_predict), :1829-1890 (multiclass_probability), :2550-2590
(svm_predict_probability), svm-predict.c:53-66 / :111-118 (the output file) and
src/calc_grasppoints_action_server.cpp:817-818, :831-846 (show_predicted_gps parsing "<label> <p0> <p1>" lines -- one line
late, because the header line is read as the first prediction).

* CPU (`ref`): the oracle's restatement against the reference's OWN svm-predict -b 1 (byte-identical output files) and against
  the reference's OWN show_predicted_gps(roll, tilt, true) inside the action server compiled in place.
* GPU: haf_svm_predict_probability / svm-predict-b200 -b 1 / haf_search(svm_with_probability = 1) against the oracle, against
  committed output of the reference program, and (where oracle/_ref travelled) against the reference program itself.
The probability model = the committed substitute model + the probA / probB lines the reference's svm-train -b 1 produced for
the same training file (tests/golden/make_prob_golden.py).
"""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import FEATURES, GOLDEN, RANGE, ROOT

LIBDIR = os.path.join(ROOT, "haf_grasping_b200", "lib")
SVM_PREDICT_B200 = os.path.join(LIBDIR, "svm-predict-b200")
CLI_DIR = os.path.join(GOLDEN, "svm_cli")


@pytest.fixture(scope="module")
def prob_model(tmp_path_factory):
    with open(os.path.join(GOLDEN, "substitute_trained.prob.json")) as fh:
        pj = json.load(fh)
    text = gzip.open(os.path.join(GOLDEN, "substitute_trained.model.gz"), "rb").read().decode()
    head, tail = text.split("nr_sv", 1)   # libsvm writes probA / probB between `label` and `nr_sv` (svm.cpp:2640-2655)
    p = str(tmp_path_factory.mktemp("prob_model") / "substitute_trained.prob.model")
    with open(p, "w") as fh:
        fh.write(head + "probA %s\nprobB %s\nnr_sv" % (pj["probA"], pj["probB"]) + tail)
    return p


@pytest.fixture(scope="module")
def clouds():
    return np.load(os.path.join(GOLDEN, "clouds.npz"))


def _dense_rows_of_file(path, dim):
    rows = []
    with open(path) as fh:
        for ln in fh:
            tok = ln.split()
            x = np.zeros(dim)
            for t in tok[1:]:
                k, v = t.split(":")
                x[int(k) - 1] = float(v)
            rows.append(x)
    return np.array(rows)


def _roll_scaled(o, xyz, roll):
    M = o.build_transform((0, 0, 0), o.normalize_approach((0, 0, 1)), 1, roll)
    integral = o.calc_intimage(o.generate_grid(xyz, M))
    mask = o.pnt_in_box(integral, roll)
    feats, _ = o.calc_featurevectors(integral, mask)
    return o.scale(feats), mask


def _write_libsvm(path, scaled):
    with open(path, "w") as fh:
        for row in scaled:   # what svm-scale writes: "%g" of every non-zero value, a blank after each pair (svm-scale.c:348-352)
            fh.write("0 " + "".join("%d:%s " % (k + 1, "%g" % v) for k, v in enumerate(row) if v != 0) + "\n")


# ------------------------------------------------------------------------------------------------------- CPU: oracle pins
def test_oracle_reproduces_the_committed_reference_output(oracle_lib, prob_model):
    """runs everywhere (no oracle/_ref needed): the reference program's own -b 1 output for the committed 20-row file"""
    o = oracle_lib.Oracle(FEATURES, RANGE, prob_model)
    assert o.check_probability_model()
    x = _dense_rows_of_file(os.path.join(CLI_DIR, "scaled_ref.txt"), 323)
    lab, pr = o.svm_predict_probability(x)
    with open(os.path.join(CLI_DIR, "prob_out_trained_ref.txt"), "rb") as fh:
        assert o.format_probability_output(lab, pr) == fh.read()


def test_oracle_rejects_probability_without_probA_probB(oracle_lib, trained_model_path):
    o = oracle_lib.Oracle(FEATURES, RANGE, trained_model_path)
    assert not o.check_probability_model()
    with pytest.raises(ValueError):
        o.svm_predict_probability(np.zeros((1, 323)))


@pytest.mark.ref
@pytest.mark.parametrize("name,roll", [("pcd2", 0), ("table1", 5), ("plastic_mug2", 9)])
def test_oracle_probability_text_equals_reference_svm_predict(oracle_lib, prob_model, clouds, tmp_path, name, roll):
    """one roll's scaled feature file through the REFERENCE'S OWN svm-predict -b 1 (child process): byte-identical output"""
    o = oracle_lib.Oracle(FEATURES, RANGE, prob_model)
    scaled, _ = _roll_scaled(o, clouds[name], roll)
    assert len(scaled) > 50
    f_in, f_out = str(tmp_path / "in.scale"), str(tmp_path / "out.txt")
    _write_libsvm(f_in, scaled)
    subprocess.run([os.path.join(oracle_lib.REF_DIR, "svm-predict"), "-b", "1", f_in, prob_model, f_out], check=True, capture_output=True)
    lab, pr = o.svm_predict_probability(scaled)
    with open(f_out, "rb") as fh:
        assert o.format_probability_output(lab, pr) == fh.read()
    assert ((pr.sum(1) - 1) < 1e-12).all() and (lab == np.where(pr[:, 1] > pr[:, 0], 1, -1)).all()


@pytest.mark.ref
@pytest.mark.parametrize("name", ["pcd2", "table1"])
def test_oracle_probability_grid_equals_reference_server(oracle_lib, prob_model, trained_model_path, clouds, name):
    """show_predicted_gps(roll, 0, true) of the REFERENCE'S OWN action server (compiled in place), fed by the reference's own
    svm-predict -b 1 child process, for every roll: per-roll tops and every published (positive) graspseval value equal the
    oracle's -- including the one-line shift (the header line is consumed as the first window's prediction)."""
    if not oracle_lib.refserver_available():
        pytest.skip("oracle/_ref/libhaf_refserver.so not built")
    srv = oracle_lib.RefServer(FEATURES, RANGE, trained_model_path)
    try:
        xyz = clouds[name]
        srv.run_goal(xyz, per_roll=False)
        top, pos, seen = srv.prob_rolls(xyz, prob_model)
    finally:
        srv.close()
    o = oracle_lib.Oracle(FEATURES, RANGE, prob_model)
    res = o.search_prob(xyz, oracle_lib.make_request())
    assert np.array_equal(res["per_roll_top"], top)
    n_pos = 0
    for roll in range(12):
        ev, mask = res["graspseval"][roll], res["mask"][roll]
        assert np.array_equal(seen[roll], mask)
        want = np.where((ev > 0) & (mask > 0), ev, 0).astype(np.float32)
        assert np.array_equal(pos[roll], want), roll
        n_pos += int((want > 0).sum())
        # the shift itself: the first valid window holds the header's value (0 * label0), the k-th one the (k-1)-th prediction
        grid = res["graspsgrid"][roll]
        cells = np.flatnonzero(mask.reshape(-1))
        if len(cells):
            assert grid.reshape(-1)[cells[0]] == 0
            pr = res["probs"][roll]
            lab = np.where(pr[:, 1] > pr[:, 0], 1, -1)
            val = np.where(lab > 0, pr[:, 1], pr[:, 0])
            val = np.array([float("%g" % v) for v in val]).astype(np.float32) * lab.astype(np.float32)
            assert np.array_equal(grid.reshape(-1)[cells[1:]], val[:-1])
    assert n_pos > 20
    # probability mode changes the answer: float scores, truncated tops
    assert (res["graspseval"] != np.round(res["graspseval"])).any()


# ------------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_predict_probability_matches_reference_output(hg, oracle_lib, prob_model, trained_model_path, clouds, tmp_path):
    o = oracle_lib.Oracle(FEATURES, RANGE, prob_model)
    # (1) the committed output of the reference program
    x = _dense_rows_of_file(os.path.join(CLI_DIR, "scaled_ref.txt"), 323)
    for mode in (hg.HAF_SVM_TENSOR_GUARD, hg.HAF_SVM_FP64_EXACT, hg.HAF_SVM_FP32_GUARD):   # estimates never depend on the handle's mode
        p = hg.SvmPredictor(prob_model, svm_mode=mode, min_dims=323)
        try:
            assert p.check_probability_model()
            lab, pr = p.predict_probability(x)
            with open(os.path.join(CLI_DIR, "prob_out_trained_ref.txt"), "rb") as fh:
                assert o.format_probability_output(lab, pr) == fh.read()
            lab0, _ = p.predict(x)      # the label-only call still works on the same handle (and agrees here)
            assert np.array_equal(lab0, lab)
        finally:
            p.close()
    # (2) three rolls of real windows against the oracle: same text; estimates to 1e-13 (exp: CUDA vs glibc)
    p = hg.SvmPredictor(prob_model, min_dims=323)
    try:
        for name, roll in (("pcd2", 0), ("table1", 5), ("plastic_mug2", 9)):
            scaled, _ = _roll_scaled(o, clouds[name], roll)
            lab, pr = p.predict_probability(scaled)
            olab, opr = o.svm_predict_probability(scaled)
            assert np.array_equal(lab, olab)
            assert np.abs(pr - opr).max() <= 1e-13
            assert o.format_probability_output(lab, pr) == o.format_probability_output(olab, opr)
        assert p.predict_probability(np.zeros((0, 323)))[1].shape == (0, 2)
    finally:
        p.close()
    # a model without probA / probB: libsvm's own refusal
    p = hg.SvmPredictor(trained_model_path, min_dims=323)
    try:
        assert not p.check_probability_model()
        with pytest.raises(hg.HafError) as e:
            p.predict_probability(x)
        assert e.value.code == -5 and "does not support probabiliy estimates" in str(e.value)
    finally:
        p.close()


@pytest.mark.gpu
def test_gpu_svm_predict_cli_b1_is_byte_identical(hg, oracle_lib, prob_model, trained_model_path, clouds, tmp_path):
    from haf_grasping_b200 import build
    if not os.path.exists(SVM_PREDICT_B200):
        build.build_svm_tools()
    scaled_file = os.path.join(CLI_DIR, "scaled_ref.txt")
    out = str(tmp_path / "out.txt")
    r = subprocess.run([SVM_PREDICT_B200, "-b", "1", scaled_file, prob_model, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with open(out, "rb") as a, open(os.path.join(CLI_DIR, "prob_out_trained_ref.txt"), "rb") as b:
        assert a.read() == b.read()
    with open(os.path.join(CLI_DIR, "prob_out_trained_ref.stdout")) as fh:
        assert r.stdout == fh.read()
    # the two messages of svm-predict.c:209-221
    r = subprocess.run([SVM_PREDICT_B200, "-b", "1", scaled_file, trained_model_path, out], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr == "Model does not support probabiliy estimates\n"
    r = subprocess.run([SVM_PREDICT_B200, scaled_file, prob_model, out], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("Model supports probability estimates, but disabled in prediction.\nAccuracy = ")
    # a whole roll against the reference program itself, where oracle/_ref is here
    ref_predict = os.path.join(oracle_lib.REF_DIR, "svm-predict")
    if os.path.exists(ref_predict):
        o = oracle_lib.Oracle(FEATURES, RANGE, prob_model)
        scaled, _ = _roll_scaled(o, clouds["table1"], 3)
        f_in = str(tmp_path / "roll.scale")
        _write_libsvm(f_in, scaled)
        a = subprocess.run([ref_predict, "-b", "1", f_in, prob_model, str(tmp_path / "ref.out")], capture_output=True, text=True, check=True)
        b = subprocess.run([SVM_PREDICT_B200, "-b", "1", f_in, prob_model, str(tmp_path / "our.out")], capture_output=True, text=True, check=True)
        assert open(str(tmp_path / "ref.out"), "rb").read() == open(str(tmp_path / "our.out"), "rb").read()
        assert a.stdout == b.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("pcd2", {}), ("table1", {}), ("table3", {"approach": (0.0, -0.5, 0.8660254)}),
                                     ("plastic_mug2", {"return_only_best": 1, "graspval_top": 60})])
def test_gpu_search_with_probability_equals_oracle(hg, oracle_lib, prob_model, clouds, name, kw):
    """haf_search with svm_with_probability = 1 (server.cpp:383-385 with the switch on): graspseval (float stencil), per-roll
    tops (truncation + tie rule), the best grasp and the early exit equal the oracle's restatement (pinned above against the
    reference server); the context's svm_mode does not matter (every window takes the FP64 exact-order path)."""
    o = oracle_lib.Oracle(FEATURES, RANGE, prob_model)
    xyz = clouds[name]
    okw = dict(kw)
    ores = o.search_prob(xyz, oracle_lib.make_request(**okw))
    for mode in (hg.HAF_SVM_TENSOR_GUARD, hg.HAF_SVM_FP64_EXACT):
        gpu = hg.GraspSearch(FEATURES, RANGE, prob_model, svm_mode=mode)
        try:
            res = gpu.search(xyz, [hg.make_request(svm_with_probability=1, **kw)])
            nr = len(ores["per_roll_top"])
            assert res["best"].astuple() == tuple(ores["best"]), (res["best"].astuple(), ores["best"])
            assert res["best"].rolls_done == nr
            assert np.array_equal(res["per_roll_top"][0][:nr], ores["per_roll_top"])
            for roll in range(nr):
                assert np.array_equal(res["mask"][0][roll], ores["mask"][roll])
                assert res["graspseval"][0][roll].tobytes() == ores["graspseval"][roll].tobytes(), roll
            # and the label-only answer of the same context is the usual one (different from the probability answer)
            res0 = gpu.search(xyz, [hg.make_request(**kw)])
            assert np.array_equal(res0["graspseval"][0], np.round(res0["graspseval"][0]))
        finally:
            gpu.close()


@pytest.mark.gpu
def test_gpu_search_with_probability_errors(hg, prob_model, trained_model_path, clouds):
    gpu = hg.GraspSearch(FEATURES, RANGE, trained_model_path)
    try:
        with pytest.raises(hg.HafError) as e:
            gpu.search(clouds["pcd2"], [hg.make_request(svm_with_probability=1)])
        assert e.value.code == -5 and "does not support probabiliy estimates" in str(e.value)
    finally:
        gpu.close()
    gpu = hg.GraspSearch(FEATURES, RANGE, prob_model)
    try:
        with pytest.raises(hg.HafError) as e:   # the switch is per goal in the reference: one value per call
            gpu.search(clouds["pcd2"], [hg.make_request(svm_with_probability=1), hg.make_request()])
        assert e.value.code == -1
        # a batch in probability mode equals the single searches
        xs = [clouds["pcd2"], clouds["pcd7"]]
        best = gpu.search_batch(xs, hg.make_request(svm_with_probability=1))
        for x, b in zip(xs, best):
            assert gpu.search(x, [hg.make_request(svm_with_probability=1)])["best"].astuple() == b.astuple()
    finally:
        gpu.close()
