"""Pins the oracle restatement against the reference's own compiled code (oracle/_ref).

Needs oracle/_ref (built in the build container from /root/reference; travels to the GPU box as
prebuilt files).  CPU only.
"""
import os

import numpy as np
import pytest

from conftest import FEATURES, GOLDEN, RANGE

pytestmark = pytest.mark.ref


@pytest.fixture(scope="module")
def pair(oracle_lib, tmp_models):
    if not oracle_lib.ref_available():
        pytest.skip("oracle/_ref not built")
    model = tmp_models(256)
    return oracle_lib.Oracle(FEATURES, RANGE, model), oracle_lib.Ref(FEATURES, RANGE, model)


@pytest.fixture(scope="module")
def clouds():
    return np.load(os.path.join(GOLDEN, "clouds.npz"))


def test_feature_table_matches_reference_parser(pair):
    o, r = pair
    assert o.F == r.F == 324  # 323 lines + the phantom feature from the trailing blank line
    reg_o, w_o = o.feature_table()
    reg_r, w_r = r.feature_table()
    assert (reg_o == reg_r).all()
    assert (w_o.astype(np.float64) == w_r).all()
    assert (w_r[:, 3] == 0).all()  # Haar.cpp:55-60: 4th weight never assigned
    assert (reg_r[323] == 0).all()


def test_feature_values_bitwise_vs_reference_class(pair, clouds):
    o, r = pair
    rng = np.random.default_rng(0)
    av = o.normalize_approach((0, 0, 1))
    for name in ("pcd2", "table1", "plastic_mug2"):
        integral = o.calc_intimage(o.generate_grid(clouds[name], o.build_transform((0, 0, 0), av, 1, 2)))
        for _ in range(40):
            row, col = rng.integers(0, 42, 2)
            patch = np.ascontiguousarray(integral[row:row + 15, col:col + 15])
            a, b = o.featurevalues(patch), r.featurevalues(patch)
            assert a.tobytes() == b.tobytes()
    # random patches incl. negative / huge / tiny magnitudes
    for scale in (1.0, 1e-6, 1e6):
        patch = np.cumsum(np.cumsum(rng.normal(size=(15, 15)) * scale, 0), 1).astype(np.float32)
        assert o.featurevalues(patch).tobytes() == r.featurevalues(patch).tobytes()


@pytest.mark.parametrize("name,roll", [("pcd2", 0), ("pcd2", 5), ("table1", 3), ("pcd4", 0), ("plastic_mug2", 11)])
def test_roll_file_exact(pair, clouds, name, roll):
    """features.txt -> svm-scale -> svm-predict through real child processes == oracle, label for label
    and scaled-text for scaled-text."""
    o, r = pair
    av = o.normalize_approach((0, 0, 1))
    integral = o.calc_intimage(o.generate_grid(clouds[name], o.build_transform((0, 0, 0), av, 1, roll)))
    mask = o.pnt_in_box(integral, roll)
    labels_ref, scaled_lines = r.roll_file_exact(integral, mask)
    feats, _ = o.calc_featurevectors(integral, mask)
    scaled = o.scale(feats)
    dec, labels = o.svm_decision(scaled)
    assert len(labels_ref) == len(labels) == int(mask.sum())
    assert (labels_ref == labels).all()
    # scaled text: the oracle's doubles printed with %g must reproduce svm-scale's file exactly
    for w, line in enumerate(scaled_lines):
        toks = line.split()
        mine = ["%d:%s" % (k + 1, "%g" % v) for k, v in enumerate(scaled[w]) if v != 0]
        assert toks[1:] == mine, (w, toks[:5], mine[:4])
    # in-process reference libsvm on the oracle's vectors: decision values bit-identical
    for w in range(0, len(labels), max(1, len(labels) // 25)):
        d_ref, l_ref = r.svm_predict(scaled[w])
        assert d_ref == dec[w] and l_ref == labels[w]


def test_label_orders(oracle_lib, tmp_models, clouds):
    """label -1 1 order flips the sign convention (svm.cpp:2516-2531)."""
    if not oracle_lib.ref_available():
        pytest.skip("oracle/_ref not built")
    m = tmp_models(256, labels=(-1, 1), rho=0.01)
    o, r = oracle_lib.Oracle(FEATURES, RANGE, m), oracle_lib.Ref(FEATURES, RANGE, m)
    av = o.normalize_approach((0, 0, 1))
    integral = o.calc_intimage(o.generate_grid(clouds["pcd3"], o.build_transform((0, 0, 0), av, 1, 7)))
    mask = o.pnt_in_box(integral, 7)
    labels_ref, _ = r.roll_file_exact(integral, mask)
    feats, _ = o.calc_featurevectors(integral, mask)
    dec, labels = o.svm_decision(o.scale(feats))
    assert (labels_ref == labels).all()
    assert ((dec > 0) == (labels == -1)).all()
