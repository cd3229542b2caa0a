"""Pins the oracle's restatement of the ROS-bound members against the REFERENCE'S OWN action server code.

oracle/_ref/libhaf_refserver.so is /root/reference/src/calc_grasppoints_action_server.cpp compiled in place and unmodified
(oracle/server_shim.cpp) against stand-in ROS / PCL / Eigen / OpenCV headers (oracle/stub_server/haf_ref_stubs.hpp lists the
third-party arithmetic restated there).  A goal runs read_pc_cb -> loop_control -> generate_grid / calc_intimage /
calc_featurevectors (pnt_in_box) / predict_bestgp_withsvm (the real svm-scale and svm-predict child processes on
/tmp/features.txt) / show_predicted_gps -> transform_gp_in_wcs_and_publish.  The oracle (oracle/haf_oracle.cpp) must
reproduce, bit for bit: the per-roll transforms (a1), height grids (a2-a3), integral images (a4), window masks (a5),
per-roll tops incl. the tie rule (a12-a14), graspseval where the server publishes it, the overall best grasp with the
loop rules (a15) and the GraspOutput numbers (a16).
"""
import gzip
import os

import numpy as np
import pytest

from conftest import FEATURES, GOLDEN, RANGE

pytestmark = pytest.mark.ref


@pytest.fixture(scope="module")
def clouds():
    return np.load(os.path.join(GOLDEN, "clouds.npz"))


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    p = str(tmp_path_factory.mktemp("trained_ref") / "substitute_trained.model")
    with gzip.open(os.path.join(GOLDEN, "substitute_trained.model.gz"), "rb") as src, open(p, "wb") as dst:
        dst.write(src.read())
    return p


@pytest.fixture(scope="module")
def pair(oracle_lib, model):
    if not oracle_lib.refserver_available():
        pytest.skip("oracle/_ref/libhaf_refserver.so not built")
    srv = oracle_lib.RefServer(FEATURES, RANGE, model)
    orc = oracle_lib.Oracle(FEATURES, RANGE, model)
    yield srv, orc
    srv.close()


CASES = [("pcd2", {}),
         ("table1", {}),
         ("table2", {"approach": (0.5, 0.0, 0.8660254), "center": (0.05, 0.2, 0.0)}),
         ("pcd7", {"width": 2, "area": (30.0, 40.0)}),
         ("table3", {"approach": (0.0, -0.5, 0.8660254)}),
         ("plastic_mug2", {"only_best": 1})]


@pytest.mark.parametrize("name,kw", CASES)
def test_oracle_equals_reference_server(pair, oracle_lib, clouds, name, kw):
    srv, o = pair
    xyz = clouds[name]
    ref = srv.run_goal(xyz, **kw)
    okw = dict(kw)
    if "only_best" in okw:
        okw["return_only_best"] = okw.pop("only_best")
    rq = oracle_lib.make_request(**okw)
    ores = o.search(xyz, rq)
    ob = ores["best"]
    n = ob.rolls_done
    # a1: the transform of every roll, bitwise
    avn = o.normalize_approach(tuple(rq.approach))
    for roll in range(srv.R):
        M_ref = srv.transform_of_roll(roll)
        M_orc = o.build_transform(tuple(rq.center), avn, rq.gripper_opening_width, roll)
        assert M_ref.tobytes() == M_orc.tobytes(), roll
    # a2-a5 on the rolls the loop evaluated
    assert ref["heights"][:n].tobytes() == ores["heights"][:n].tobytes()
    assert ref["integral"][:n].tobytes() == ores["integral"][:n].tobytes()
    assert np.array_equal(ref["mask"][:n], ores["mask"][:n])
    # a12-a14: per-roll tops (the shim's second pass evaluates all 12 rolls with the goal's members)
    full = o.search(xyz, oracle_lib.make_request(**{k: v for k, v in okw.items() if k != "return_only_best"}))
    assert np.array_equal(ref["per_roll_top"], full["per_roll_top"]), (ref["per_roll_top"], full["per_roll_top"])
    # graspseval where the server publishes it (markers of the mask-true cells; positive values are recoverable exactly)
    assert np.array_equal(ref["eval_seen"], full["mask"])
    pos = full["graspseval"] > 0
    assert np.array_equal(ref["eval_pos"][pos], full["graspseval"][pos])
    assert (ref["eval_pos"][~pos] == 0).all()
    # a15: overall best of the goal, with the loop rules (strict >, early exit when only_best)
    assert tuple(ref["best"]) == (ob.row, ob.col, ob.roll, ob.tilt, ob.topval)
    # a16: GraspOutput
    pose = o.transform_gp_in_wcs(rq, ores["heights"][max(ob.roll, 0)], ob.row, ob.col, ob.roll)
    g = ref["grasp"]
    assert int(g[0]) == ob.eval
    assert np.array_equal(g[1:13].astype(np.float32), pose[:12].astype(np.float32)), (g[1:13], pose[:12])
    assert np.float32(g[13]) == np.float32(pose[12]) == np.float32(ob.roll_rad)


def test_reference_server_empty_cloud_and_outside_points(pair, oracle_lib):
    srv, o = pair
    far = np.full((50, 3), 5.0, np.float32)   # every point outside the 56 cm box
    ref = srv.run_goal(far)
    ores = o.search(far, oracle_lib.make_request())
    assert tuple(ref["best"]) == ores["best"].astuple()
    assert np.array_equal(ref["mask"], ores["mask"]) and ref["mask"].sum() == 0
    assert ref["heights"].tobytes() == ores["heights"].tobytes()
