"""Several GPUs behind ONE C-ABI context (haf_config.n_devices / devices): the path a C++ host takes.
haf_search shards the (approach vector, roll) units of a goal over the GPUs and merges the per-unit tops with the loop's own
rules (server.cpp:362-365 early exit, :953-960 strict >); haf_search_batch* shards the clouds.  Everything must equal the
one-GPU answer exactly.  Needs >= 2 GPUs (skipped otherwise: the driver's single-GPU test run; run with `gpurun --gpus 2`)."""
import numpy as np
import pytest

from conftest import FEATURES, RANGE

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def pair(hg, tmp_models):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    devs = list(range(min(n, 4)))
    model = tmp_models(256)
    one = hg.GraspSearch(FEATURES, RANGE, model)
    grp = hg.GraspSearch(FEATURES, RANGE, model, devices=devs)
    yield one, grp, len(devs)
    one.close()
    grp.close()


TILTED = [(0.0, 0.0, 1.0), (0.5, 0.0, 0.8660254), (-0.5, 0.0, 0.8660254), (0.0, 0.5, 0.8660254), (0.0, -0.5, 0.8660254)]


def test_one_goal_sharded_by_unit_equals_one_gpu(pair, hg, clouds_npz):
    one, grp, n = pair
    for name, reqs in (("table1", [hg.make_request(approach=a) for a in TILTED]),                      # configs[2]: 60 units
                       ("pcd2", [hg.make_request()]),                                                   # 12 units
                       ("plastic_mug2", [hg.make_request(return_only_best=1, graspval_top=60)]),        # early exit replayed on the merge
                       ("table3", [hg.make_request(roll_begin=3, roll_limit=10), hg.make_request(approach=TILTED[1], roll_limit=5)])):
        xyz = clouds_npz[name]
        a, b = one.search(xyz, reqs), grp.search(xyz, reqs)
        assert a["best"].astuple() == b["best"].astuple(), name
        assert (a["best"].approach_idx, a["best"].rolls_done, a["best"].n_windows_scored) == (b["best"].approach_idx, b["best"].rolls_done, b["best"].n_windows_scored), name
        assert np.array_equal(a["per_roll_top"], b["per_roll_top"]), name
        for k in ("graspseval", "mask", "heights"):
            ra = [r for r in range(one.R)]
            # rolls the one-GPU loop evaluated: the sharded run evaluates at least those (it replays the early exit afterwards)
            done = a["best_per_request"][0].rolls_done if len(reqs) == 1 and reqs[0].return_only_best else None
            if done is None:
                assert np.array_equal(a[k], b[k]), (name, k)
            else:
                assert np.array_equal(a[k][0][:done], b[k][0][:done]), (name, k)
    # a cloud that lives on GPU 0: the other members fetch their own copy
    import torch
    xyz = clouds_npz["table2"]
    dev = torch.from_numpy(xyz).cuda(0)
    assert grp.search(dev, [hg.make_request()], n_points=len(xyz))["best"].astuple() == one.search(xyz, [hg.make_request()])["best"].astuple()


def test_batch_sharded_by_cloud_equals_one_gpu(pair, hg):
    from haf_grasping_b200 import synth
    one, grp, n = pair
    clouds = [synth.synth_cloud(500 + i, 20000 + 1000 * (i % 7)) for i in range(37)]     # ragged, not a multiple of the GPU count
    off = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int64)
    allpts = np.concatenate(clouds)
    a = one.search_batch_packed(allpts, off)
    b = grp.search_batch_packed(allpts, off)
    assert [x.astuple() for x in a] == [x.astuple() for x in b]
    assert [x.n_windows_scored for x in a] == [x.n_windows_scored for x in b]
    c = grp.search_batch(clouds)
    assert [x.astuple() for x in a] == [x.astuple() for x in c]
    t = grp.timing()
    assert t.n_windows == sum(x.n_windows_scored for x in a) and t.n_units == 37 * one.R
    # fewer clouds than GPUs
    d = grp.search_batch(clouds[:1])
    assert d[0].astuple() == a[0].astuple()
