"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: unit sharding, the record exchange and the merge rules.
The per-rank evaluation is done by the CPU oracle here (on a GPU box the same code path calls GraspSearch.search)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import FEATURES, GOLDEN, RANGE, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, model, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from haf_grasping_b200 import distributed as hd
    from oracle import orc
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    clouds = np.load(os.path.join(GOLDEN, "clouds.npz"))
    o = orc.Oracle(FEATURES, RANGE, model)
    avs = [(0.0, 0.0, 1.0), (0.5, 0.0, 0.8660254), (0.0, -0.5, 0.8660254)]
    xyz = clouds["pcd2"]
    R = 12

    def evaluate(a, rb, re):  # the oracle evaluates rolls [0, re) and we keep [rb, re): same tops as a GPU roll_begin call
        res = o.search(xyz, orc.make_request(approach=avs[a], roll_limit=re))
        return res["per_roll_top"][rb:re]

    for only_best, top in ((0, 119), (1, 60)):
        per, overall, tops = hd.sharded_goal_search(evaluate, len(avs), R, [only_best] * 3, [top] * 3, rank, world)
        np.save(os.path.join(out_dir, "tops_%d_%d.npy" % (only_best, rank)), tops)
        with open(os.path.join(out_dir, "res_%d_%d.txt" % (only_best, rank)), "w") as fh:
            fh.write(repr((per, overall)))
    # by-cloud sharding + record gather
    names = ["pcd1", "pcd3", "pcd4", "pcd6", "plastic_mug2"]
    b, e = hd.shard_range(len(names), rank, world)
    rec = np.array([o.search(clouds[n], orc.make_request(), full=False)["best"].astuple() for n in names[b:e]], np.int32).reshape(-1, 5)
    pad = np.full((3, 5), -7, np.int32)  # equal shapes for all_gather
    pad[:len(rec)] = rec
    allrec = hd.all_gather_records(pad)
    np.save(os.path.join(out_dir, "gather_%d.npy" % rank), allrec)
    dist.destroy_process_group()


def test_unit_blocks_partition_every_unit_once():
    from haf_grasping_b200 import distributed as hd
    for nreq, R, world in ((5, 12, 8), (1, 12, 8), (3, 12, 2), (2, 7, 3), (1, 3, 8)):
        seen = np.zeros(nreq * R, int)
        for rank in range(world):
            for a, rb, re in hd.unit_blocks(nreq, R, rank, world):
                assert 0 <= rb < re <= R
                seen[a * R + rb:a * R + re] += 1
        assert (seen == 1).all()
        sizes = [sum(re - rb for _, rb, re in hd.unit_blocks(nreq, R, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_merge_rules_match_the_reference_loop():
    from haf_grasping_b200 import distributed as hd
    tops = np.array([[1, 1, 50], [2, 2, 90], [3, 3, 90], [4, 4, 120], [5, 5, 121]] + [[0, 0, 0]] * 7, np.int32)
    per, overall = hd.merge_unit_tops(tops, 1, 12, [0], [119])
    assert per[0] == (5, 5, 4, 121, 12)             # all rolls, strict > keeps the first 90 until beaten
    per, overall = hd.merge_unit_tops(tops, 1, 12, [1], [119])
    assert per[0] == (4, 4, 3, 120, 4)              # early exit once topval >= 119: roll 4 is never evaluated
    two = np.concatenate([tops, tops])
    per, overall = hd.merge_unit_tops(two, 2, 12, [0, 0], [119, 119])
    assert overall[0] == 0                           # equal tops: the earlier request wins


def test_two_rank_gloo_sharded_goal_and_gather(oracle_lib, tmp_models, tmp_path):
    model = tmp_models(256)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, model, str(tmp_path)), nprocs=2, join=True)
    from oracle import orc
    o = orc.Oracle(FEATURES, RANGE, model)
    clouds = np.load(os.path.join(GOLDEN, "clouds.npz"))
    avs = [(0.0, 0.0, 1.0), (0.5, 0.0, 0.8660254), (0.0, -0.5, 0.8660254)]
    for only_best, top in ((0, 119), (1, 60)):
        r0 = open(tmp_path / ("res_%d_0.txt" % only_best)).read()
        r1 = open(tmp_path / ("res_%d_1.txt" % only_best)).read()
        assert r0 == r1                                  # both ranks hold the same merged answer
        per, overall = eval(r0)
        t0 = np.load(tmp_path / ("tops_%d_0.npy" % only_best))
        tops_best = []
        for a, av in enumerate(avs):
            ref = o.search(clouds["pcd2"], orc.make_request(approach=av, return_only_best=only_best, graspval_top=top))
            b = ref["best"]
            assert per[a] == (b.row, b.col, b.roll, b.topval, b.rolls_done)
            full = o.search(clouds["pcd2"], orc.make_request(approach=av))
            assert np.array_equal(t0[a * 12:(a + 1) * 12], full["per_roll_top"])
            tops_best.append(b.topval)
        assert overall[0] == int(np.argmax(tops_best)) and overall[4] == max(tops_best)
    g0, g1 = np.load(tmp_path / "gather_0.npy"), np.load(tmp_path / "gather_1.npy")
    assert np.array_equal(g0, g1) and g0.shape == (6, 5)
    names = ["pcd1", "pcd3", "pcd4", "pcd6", "plastic_mug2"]
    exp = [o.search(clouds[n], orc.make_request(), full=False)["best"].astuple() for n in names]
    got = [tuple(r) for r in g0 if r[0] != -7]
    assert got == exp
