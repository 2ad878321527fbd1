"""World-size-2 gloo tests of the sharding / collective logic.

The compute steps of pybot_b200.distributed are injected (`kernels=`,
`problem_factory=`); here they are the CPU oracle, so what is under test is the
HOST logic the NCCL path shares: shard ranges, global-row keys, the all-gather
+ unsigned-min merge, the owner-contributes all-reduce of the matched rows for
the mutual check and the cost all-reduce.  The result must equal the unsharded
oracle bit for bit.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model, port
from pybot_b200 import synthetic as syn
from pybot_b200.distributed import ShardedBA, ShardedKeyframeMatcher, shard_range

KEY_NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


class OracleMatchKernels(object):
    """numpy stand-ins with the exact contracts of the pbk_hamming_* entry points."""

    def knn2(self, q, t, train_base):
        idx, d = model.hamming_knn(q.numpy(), t.numpy(), 2)
        keys = (d.astype(np.uint64) << np.uint64(32)) | (idx.astype(np.int64) + int(train_base)).astype(np.uint64)
        keys[idx < 0] = KEY_NONE
        return torch.from_numpy(keys.view(np.int64))

    def merge(self, partial):
        p = partial.numpy().view(np.uint64)                     # (parts, nq, 2)
        allk = np.concatenate([p[i] for i in range(p.shape[0])], axis=1)
        allk.sort(axis=1)
        return torch.from_numpy(np.ascontiguousarray(allk[:, :2]).view(np.int64))

    def gather_best(self, t, keys, train_base):
        k = keys.numpy().view(np.uint64)[:, 0]
        rows = np.zeros((len(k), 32), np.uint8)
        g = (k & np.uint64(0xffffffff)).astype(np.int64) - int(train_base)
        own = (k != KEY_NONE) & (g >= 0) & (g < len(t))
        rows[own] = t.numpy()[g[own]]
        return torch.from_numpy(rows)

    def knnk(self, q, t, k, mask, train_base):
        idx, d = model.hamming_knn(np.asarray(q), np.asarray(t), k, mask=None if mask is None else np.asarray(mask))
        return (np.where(idx >= 0, idx.astype(np.int64) + int(train_base), -1), d, (idx >= 0).sum(1).astype(np.int32))

    def ratio_mutual(self, keys, rev, ratio):
        k = keys.numpy().view(np.uint64)
        has = k != KEY_NONE
        idx = np.where(has, (k & np.uint64(0xffffffff)).astype(np.int64), -1)
        d = np.where(has, (k >> np.uint64(32)).astype(np.int64), -1).astype(np.int32)
        r_ok = has[:, 1] & (d[:, 0].astype(np.float64) < ratio * d[:, 1].astype(np.float64))
        m_ok = np.zeros(len(k), bool)
        if rev is not None:
            rk = rev.numpy().view(np.uint64)[:, 0]
            m_ok = has[:, 0] & (rk != KEY_NONE) & ((rk & np.uint64(0xffffffff)).astype(np.int64) == np.arange(len(k)))
        return idx, d, r_ok, m_ok, r_ok & m_ok


class OracleBAProblem(object):
    def __init__(self, rvecs, tvecs, K, points, obs_cam, obs_pt, obs_uv):
        self.args = (rvecs, tvecs, K, points, obs_cam, obs_pt, obs_uv)
        self.cost = torch.zeros(1, dtype=torch.float64)
        self.n_obs = len(obs_cam)

    def evaluate(self):
        if self.n_obs:
            self.r, self.Jc, self.Jp, c = port.ba_residual_jacobian(*self.args, use_cv2=False)
            self.cost = torch.tensor([c], dtype=torch.float64)
        return self


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, fn, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), fn, out), nprocs=world, join=True)
        return [out[r] for r in range(world)]


def _match_job(rank, world):
    per_kf, nkf = 300, 7                                   # 7 keyframes: uneven shards (4 + 3)
    q, db = syn.c3_database(nq=160, n_keyframes=nkf, per_kf=per_kf, n_revisit=2)
    db[1900] = db[40]                                      # duplicate across the shard boundary
    q[5] = db[40]                                          # ... which query 5 hits exactly: tie on global row
    lo, hi = shard_range(nkf, rank, world)
    shard = torch.from_numpy(db[lo * per_kf:hi * per_kf].copy())
    m = ShardedKeyframeMatcher(shard, lo * per_kf, kernels=OracleMatchKernels())
    idx, d, r_ok, m_ok, valid = m.match(torch.from_numpy(q), 0.75)
    return [np.asarray(x) for x in (idx, d, r_ok, m_ok, valid)]


def test_sharded_keyframe_matching_equals_unsharded_oracle():
    per_kf, nkf = 300, 7
    q, db = syn.c3_database(nq=160, n_keyframes=nkf, per_kf=per_kf, n_revisit=2)
    db[1900] = db[40]
    q[5] = db[40]
    want = model.ratio_mutual(q, db, 0.75)
    got = _run(_match_job)
    for rank_out in got:                                   # identical on every rank
        for a, b in zip(rank_out, want):
            assert np.array_equal(np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64))
    assert want[0][5, 0] == 40 and want[0][5, 1] == 1900 and want[1][5, 1] == 0   # lowest global row first


def _knnk_case():
    per_kf, nkf = 120, 5                                   # 5 keyframes: uneven shards (3 + 2)
    q, db = syn.low_entropy_pair(seed=31, nq=60, nt=per_kf * nkf)     # tie-heavy: the merge order matters
    db[400] = db[7]                                        # duplicates across the shard boundary
    q[3] = db[7]
    mask = (np.random.default_rng(8).random((len(q), len(db))) < 0.05).astype(np.uint8)
    mask[11] = 0                                           # a query with no allowed row anywhere
    mask[12] = 0
    mask[12, 599] = 1                                      # ... and one whose only row is on the last rank
    return per_kf, nkf, q, db, mask


def _knnk_job(rank, world):
    per_kf, nkf, q, db, mask = _knnk_case()
    lo, hi = shard_range(nkf, rank, world)
    m = ShardedKeyframeMatcher(torch.from_numpy(db[lo * per_kf:hi * per_kf].copy()), lo * per_kf,
                               kernels=OracleMatchKernels())
    out = []
    for k, msk in ((1, None), (5, None), (9, mask), (700, None)):
        out.append([np.asarray(x) for x in m.knnk(torch.from_numpy(q), k, msk)])
    return out


def test_sharded_general_k_and_mask_equals_unsharded_oracle():
    """BFMatcher(group=...).knnMatch(k > 2 / mask): per-rank k best + all-gather + (distance, row)
    merge equals the unsharded model, ties across the shard boundary included."""
    per_kf, nkf, q, db, mask = _knnk_case()
    got = _run(_knnk_job)
    for i, (k, msk) in enumerate(((1, None), (5, None), (9, mask), (700, None))):
        idx, d = model.hamming_knn(q, db, k, mask=msk)
        for rank_out in got:
            assert np.array_equal(rank_out[i][0], idx) and np.array_equal(rank_out[i][1], d)
            assert np.array_equal(rank_out[i][2], (idx >= 0).sum(1))
    idx5 = model.hamming_knn(q, db, 5)[0]
    assert idx5[3, 0] == 7 and idx5[3, 1] == 400           # the duplicate pair, lowest global row first


def _ba_job(rank, world):
    prob = syn.c4_problem(n_cams=6, n_points=301)
    sb = ShardedBA(*prob, problem_factory=OracleBAProblem)
    cost = float(sb.evaluate().item())
    return (sb.range, cost, float(sb.problem.cost.item()))


def test_sharded_ba_cost_all_reduce():
    prob = syn.c4_problem(n_cams=6, n_points=301)
    n_obs = len(prob[4])
    _, _, _, want = port.ba_residual_jacobian(*prob, use_cv2=False)
    got = _run(_ba_job)
    ranges = [g[0] for g in got]
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == n_obs
    assert abs(got[0][1] - got[1][1]) == 0.0               # same global cost on both ranks
    assert abs(got[0][1] - want) <= 1e-9 * want
    assert abs(got[0][2] + got[1][2] - want) <= 1e-9 * want


@pytest.mark.parametrize("n,world", [(10, 3), (7, 8), (0, 2), (10000, 8)])
def test_shard_range_partitions(n, world):
    cuts = [shard_range(n, r, world) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == n
    for a, b in zip(cuts[:-1], cuts[1:]):
        assert a[1] == b[0]
    sizes = [hi - lo for lo, hi in cuts]
    assert max(sizes) - min(sizes) <= 1
