"""GPU parity: the rest of cv2.BFMatcher.knnMatch's signature - any k, masks,
compactResult, the collection overload - through pbk_hamming_knnk, against golden g12 (generated
by cv2.BFMatcher, oracle/make_golden.py) and the numpy model on seeded inputs."""
import numpy as np
import pytest

from oracle import model
from pybot_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g12(golden_dir):
    return np.load(golden_dir + "/g12_knnk.npz")


def _rows(res, k, field):
    a = np.full((len(res), k), -1, np.int32)
    for i, r in enumerate(res):
        a[i, :len(r)] = [int(getattr(m, field)) for m in r]
    return a


def test_general_k_golden(g12):
    from pybot_b200.vision.matcher import BFMatcher
    q, t = g12["q"], g12["t"]
    for k in (3, 7, 40):
        res = BFMatcher().knnMatch(q, t, k=k)
        assert len(res) == len(q) and all(len(r) == k for r in res)
        assert np.array_equal(_rows(res, k, "trainIdx"), g12["k%d_idx" % k])
        assert np.array_equal(_rows(res, k, "distance"), g12["k%d_dist" % k])
        assert all(m.queryIdx == i and m.imgIdx == 0 for i, r in enumerate(res) for m in r)


def test_mask_golden_and_compact(g12):
    from pybot_b200.vision.matcher import BFMatcher
    q, t, mask = g12["q"], g12["t"], g12["mask"]
    for k in (2, 5):
        res = BFMatcher().knnMatch(q, t, k=k, mask=mask)
        assert np.array_equal(np.array([len(r) for r in res]), g12["m%d_len" % k])
        assert np.array_equal(_rows(res, k, "trainIdx"), g12["m%d_idx" % k])
        assert np.array_equal(_rows(res, k, "distance"), g12["m%d_dist" % k])
    assert res[3] == [] and len(res[5]) == 1 and res[5][0].trainIdx == 600
    res = BFMatcher().knnMatch(q, t, k=5, mask=mask.astype(bool), compactResult=True)
    assert np.array_equal(np.array([r[0].queryIdx for r in res]), g12["m5_compact_query"])
    one = BFMatcher().match(q, t, mask=mask)                   # k = 1 with a mask, empty rows dropped
    assert [m.trainIdx for m in one] == g12["m2_idx"][g12["m2_len"] > 0, 0].tolist()


def test_collection_overload_with_masks(g12):
    from pybot_b200.vision.matcher import BFMatcher
    q, t, mask, cuts = g12["q"], g12["t"], g12["mask"], g12["cuts"]
    bf = BFMatcher()
    bf.add([t[a:b] for a, b in zip(cuts[:-1], cuts[1:])])
    res = bf.knnMatch(q, 6)                                    # cv2's (query, k) overload
    assert np.array_equal(_rows(res, 6, "imgIdx"), g12["c6_img"])
    assert np.array_equal(_rows(res, 6, "trainIdx"), g12["c6_idx"])
    assert np.array_equal(_rows(res, 6, "distance"), g12["c6_dist"])
    res = bf.knnMatch(q, k=4, masks=[mask[:, a:b] for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.array_equal(_rows(res, 4, "imgIdx"), g12["c4_img"])
    assert np.array_equal(_rows(res, 4, "trainIdx"), g12["c4_idx"])
    assert np.array_equal(_rows(res, 4, "distance"), g12["c4_dist"])
    with_none = bf.knnMatch(q, k=4, masks=[None, mask[:, cuts[1]:cuts[2]], None])
    full = np.ones_like(mask)
    full[:, cuts[1]:cuts[2]] = mask[:, cuts[1]:cuts[2]]
    idx, dist = model.hamming_knn(q, t, 4, mask=full)
    assert np.array_equal(with_none.arrays["idx"], idx) and np.array_equal(with_none.arrays["dist"], dist)


@pytest.mark.parametrize("nq,nt,k,density", [
    (1, 1, 1, None), (3, 5, 8, None), (70, 255, 3, None), (70, 256, 256, None), (33, 257, 31, 0.5),
    (257, 1000, 1024, None), (64, 5000, 100, 0.02), (9, 70001, 17, None), (300, 4099, 5, 0.001),
    (1500, 300, 4, 0.3)])      # more queries than resident CTAs: the per-CTA query loop
def test_seeded_shapes_against_model(nq, nt, k, density):
    """tie-heavy descriptors around the 256-row chunk boundaries of the ordered tie selection,
    k above / below the row count, k at the supported maximum, sparse and dense masks."""
    from pybot_b200 import ops
    q, t = syn.low_entropy_pair(seed=nq * 7 + k, nq=nq, nt=nt)
    mask = None
    if density is not None:
        mask = (np.random.default_rng(nt).random((nq, nt)) < density).astype(np.uint8)
    idx, dist, count = ops.hamming_knnk(q, t, k, mask)
    ridx, rdist = model.hamming_knn(q, t, k, mask=mask)
    assert idx.dtype == np.int64 and dist.dtype == np.int32
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)
    assert np.array_equal(count, (ridx >= 0).sum(1))


def test_identical_database_takes_lowest_rows():
    """every row equal: all distances tie, the k lowest rows in order."""
    from pybot_b200 import ops
    q = np.full((5, 32), 0xA5, np.uint8)
    t = np.full((3000, 32), 0xA5, np.uint8)
    t[:, 0] = 0xA4                                             # distance 1 everywhere
    idx, dist, count = ops.hamming_knnk(q, t, 600, train_base=1000)
    assert np.array_equal(idx, np.tile(np.arange(1000, 1600), (5, 1))) and (dist == 1).all() and (count == 600).all()


def test_k2_equals_hot_path():
    """k = 2 without a mask through the general kernel equals the hot path's keys."""
    from pybot_b200 import ops
    from pybot_b200.vision.matcher import knn_ratio_mutual
    q, t = syn.orb_pair(seed=21, nq=500, nt=3000)
    idx, dist, _ = ops.hamming_knnk(q, t, 2)
    hidx, hdist = knn_ratio_mutual(q, t, 0.75)[:2]
    assert np.array_equal(idx, np.asarray(hidx)) and np.array_equal(dist, np.asarray(hdist))


def test_device_tensors_and_errors():
    import torch
    from pybot_b200 import ops
    from pybot_b200.vision.matcher import BFMatcher
    q, t = syn.orb_pair(seed=4, nq=20, nt=90)
    idx, dist, count = ops.hamming_knnk(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), 3)
    assert idx.is_cuda and np.array_equal(idx.cpu().numpy(), model.hamming_knn(q, t, 3)[0])
    empty = ops.hamming_knnk(q, t[:0], 3)
    assert (empty[0] == -1).all() and (empty[2] == 0).all()
    with pytest.raises(NotImplementedError):
        ops.hamming_knnk(q, t, ops.KNNK_MAX_K + 1)
    with pytest.raises(ValueError):
        ops.hamming_knnk(q, t, 0)
    with pytest.raises(ValueError):
        ops.hamming_knnk(q, t, 3, mask=np.ones((20, 91), np.uint8))
    with pytest.raises(TypeError):
        ops.hamming_knnk(q, t, 3, mask=np.ones((20, 90), np.float32))
    with pytest.raises(ValueError):
        BFMatcher(crossCheck=True).knnMatch(q, t, k=3)
    with pytest.raises(ValueError):
        BFMatcher().knnMatch(q, t, k=0)


def test_sharded_keys_merge_equals_whole_database():
    """the sharded form (BFMatcher(group=...).knnMatch with k > 2 / a mask): per-shard k best as
    (distance, global row) keys, merged - here three shards of one GPU, the all-gather replaced by
    a concatenation - equals the single call over the whole database, ties across shards included."""
    import torch
    from pybot_b200 import ops
    from pybot_b200.distributed import ShardedKeyframeMatcher
    q, t = syn.low_entropy_pair(seed=77, nq=150, nt=3000)
    t[2500] = t[9]
    q[4] = t[9]
    mask = (np.random.default_rng(3).random((150, 3000)) < 0.03).astype(np.uint8)
    mask[7] = 0
    qd = torch.from_numpy(q).cuda()
    cuts = [0, 1100, 1101, 3000]
    shards = [ShardedKeyframeMatcher(torch.from_numpy(t[a:b]).cuda(), a) for a, b in zip(cuts[:-1], cuts[1:])]
    for k, m in ((6, None), (6, mask), (64, torch.from_numpy(mask).cuda())):
        keys = torch.cat([s.knnk_local_keys(qd, k, m) for s in shards], dim=1)
        idx, dist, count = (x.cpu().numpy() for x in ShardedKeyframeMatcher.knnk_merge(keys, k))
        mm = None if m is None else mask
        ridx, rdist = model.hamming_knn(q, t, k, mask=mm)
        assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist) and np.array_equal(count, (ridx >= 0).sum(1))
        one = ShardedKeyframeMatcher(torch.from_numpy(t).cuda(), 0).knnk(qd, k, m)     # world 1: same call path
        assert np.array_equal(one[0].cpu().numpy(), ridx)
    assert model.hamming_knn(q, t, 2)[0][4].tolist() == [9, 2500]      # the duplicate pair sits in two shards
