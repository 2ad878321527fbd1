"""GPU parity: Hamming kNN k=2 + ratio + mutual with BFMatcher semantics (a9)."""
import os

import numpy as np
import pytest

from oracle import model, port
from pybot_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["popc", "tc", "tc8"])
def engine(request, monkeypatch):
    """Every test of this file runs on all three scan engines: the integer-pipe kernel, the
    tcgen05 kind::mxf4 kernel ("tc": the default expanded format, 4) and the tcgen05 kind::i8
    kernel ("tc8": format 8) - the tensor-core ones forced here even for tiny databases and the
    reverse pass."""
    from pybot_b200 import ops
    monkeypatch.setenv("PBK_HAMMING_ENGINE", request.param[:2] if request.param != "popc" else "popc")
    before = ops.hamming_tc_format()
    ops.hamming_tc_set_format(8 if request.param == "tc8" else 4)
    yield request.param
    ops.hamming_tc_set_format(before)


@pytest.fixture(scope="module")
def g3(golden_dir):
    return np.load(golden_dir + "/g3_matching.npz")


def _check_against_model(q, t, ratio=0.75):
    from pybot_b200.vision.matcher import knn_ratio_mutual
    idx, dist, r_ok, m_ok, valid = knn_ratio_mutual(q, t, ratio)
    ridx, rdist, rr, rm, rv = model.ratio_mutual(q, t, ratio)
    assert np.array_equal(idx, ridx.astype(np.int64))
    assert np.array_equal(dist, rdist)
    assert np.array_equal(r_ok, rr) and np.array_equal(m_ok, rm) and np.array_equal(valid, rv)
    return idx, dist, valid


def test_low_entropy_ties_golden(g3):
    """46%-tied descriptors: lowest-train-index tie-break, exact duplicates."""
    from pybot_b200.vision.matcher import BFMatcher
    q, t = g3["le_q"], g3["le_t"]
    idx, dist, valid = _check_against_model(q, t)
    assert np.array_equal(idx, g3["le_idx"]) and np.array_equal(dist, g3["le_dist"])
    assert (dist[:, 0] == dist[:, 1]).mean() > 0.3
    res = BFMatcher().knnMatch(q, t, k=2)
    assert [[m.trainIdx for m in r] for r in res] == g3["le_idx"].tolist()
    assert [[int(m.distance) for m in r] for r in res] == g3["le_dist"].tolist()
    assert all(m.imgIdx == 0 and m.queryIdx == i for i, r in enumerate(res) for m in r)
    cross = BFMatcher(crossCheck=True).match(q, t)
    assert [[m.queryIdx, m.trainIdx] for m in cross] == g3["le_cross"].tolist()


def test_train_smaller_than_k(g3):
    from pybot_b200.vision.matcher import BFMatcher
    res = BFMatcher().knnMatch(g3["le_q"][:4], g3["le_t"][:1], k=2)
    assert [len(r) for r in res] == g3["nt1_len"].tolist()
    _check_against_model(g3["le_q"][:4], g3["le_t"][:1])
    empty = BFMatcher().knnMatch(g3["le_q"][:4], np.zeros((0, 32), np.uint8), k=2)
    assert empty == [[], [], [], []]


def test_keyframe_collection_golden(g3):
    """BFMatcher.add([kf0, kf1, kf2]) ordering incl. a duplicate across keyframes."""
    from pybot_b200.vision.matcher import BFMatcher
    bf = BFMatcher()
    bf.add([g3["kf_0"], g3["kf_1"], g3["kf_2"]])
    res = bf.knnMatch(g3["kf_q"], k=2)
    assert [[m.imgIdx for m in r] for r in res] == g3["kf_img"].tolist()
    assert [[m.trainIdx for m in r] for r in res] == g3["kf_idx"].tolist()
    assert [[int(m.distance) for m in r] for r in res] == g3["kf_dist"].tolist()
    assert len(bf.getTrainDescriptors()) == 3
    bf.clear()
    assert bf.empty()


def test_c1_orb_2000x2000():
    q, t = syn.orb_pair()
    idx, dist, valid = _check_against_model(q, t)
    assert 0.3 < valid.mean() < 0.7          # planted matches survive ratio + mutual
    pidx, pdist, pr, pm, pv = port.knn_ratio_mutual(q, [t])       # cv2 engine when importable
    assert np.array_equal(idx, pidx) and np.array_equal(dist, pdist) and np.array_equal(valid, pv)


@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 2), (5, 255), (511, 256), (513, 257), (1000, 1025),
                                   (37, 70001)])
def test_ragged_sizes(nq, nt):
    rng = np.random.default_rng(nq * 7919 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    t[rng.integers(0, nt)] = q[0]
    _check_against_model(q, t)


@pytest.mark.parametrize("nq,nt", [(520, 150 * 256 + 200), (700, 160 * 256 + 128), (513, 149 * 256 + 129),
                                   (2000, 148 * 256 + 1)])
def test_ragged_last_tile_in_the_throughput_shape(nq, nt):
    """Sizes that take the 4-queries-per-thread shape (packed 16-bit top-2 over 128-row blocks)
    and end in a partial tile whose second block is ragged / empty / one row; ties planted across
    the block and tile boundaries of that last tile.  Checked against the C restatement."""
    from oracle import c_oracle
    from pybot_b200.vision.matcher import knn_ratio_mutual
    q, t = syn.low_entropy_pair(seed=nq + nt, nq=nq, nt=nt)
    last = (nt // 256) * 256
    for k, row in enumerate((last - 1, last, min(last + 127, nt - 1), min(last + 128, nt - 1), nt - 1)):
        t[row] = q[k % 3]
    idx, dist = knn_ratio_mutual(q, t)[:2]
    ridx, rdist = c_oracle.hamming_knn2(q, t)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)


@pytest.mark.parametrize("variant", ["0", "1", "2"])
def test_both_popc_variants_agree(variant):
    """PBK_HAMMING_VARIANT selects 8-POPC / carry-save 4-POPC / carry-save with the packed
    u16x2 top-2 network: same keys.  The second problem is large enough for the
    4-queries-per-thread shape (the only one variant 2 applies to) and is full of ties."""
    import subprocess, sys
    code = (
        "import numpy as np, sys; sys.path.insert(0, %r);"
        "from pybot_b200 import synthetic as syn; from pybot_b200.vision.matcher import knn_ratio_mutual;"
        "from oracle import model;"
        "q,t = syn.low_entropy_pair(seed=3, nq=700, nt=3000);"
        "a = knn_ratio_mutual(q,t); b = model.ratio_mutual(q,t);"
        "assert all(np.array_equal(np.asarray(x).astype(np.int64), np.asarray(y).astype(np.int64)) for x,y in zip(a,b));"
        "from oracle import c_oracle;"
        "q,t = syn.low_entropy_pair(seed=4, nq=700, nt=50021);"
        "t[130] = ~q[5]; t[40000] = q[5]; t[40001] = q[5]; t[127] = q[9]; t[128] = q[9];"
        "idx, dist = knn_ratio_mutual(q,t)[:2]; ridx, rdist = c_oracle.hamming_knn2(q,t);"
        "assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist);"
        "print('ok')") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PBK_HAMMING_VARIANT=variant)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_database_sized_vs_c_oracle():
    """2000 queries x 400k rows (200 keyframes): GPU vs the C restatement of
    BFMatcher (oracle/c), global indices, with planted revisits."""
    from oracle import c_oracle
    from pybot_b200.vision.matcher import knn_ratio_mutual
    q, db = syn.c3_database(nq=2000, n_keyframes=200, per_kf=2000, n_revisit=3)
    idx, dist, r_ok, m_ok, valid = knn_ratio_mutual(q, db, 0.75)
    ridx, rdist = c_oracle.hamming_knn2(q, db)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)
    assert valid.sum() > 300          # planted in 3 keyframes: many top-2 pairs are both true matches


def test_sharded_merge_equals_single_scan():
    """Split the DB in 3 uneven shards, merge the partial keys: identical to one scan
    (the multi-GPU path run on one device)."""
    import torch
    from pybot_b200 import ops
    q, db = syn.c3_database(nq=600, n_keyframes=30, per_kf=1000, n_revisit=2)
    db[12345] = db[777]                           # cross-shard duplicate -> tie on global index
    qd, dbd = torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda()
    whole = ops.hamming_knn2_keys(qd, dbd, 0)
    cuts = [0, 7000, 19000, 30000]
    parts = torch.stack([ops.hamming_knn2_keys(qd, dbd[a:b], a) for a, b in zip(cuts[:-1], cuts[1:])])
    merged = ops.hamming_merge(parts)
    assert torch.equal(whole, merged)


def test_full_size_properties_c3_shard():
    """One 8-GPU shard of C3 (2.5M rows): self-match property - planting each
    query in the shard must return distance 0 at the planted (lowest) row."""
    import torch
    from pybot_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1003)
    nt, nq = 2_500_000, 2000
    db = torch.randint(0, 256, (nt, 32), device="cuda", dtype=torch.uint8, generator=g)
    q = torch.randint(0, 256, (nq, 32), device="cuda", dtype=torch.uint8, generator=g)
    rows = torch.randperm(nt, device="cuda", generator=g)[:nq]
    db[rows] = q
    keys = ops.hamming_knn2_keys(q, db, 5_000_000)
    assert torch.equal(keys[:, 0] >> 32, torch.zeros(nq, dtype=torch.int64, device="cuda"))
    assert torch.equal(keys[:, 0] & 0xffffffff, rows + 5_000_000)
    d2 = keys[:, 1] >> 32
    assert int(d2.min()) > 60 and int(d2.max()) < 110          # random 256-bit: ~ 85-95


def test_popc_probe_runs():
    from pybot_b200 import ops
    assert ops.popc_probe(mode=0, ctas=148, iters=10) == 16 * 8 * 10 * 256 * 148


def test_prepared_matcher_cuda_graph_equals_eager():
    """PreparedMatcher (buffers allocated once, launches captured in a CUDA graph) on several
    frames: identical to the eager pipeline and to the oracle."""
    import torch
    from oracle import model
    from pybot_b200 import ops
    pm = ops.PreparedMatcher(2000, 2000, 0.75)
    assert pm.graph is not None
    for seed in (1, 2, 3):
        q, t = syn.orb_pair(seed=seed)
        got = [x.cpu().numpy() for x in pm.run(q, t)]
        want = model.ratio_mutual(q, t, 0.75)
        for a, b in zip(got, want):
            assert np.array_equal(np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64))
        eager = ops.knn_ratio_mutual(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), 0.75)
        for a, b in zip(pm.run(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()), eager):
            assert torch.equal(a, b)
    with pytest.raises(ValueError):
        pm.run(q[:10], t)


def test_all_rows_tied_and_extreme_distances():
    """Every train row identical (every pair tied), across stage tiles, train splits
    and a ragged last tile: the two LOWEST rows win for every query, at distance 0,
    d and 256 (the complement) - the packed-key range end to end."""
    from pybot_b200.vision.matcher import knn_ratio_mutual
    rng = np.random.default_rng(11)
    nt = 300_001
    row = rng.integers(0, 256, 32, dtype=np.uint8)
    t = np.tile(row, (nt, 1))
    q = rng.integers(0, 256, (700, 32), dtype=np.uint8)
    q[0] = row                                   # distance 0
    q[1] = ~row                                  # distance 256
    idx, dist, r_ok, m_ok, valid = knn_ratio_mutual(q, t, 0.75)
    d = np.unpackbits(q ^ row, axis=1).sum(1)
    assert np.array_equal(idx, np.tile(np.int64([0, 1]), (700, 1)))
    assert np.array_equal(dist, np.stack([d, d], 1)) and dist[0, 0] == 0 and dist[1, 0] == 256
    assert not r_ok[1:].any()                    # d1 == d2 never passes the ratio test (0 < 0.75*0 is false too)


def test_duplicates_on_tile_and_split_boundaries():
    """Exact copies of a query planted on both sides of every 256-row stage boundary
    in a window, and at the very end: the top-2 are the two lowest planted rows."""
    from oracle import c_oracle
    from pybot_b200.vision.matcher import knn_ratio_mutual
    rng = np.random.default_rng(12)
    nt = 1_000_003
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (513, 32), dtype=np.uint8)     # one more than a query tile
    planted = {}
    for i, base in enumerate(range(256 * 7, nt - 600, 61_184)):        # 61184 = 239 * 256
        rows = [base - 1, base, base + 255, base + 256, nt - 1 - i]
        t[rows] = q[i]
        planted[i] = sorted(rows)
    idx, dist, _, _, _ = knn_ratio_mutual(q, t, 0.75)
    for i, rows in planted.items():
        assert list(idx[i]) == rows[:2] and list(dist[i]) == [0, 0], (i, idx[i], rows)
    ridx, rdist = c_oracle.hamming_knn2(q[:64], t)
    assert np.array_equal(idx[:64], ridx) and np.array_equal(dist[:64], rdist)


def test_keyframe_database_grows_in_place():
    """Keyframes appended one at a time (across several capacity doublings) give the
    same global rows as one flat matrix; appends within capacity do not move the data."""
    from pybot_b200.vision.matcher import BFMatcher, KeyframeDatabase
    rng = np.random.default_rng(21)
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    kfs = [rng.integers(0, 256, (int(n), 32), dtype=np.uint8) for n in rng.integers(1, 9000, 40)]
    kfs[7][3] = q[5]
    kfs[31][0] = q[5]                                      # same descriptor in two keyframes: lower image wins
    m = BFMatcher()
    ptrs = set()
    for kf in kfs:
        m.add([kf])
        ptrs.add(m.db.flat.data_ptr())
    assert len(m.db) == sum(len(k) for k in kfs) and len(ptrs) <= 6      # geometric growth, not one copy per add
    got = m.knnMatch(q, k=2)
    flat = np.concatenate(kfs)
    idx, dist, _, _, _ = __import__("pybot_b200.vision.matcher", fromlist=["x"]).knn_ratio_mutual(q, flat, 0.75)
    off = np.cumsum([0] + [len(k) for k in kfs])
    for i in (0, 5, 17, 299):
        g = [off[d.imgIdx] + d.trainIdx for d in got[i]]
        assert g == list(idx[i]) and [int(d.distance) for d in got[i]] == list(dist[i])
    assert (got[5][0].imgIdx, got[5][0].trainIdx, got[5][1].imgIdx, got[5][1].trainIdx) == (7, 3, 31, 0)
    db = KeyframeDatabase()
    db.reserve(1_000_000)
    p0 = db.flat.data_ptr() if len(db) else None
    db.add(kfs[:3])
    p1 = db.flat.data_ptr()
    db.add(kfs[3:])
    assert db.flat.data_ptr() == p1 and p0 is None
