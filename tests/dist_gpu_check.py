"""NCCL parity check of the sharded paths (launched by tests/test_gpu_distributed.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29531 tests/dist_gpu_check.py

Every rank matches the replicated queries against its keyframe shard and evaluates its
chunk of BA observations; the merged results must be bit-identical to the single-GPU
run of the whole problem on the same device (keys, masks) / equal within 1e-12 (cost).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import ops, synthetic as syn  # noqa: E402
from pybot_b200.distributed import ShardedBA, ShardedKeyframeMatcher, shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per_kf, nkf = 1000, 61                                 # uneven shards
    q, db = syn.c3_database(nq=777, n_keyframes=nkf, per_kf=per_kf, n_revisit=3)
    db[40_123] = db[123]                                   # duplicate across shards: tie on the global row
    q[9] = db[123]
    lo, hi = shard_range(nkf, rank, world)
    shard = torch.from_numpy(db[lo * per_kf:hi * per_kf]).cuda()
    want = ops.knn_ratio_mutual(torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda(), 0.75)
    m = ShardedKeyframeMatcher(shard, lo * per_kf)         # fused peer-memory exchanges
    assert m.fused
    for rep in range(4):                                   # repeated frames: sequence numbers / buffer parity
        got = m.match(torch.from_numpy(q).cuda(), 0.75)
        for a, b in zip(got, want):
            assert torch.equal(a, b), "fused sharded matching differs from the single-GPU scan"
    got = ShardedKeyframeMatcher(shard, lo * per_kf, exchange="nccl").match(torch.from_numpy(q).cuda(), 0.75)
    for a, b in zip(got, want):
        assert torch.equal(a, b), "NCCL sharded matching differs from the single-GPU scan"
    q2 = q[:500].copy()                                    # a different query count re-creates the exchange
    got = m.match(torch.from_numpy(q2).cuda(), 0.75)
    for a, b in zip(got, ops.knn_ratio_mutual(torch.from_numpy(q2).cuda(), torch.from_numpy(db).cuda(), 0.75)):
        assert torch.equal(a, b)
    assert int(want[0][9, 0]) == 123 and int(want[0][9, 1]) == 40_123
    m.check()

    # the cv2-shaped facade over the sharded collection: every rank adds the WHOLE list of
    # keyframes and gets the global result, identical to the single-GPU facade
    from pybot_b200.vision.matcher import BFMatcher
    kfs = [db[i * per_kf:(i + 1) * per_kf] for i in range(nkf)]
    bf = BFMatcher(group=True)
    bf.add(kfs)
    got_rows = bf.knnMatch(q, k=2)
    one = BFMatcher()
    one.add(kfs)
    assert got_rows == one.knnMatch(q, k=2), "sharded BFMatcher.knnMatch differs from the single-GPU facade"
    assert (got_rows[9][0].imgIdx, got_rows[9][0].trainIdx, got_rows[9][1].imgIdx, got_rows[9][1].trainIdx) == (0, 123, 40, 123)

    # ... and the rest of knnMatch's signature over the sharded collection (k > 2, masks)
    kmask = (np.random.default_rng(2).random((len(q), len(db))) < 0.001).astype(np.uint8)
    assert bf.knnMatch(q, k=5) == one.knnMatch(q, k=5), "sharded knnMatch(k=5) differs from the single-GPU facade"
    assert bf.knnMatch(q, k=3, mask=kmask) == one.knnMatch(q, k=3, mask=kmask), "sharded masked knnMatch differs"

    prob = syn.c4_problem(n_cams=50, n_points=20_000)
    sb = ShardedBA(*prob)                                 # fused peer-memory exchange
    assert sb.exchange is not None
    cost = float(sb.evaluate().item())
    for _ in range(5):                                    # repeated calls: sequence numbers / slot parity
        assert float(sb.evaluate().item()) == cost
    for _ in range(4):                                    # ... and replayed from a captured CUDA graph
        assert sb.cost_host(graph=True) == cost
    assert float(sb.evaluate().item()) == cost            # eager again: the device-resident sequence kept in step
    nccl_cost = float(ShardedBA(*prob, exchange="nccl").evaluate().item())
    assert abs(nccl_cost - cost) <= 1e-12 * cost
    gathered = [None] * world
    dist.all_gather_object(gathered, cost)
    assert all(c == gathered[0] for c in gathered), "ranks disagree on the fused cost"
    whole = ops.BAProblemDevice(*prob).evaluate()
    ref = float(whole.cost.item())
    assert abs(cost - ref) <= 1e-12 * ref, (cost, ref)
    a, b = sb.range
    assert torch.equal(sb.problem.r, whole.r[a:b]) and torch.equal(sb.problem.Jc, whole.Jc[a:b])
    # a starved rank: rank 0 calls alone with a 50 ms bound -> poisoned outputs + RuntimeError,
    # never a hang and never stale data (fresh objects: the sequence numbers diverge here)
    from pybot_b200 import _ffi
    dist.barrier()
    m2 = ShardedKeyframeMatcher(shard, lo * per_kf)
    m2.match(torch.from_numpy(q).cuda(), 0.75)             # creates the exchange collectively
    sb2 = ShardedBA(*prob)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        _ffi.check(_ffi.lib().pbk_set_exchange_timeout_ms(50.0))
        try:
            m2.match_host(torch.from_numpy(q).cuda(), 0.75)
            raise AssertionError("a starved exchange must raise")
        except RuntimeError as e:
            assert "did not arrive" in str(e)
        try:
            sb2.cost_host()
            raise AssertionError("a starved cost exchange must raise")
        except RuntimeError as e:
            assert "did not arrive" in str(e)
        assert np.isnan(float(sb2.problem.cost.item()))
        _ffi.check(_ffi.lib().pbk_set_exchange_timeout_ms(2000.0))
    dist.barrier()
    if rank == 0:
        print("dist gpu check ok: world %d, valid matches %d, cost %.6e" % (world, int(want[4].sum()), cost))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
