"""Multi-GPU (NCCL) parity of the sharded paths; needs >= 2 visible GPUs, skipped otherwise
(the host-side sharding logic is covered on CPU by tests/test_distributed_cpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_sharded_paths_match_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "dist gpu check ok" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


def _emulated_matchers(db, per_kf, nkf, nq, world):
    """`world` ranks of the fused-exchange matcher on ONE GPU (one stream per rank, plain device
    buffers as the peers' exchange buffers; over NVLink the protocol is the same)."""
    import torch
    from pybot_b200.distributed import PeerKeyExchange, ShardedKeyframeMatcher, shard_range
    xs = PeerKeyExchange.emulated(nq, world)
    ms = []
    for r in range(world):
        lo, hi = shard_range(nkf, r, world)
        shard = torch.from_numpy(db[lo * per_kf:hi * per_kf]).cuda()
        ms.append(ShardedKeyframeMatcher(shard, lo * per_kf).use_exchange(xs[r]))
    # Load every kernel of the fused path before the ranks run concurrently: in ONE process the
    # first launch of a kernel (lazy module loading) can block the host behind a kernel that is
    # spinning for a peer that has not been launched yet.  One call per rank with a 1 ms bound
    # (they time out against each other), then clear the status words.
    from pybot_b200 import _ffi
    _ffi.check(_ffi.lib().pbk_set_exchange_timeout_ms(1.0))
    qw = torch.zeros((nq, 32), dtype=torch.uint8, device="cuda")
    for m in ms:
        m.match(qw, 0.75)
    torch.cuda.synchronize()
    for x in xs:
        x.status.word.zero_()
    _ffi.check(_ffi.lib().pbk_set_exchange_timeout_ms(2000.0))
    return ms, [torch.cuda.Stream() for _ in range(world)]


@pytest.fixture
def restore_tc_format():
    from pybot_b200 import ops
    before = ops.hamming_tc_format()
    yield
    ops.hamming_tc_set_format(before)


@pytest.mark.parametrize("engine", ["popc", "tc", "tc8"])
@pytest.mark.parametrize("world", [2, 4])
def test_fused_key_and_row_exchange_protocol_on_one_gpu(world, engine, monkeypatch, restore_tc_format):
    """Both fused exchanges (top-2 keys, matched rows) with all ranks on one device: every rank's
    result is bit-identical to the single-GPU scan of the whole database, frame after frame."""
    import torch
    from pybot_b200 import ops, synthetic as syn
    monkeypatch.setenv("PBK_HAMMING_ENGINE", engine[:2] if engine != "popc" else engine)
    ops.hamming_tc_set_format(8 if engine == "tc8" else 4)
    per_kf, nkf, nq = 1000, 37, 777
    q, db = syn.c3_database(nq=nq, n_keyframes=nkf, per_kf=per_kf, n_revisit=3)
    db[30_123] = db[123]                                   # duplicate across shards: tie on the global row
    q[9] = db[123]
    qd = torch.from_numpy(q).cuda()
    want = ops.knn_ratio_mutual(qd, torch.from_numpy(db).cuda(), 0.75)
    ms, streams = _emulated_matchers(db, per_kf, nkf, nq, world)
    torch.cuda.synchronize()
    for rep in range(3):
        outs = []
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                outs.append(ms[r].match(qd, 0.75))
        torch.cuda.synchronize()
        for r in range(world):
            ms[r].check()
            for a, b in zip(outs[r], want):
                assert torch.equal(a, b), "rank %d differs from the single-GPU scan" % r
    assert int(want[0][9, 0]) == 123 and int(want[0][9, 1]) == 30_123


def test_fused_exchange_timeout_raises():
    """A starved rank: the waiting rank's outputs are all 'no match' (never stale exchange
    words), its status word is set, and the wrapper raises RuntimeError."""
    import torch
    from pybot_b200 import _ffi, synthetic as syn
    per_kf, nkf, nq = 500, 8, 300
    q, db = syn.c3_database(nq=nq, n_keyframes=nkf, per_kf=per_kf, n_revisit=2)
    ms, streams = _emulated_matchers(db, per_kf, nkf, nq, 2)
    qd = torch.from_numpy(q).cuda()
    _ffi.check(_ffi.lib().pbk_set_exchange_timeout_ms(20.0))
    try:
        idx, dist, r_ok, m_ok, valid = ms[0].match(qd, 0.75)          # rank 1 never calls
        torch.cuda.synchronize()
        assert bool((idx == -1).all()) and not bool(valid.any())
        with pytest.raises(RuntimeError, match="did not arrive"):
            ms[0].check()
        with pytest.raises(RuntimeError, match="did not arrive"):
            ms[0].match_host(qd, 0.75)
    finally:
        _ffi.check(_ffi.lib().pbk_set_exchange_timeout_ms(2000.0))
