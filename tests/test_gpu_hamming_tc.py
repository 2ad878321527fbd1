"""GPU parity of the tensor-core Hamming scans (tcgen05 kind::mxf4 over e2m1 nibbles - format 4,
the default - and kind::i8 over signed bytes - format 8): their top-2 keys must be bit-identical
to the integer-pipe kernel's (itself pinned to cv2.BFMatcher and the C oracle in
tests/test_gpu_hamming.py) and to the oracle directly.  Every test runs in both formats."""
import numpy as np
import pytest

from pybot_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[4, 8], ids=["mxf4", "i8"])
def tc_format(request):
    from pybot_b200 import ops
    before = ops.hamming_tc_format()
    ops.hamming_tc_set_format(request.param)
    yield request.param
    ops.hamming_tc_set_format(before)


def _keys_both(q, t, base=0):
    import torch
    from pybot_b200 import ops
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    signs = ops.hamming_tc_expand(td)
    a = ops.hamming_knn2_keys_tc(qd, signs, len(t), base)
    b = ops.hamming_knn2_keys(qd, td, base)
    torch.cuda.synchronize()
    return a.cpu().numpy(), b.cpu().numpy()


def test_expansion_layout(tc_format):
    """Train rows, tile by tile (128 rows).  Format 8: bit k of row r -> byte (r/128)*32768 + (k/16)*2048
    + (r%128)*16 + k%16 = -64 / +64.  Format 4: nibble k%2 of byte (r/128)*16384 + (k/32)*2048 + (r%128)*16
    + (k%32)/2 = e2m1 -1 (0xA) / +1 (0x2).  Rows past the end are zero."""
    import torch
    from pybot_b200 import ops
    rng = np.random.default_rng(0)
    n = 300
    d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    s = ops.hamming_tc_expand(torch.from_numpy(d).cuda()).cpu().numpy()
    rb = 256 if tc_format == 8 else 128
    assert ops.hamming_tc_row_bytes() == rb
    assert s.dtype == np.int8 and s.size == 512 * rb
    bits = np.unpackbits(d, axis=1, bitorder="little").astype(np.int16)                  # (n, 256)
    r, k = np.meshgrid(np.arange(n), np.arange(256), indexing="ij")
    rows = np.arange(512)
    if tc_format == 8:
        addr = (r // 128) * 32768 + (k // 16) * 2048 + (r % 128) * 16 + k % 16
        assert np.array_equal(s[addr], np.where(bits == 1, -64, 64))
        base = (rows // 128) * 32768 + (rows % 128) * 16
        assert not s[base[n:, None] + (np.arange(16) * 2048)[None, :]].any()             # pad rows are zero
    else:
        addr = (r // 128) * 16384 + (k // 32) * 2048 + (r % 128) * 16 + (k % 32) // 2
        nib = (s.view(np.uint8)[addr] >> (4 * (k % 2))) & 15
        assert np.array_equal(nib, np.where(bits == 1, 0xA, 0x2))
        base = (rows // 128) * 16384 + (rows % 128) * 16
        pad = base[n:, None, None] + (np.arange(8) * 2048)[None, :, None] + np.arange(16)[None, None, :]
        assert not s[pad].any()                                                          # pad rows are zero
    # appending: expanding only from the last whole block of 512 rows gives the same buffer
    d2 = rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    whole = ops.hamming_tc_expand(torch.from_numpy(d2).cuda())
    part = ops.hamming_tc_expand(torch.from_numpy(d2[:700]).cuda())
    grown = torch.zeros_like(whole)
    grown[:part.numel()] = part
    ops.hamming_tc_expand(torch.from_numpy(d2).cuda(), grown, first_row=512)
    assert torch.equal(grown, whole)


def test_queries_expand_with_the_opposite_sign(tc_format):
    """A query equal to a train row must come out at distance 0 and its complement at 256 (the two
    ends of the accumulator range of either format)."""
    import torch
    from pybot_b200 import ops
    rng = np.random.default_rng(5)
    t = rng.integers(0, 256, (1000, 32), dtype=np.uint8)
    q = np.stack([t[17], ~t[900], t[999]])
    a, b = _keys_both(q, t)
    assert np.array_equal(a, b)
    assert (a[0, 0] >> 32, a[0, 0] & 0xFFFFFFFF) == (0, 17) and (a[2, 0] >> 32, a[2, 0] & 0xFFFFFFFF) == (0, 999)
    t2 = np.repeat(t[900:901], 200, axis=0)                                              # every row at distance 256
    a2, b2 = _keys_both(q[1:2], t2)
    assert np.array_equal(a2, b2) and (a2[0, 0] >> 32, a2[0, 1] >> 32) == (256, 256)


@pytest.mark.parametrize("nq,nt", [(1, 1), (5, 3), (128, 128), (300, 1000), (512, 129), (513, 4096 + 77),
                                   (2000, 2000), (777, 50_000)])
def test_tc_keys_equal_integer_pipe_keys(nq, nt):
    q, t = syn.low_entropy_pair(seed=nq * 31 + nt, nq=nq, nt=nt)                         # tie-heavy descriptors
    a, b = _keys_both(q, t, base=12345)
    assert np.array_equal(a, b)


def test_tc_random_shapes_sweep(tc_format):
    """Twenty random (nq, nt) shapes - ragged last tiles, one to five query groups, splits of one to
    hundreds of tiles - on tie-heavy and on uniform descriptors, with a random global row base."""
    rng = np.random.default_rng(2024)
    for case in range(20):
        nq = int(rng.integers(1, 2600))
        nt = int(rng.integers(1, 400_000)) if case % 4 else int(rng.integers(1, 700))
        if case % 2:
            q, t = syn.low_entropy_pair(seed=case, nq=nq, nt=nt)
        else:
            q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
            t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        a, b = _keys_both(q, t, base=int(rng.integers(0, 1 << 30)))
        assert np.array_equal(a, b), (case, nq, nt)


def test_tc_against_c_oracle_and_ties():
    from oracle import c_oracle
    q, t = syn.low_entropy_pair(seed=9, nq=520, nt=150 * 256 + 200)
    t[30_000] = t[7]                                                                      # exact duplicate far apart
    q[3] = t[7]
    a, _ = _keys_both(q, t)
    idx, dist = (a & 0xFFFFFFFF).astype(np.int64), (a >> 32).astype(np.int32)
    ridx, rdist = c_oracle.hamming_knn2(q, t)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)
    assert list(idx[3]) == [7, 30_000] and list(dist[3]) == [0, 0]


def test_ties_across_every_split_resolve_to_the_lowest_rows(tc_format):
    """The CTAs of a format-4 scan share a bound of the second-best distance and skip rows beyond
    it: ties AT the bound must still resolve to the lowest global rows.  (a) every row of the
    database identical - every distance equal, the answer is rows 0 and 1 for every query;
    (b) the best match of a query duplicated into every split and column half."""
    rng = np.random.default_rng(11)
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    t = np.repeat(rng.integers(0, 256, (1, 32), dtype=np.uint8), 40_000, axis=0)
    a, b = _keys_both(q, t, base=7)
    assert np.array_equal(a, b)
    assert np.array_equal(a & 0xFFFFFFFF, np.tile(np.array([7, 8], dtype=np.int64), (300, 1)))
    t = rng.integers(0, 256, (60_000, 32), dtype=np.uint8)
    q[5] = t[1234]
    q[6] = t[1234]
    q[6, 0] ^= 1                                       # distance 1 to all the copies: ties at the second best too
    t[1234 + 61::61] = t[1234]                         # ~960 copies: several in every 128-row tile's neighbourhood
    a, b = _keys_both(q, t)
    assert np.array_equal(a, b)
    assert list(a[5] & 0xFFFFFFFF) == [1234, 1295] and list(a[5] >> 32) == [0, 0]
    assert list(a[6] & 0xFFFFFFFF) == [1234, 1295] and list(a[6] >> 32) == [1, 1]


def test_tc_uniform_random_large():
    """2000 x 2.5 M uniformly random descriptors (one 8-GPU shard of C3): every split / tile / slot path."""
    import torch
    from pybot_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randint(0, 256, (2000, 32), device="cuda", dtype=torch.uint8, generator=g)
    t = torch.randint(0, 256, (2_500_000, 32), device="cuda", dtype=torch.uint8, generator=g)
    signs = ops.hamming_tc_expand(t)
    a = ops.hamming_knn2_keys_tc(q, signs, t.shape[0], 0)
    b = ops.hamming_knn2_keys(q, t, 0)
    assert torch.equal(a, b)
