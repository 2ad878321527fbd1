"""Multi-GPU sharding for the two paths that partition.

One process per GPU, ``torch.distributed`` (NCCL over NVLink on the B200 box,
gloo in the CPU tests) for the plumbing.

* Loop-closure matching: the keyframe database is split into contiguous
  keyframe ranges, one per rank; queries are replicated.  Every rank computes
  its top-2 packed keys ``(distance << 32 | global_row)``; the 32 KB results
  reach every rank and are merged with the same unsigned min the kernel uses, so
  the result is bit-identical to the single-GPU scan.  For the mutual check the
  owner of each matched train row sends it to every rank, then every rank runs
  the tiny reverse pass locally.  Both exchanges are FUSED into the producing
  kernels (stores into the peers' symmetric-memory buffers over NVLink + a
  sequence number the consumer waits for); the NCCL variant (all-gather of the
  keys, sum-all-reduce of an owner-filled row buffer) remains as the alternative
  and is what the CPU tests drive.
* BA residual/Jacobian: observations are split into contiguous chunks; cameras
  and points are replicated; the only exchange is the sum-all-reduce of the fp64
  cost, which is FUSED into the streaming kernel: its last CTA stores the rank's
  cost into every peer's symmetric-memory buffer over NVLink and adds the peers'
  values in rank order (no NCCL call on the data path; an NCCL all-reduce remains
  as the alternative and is what the CPU tests drive).

Per-frame reprojection / projection do not shard (replicas only).

The compute steps are injected through ``kernels`` so the CPU tests can drive
the sharding and collective logic with the oracle; the default is the CUDA
library and it raises without a GPU.
"""
import numpy as np

from . import _ffi


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) of n_items for `rank` of `world`."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _CudaMatchKernels(object):
    """Default compute backend: libpybot_b200.so."""

    def knn2(self, q, t, train_base, signs=None):
        from . import ops
        return ops.hamming_scan_keys(q, t, train_base, signs)

    def merge(self, partial):
        from . import ops
        return ops.hamming_merge(partial)

    def gather_best(self, t, keys, train_base):
        from . import ops
        return ops.hamming_gather_best(t, keys, train_base)

    def ratio_mutual(self, keys, rev, ratio):
        from . import ops
        return ops.hamming_ratio_mutual(keys, rev, ratio)

    def knnk(self, q, t, k, mask, train_base):
        from . import ops
        return ops.hamming_knnk(q, t, k, mask, train_base)


class PeerKeyExchange(object):
    """Peer-mapped exchange buffer of the fused matcher exchanges
    (pbk_hamming_knn2_dist / pbk_hamming_gather_best_dist) for `nq` queries: a
    torch symmetric-memory allocation whose per-rank addresses are handed to the
    kernels.  Collective: every rank of `group` must construct it.

    `status` is the word the kernels set when a peer does not arrive within the
    exchange timeout (pbk_set_exchange_timeout_ms); `check()` raises on it."""

    def __init__(self, nq, group=None, _emulated=None):
        import ctypes
        import torch
        self.nq = int(nq)
        self.seq = 0
        self.status = _ffi.StatusWord()
        if _emulated is not None:
            # (rank, world, buffers): `world` ranks of ONE process on one device, each with its
            # own stream - the single-GPU tests of the exchange protocol
            self.rank, self.world, bufs = _emulated
            self.buf = bufs[self.rank]
            self.done = torch.zeros(4, dtype=torch.int32, device=self.buf.device)
            self.ptrs = (ctypes.c_uint64 * self.world)(*[int(b.data_ptr()) for b in bufs])
            return
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        words = int(_ffi.lib().pbk_hamming_exchange_words(self.nq, self.world))
        self.buf = symm_mem.empty(max(words, 64), dtype=torch.int64, device=dev)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group=group)
        self.done = torch.zeros(4, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        dist.barrier(group)                                  # every buffer is zero before the first store
        self.ptrs = (ctypes.c_uint64 * self.world)(*[int(p) for p in self.handle.buffer_ptrs])

    @classmethod
    def emulated(cls, nq, world):
        """`world` exchanges that share plain device buffers of the current device."""
        import torch
        words = int(_ffi.lib().pbk_hamming_exchange_words(int(nq), int(world)))
        bufs = [torch.zeros(max(words, 64), dtype=torch.int64, device="cuda") for _ in range(world)]
        return [cls(nq, _emulated=(r, world, bufs)) for r in range(world)]

    def check(self):
        self.status.check("fused key/row exchange of the sharded matcher")


class ShardedKeyframeMatcher(object):
    """Holds this rank's shard of the keyframe database and matches replicated
    queries against the whole database.

    exchange: "p2p" fuses both exchanges (top-2 keys, matched rows) into the
    kernels over NVLink peer memory (default on the CUDA path when world > 1);
    "nccl" all-gathers / all-reduces with NCCL between the kernels (also what the
    gloo CPU tests drive)."""

    def __init__(self, shard, shard_base, group=None, kernels=None, exchange=None):
        """shard: (n_local, 32) uint8 tensor on this rank's device;
        shard_base: global row index of shard[0]."""
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.shard = shard
        self.shard_base = int(shard_base)
        cuda_path = kernels is None
        self.kernels = kernels if kernels is not None else _CudaMatchKernels()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if exchange is None:
            exchange = "p2p" if (cuda_path and self.world > 1) else "nccl"
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.fused = exchange == "p2p" and self.world > 1
        self._px = None
        # large shards are scanned on the tensor cores: keep the expanded copy resident
        # (in the format selected NOW - ops.hamming_tc_format(); scans check it has not changed)
        self.signs = None
        self.signs_format = 0
        if cuda_path and int(shard.shape[0]) > 0:
            from . import ops
            if ops.hamming_engine(int(shard.shape[0])) == "tc":
                self.signs_format = ops.hamming_tc_format()
                self.signs = ops.hamming_tc_expand(shard)

    def use_exchange(self, px):
        """Adopt a prepared PeerKeyExchange (e.g. an emulated one): rank / world come from it."""
        self._px, self.rank, self.world, self.fused = px, px.rank, px.world, True
        return self

    def check(self):
        """Raises RuntimeError if a fused exchange of an EARLIER, finished call timed out
        (call after synchronising; match() also checks at its start)."""
        if self._px is not None:
            self._px.check()

    def _check_format(self):
        if self.signs is not None:
            from . import ops
            if ops.hamming_tc_format() != self.signs_format:
                raise RuntimeError("the tensor-core format changed after this shard was expanded "
                                   "(expanded in %d, now %d)" % (self.signs_format, ops.hamming_tc_format()))

    def local_keys(self, q):
        self._check_format()
        if self.signs is not None:
            return self.kernels.knn2(q, self.shard, self.shard_base, self.signs)
        return self.kernels.knn2(q, self.shard, self.shard_base)

    def global_keys(self, q):
        """(nq,2) global top-2 keys, identical on every rank."""
        T = _ffi.torch()
        local = self.local_keys(q)
        if self.world == 1:
            return local
        # flat (world*nq, 2) output: the layout both NCCL and gloo accept
        gathered = T.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                           device=local.device)
        self.dist.all_gather_into_tensor(gathered, local.contiguous(), group=self.group)
        return self.kernels.merge(gathered.view((self.world,) + tuple(local.shape)))

    def _match_fused(self, q, ratio, mutual, events=None):
        """knn2 -> [keys pushed to all peers] -> world merge -> gather -> [rows pushed
        to all peers] -> reverse knn2 -> ratio/mutual, without a collective call."""
        import ctypes
        from . import ops
        T = _ffi.torch()
        nq, nt = int(q.shape[0]), int(self.shard.shape[0])
        if self._px is None or self._px.nq != nq:
            self._px = PeerKeyExchange(nq, self.group)               # collective, once per query count
        px = self._px
        px.check()                                                   # a timeout of the previous call
        self._check_format()
        px.seq += 1
        L, s = _ffi.lib(), _ffi.stream_ptr()
        splits, nbytes = ctypes.c_int(0), ctypes.c_size_t(0)
        tc = self.signs is not None
        _ffi.check((L.pbk_hamming_tc_plan if tc else L.pbk_hamming_plan)(nq, nt, ctypes.byref(splits), ctypes.byref(nbytes)))
        scratch = T.empty((max(int(nbytes.value), 16),), dtype=T.uint8, device=q.device)
        keys = T.empty((nq, 2), dtype=T.int64, device=q.device)
        if events is not None:
            events[0].record()
        if tc:       # the shard scan on the tensor cores; same fused key exchange
            _ffi.check(L.pbk_hamming_tc_knn2_dist(_ffi.ptr(q), nq, _ffi.ptr(self.signs), nt, self.shard_base,
                                                  _ffi.ptr(keys), _ffi.ptr(scratch), px.ptrs, px.rank, px.world, px.seq,
                                                  _ffi.ptr(px.done), px.status.ptr, s))
        else:
            _ffi.check(L.pbk_hamming_knn2_dist(_ffi.ptr(q), nq, _ffi.ptr(self.shard) if nt else None, nt,
                                               self.shard_base, _ffi.ptr(keys), _ffi.ptr(scratch), px.ptrs, px.rank,
                                               px.world, px.seq, _ffi.ptr(px.done), px.status.ptr, s))
        if events is not None:
            events[1].record()
        rev = None
        if mutual:
            rows = T.empty((nq, 32), dtype=T.uint8, device=q.device)
            _ffi.check(L.pbk_hamming_gather_best_dist(_ffi.ptr(self.shard) if nt else None, nt, self.shard_base,
                                                      _ffi.ptr(keys), nq, _ffi.ptr(rows), px.ptrs, px.rank,
                                                      px.world, px.seq, _ffi.ptr(px.done), px.status.ptr, s))
            rev = ops.hamming_knn2_keys(rows, q, 0)
        return ops.hamming_ratio_mutual(keys, rev, ratio)

    def match(self, q, ratio=0.75, mutual=True, events=None):
        """idx (nq,2) global rows, dist (nq,2), ratio_ok, mutual_ok, valid.
        `events` (fused path): (start, end) CUDA events recorded around the scan of
        this rank's shard including its key exchange."""
        if self.fused and int(q.shape[0]) > 0:
            return self._match_fused(q, ratio, mutual, events)
        keys = self.global_keys(q)
        rev = None
        if mutual:
            rows = self.kernels.gather_best(self.shard, keys, self.shard_base)
            if self.world > 1:
                self.dist.all_reduce(rows, op=self.dist.ReduceOp.SUM, group=self.group)
            rev = self.kernels.knn2(rows, q, 0)
        return self.kernels.ratio_mutual(keys, rev, ratio)

    def knnk_local_keys(self, q, k, mask=None):
        """(nq, k) int64 keys (distance << 32 | global row; INT64_MAX = none) of this shard's k best
        rows per query.  mask: (nq, rows of the WHOLE database) or None; the shard reads its own
        column range."""
        T = _ffi.torch()
        nl = int(self.shard.shape[0])
        if mask is not None:
            mask = mask[:, self.shard_base:self.shard_base + nl]
            mask = mask.contiguous() if isinstance(mask, T.Tensor) else np.ascontiguousarray(mask)
        idx, dist, _ = self.kernels.knnk(q, self.shard, int(k), mask, self.shard_base)
        idx, dist = T.as_tensor(idx), T.as_tensor(dist)
        return T.where(idx >= 0, (dist.to(T.int64) << 32) | idx, T.full_like(idx, T.iinfo(T.int64).max))

    @staticmethod
    def knnk_merge(keys, k):
        """keys (nq, parts * k) -> (idx (nq,k) global rows, dist (nq,k) int32, count (nq,) int32):
        the k smallest by (distance, global row) - BFMatcher's order over the whole collection."""
        T = _ffi.torch()
        keys = T.sort(keys, dim=1).values[:, :int(k)]
        found = keys != T.iinfo(T.int64).max
        return (T.where(found, keys & 0xFFFFFFFF, T.full_like(keys, -1)),
                T.where(found, keys >> 32, T.full_like(keys, -1)).to(T.int32),
                found.sum(1).to(T.int32))

    def knnk(self, q, k, mask=None):
        """knnMatch's general form over the sharded database: any k (<= ops.KNNK_MAX_K), optional
        per-pair mask (every rank passes the same (nq, total rows) matrix).  Returns (idx, dist,
        count), identical on every rank and equal to the single-GPU call: each rank's k best
        (pbk_hamming_knnk), ONE all-gather of the (nq, k) keys, the k smallest of the world x k
        candidates per query.  Not the hot path: k = 2 without a mask is match()."""
        T = _ffi.torch()
        nq, k = int(q.shape[0]), int(k)
        keys = self.knnk_local_keys(q, k, mask)
        if self.world > 1:
            gathered = T.empty((self.world * nq, k), dtype=T.int64, device=keys.device)
            self.dist.all_gather_into_tensor(gathered, keys.contiguous(), group=self.group)
            keys = gathered.view(self.world, nq, k).permute(1, 0, 2).reshape(nq, self.world * k)
        return self.knnk_merge(keys, k)

    def match_host(self, q, ratio=0.75, mutual=True):
        """match() with numpy results; raises RuntimeError if a peer timed out in this call."""
        r = self.match(q, ratio, mutual)
        out = list(r.to_host()) if hasattr(r, "to_host") else [_ffi.to_host(x) for x in r]
        self.check()
        return out


class PeerCostExchange(object):
    """Peer-mapped exchange buffer for the fused BA cost all-reduce
    (pbk_ba_evaluate / pbk_ba_residual_jacobian_dist): a torch symmetric-memory allocation whose
    per-rank addresses are handed to the kernel, which stores / polls them over
    NVLink itself.  Collective: every rank of `group` must construct it.  The call
    sequence lives on the device (behind the problem's cost partials), so the call can
    be replayed from a CUDA graph."""

    def __init__(self, group=None, _emulated=None):
        import ctypes
        import torch
        self.status = _ffi.StatusWord()
        if _emulated is not None:
            self.rank, self.world, bufs = _emulated
            self.buf = bufs[self.rank]
            self.ptrs = (ctypes.c_uint64 * self.world)(*[int(b.data_ptr()) for b in bufs])
            return
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm_mem.empty(max(64, 4 * self.world), dtype=torch.int64, device=dev)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group=group)
        torch.cuda.synchronize()
        dist.barrier(group)                                  # every buffer is zero before the first store
        self.ptrs = (ctypes.c_uint64 * self.world)(*[int(p) for p in self.handle.buffer_ptrs])

    @classmethod
    def emulated(cls, world):
        """`world` exchanges sharing plain device buffers of the current device (tests)."""
        import torch
        bufs = [torch.zeros(max(64, 4 * world), dtype=torch.int64, device="cuda") for _ in range(world)]
        return [cls(_emulated=(r, world, bufs)) for r in range(world)]

    def check(self):
        self.status.check("fused cost all-reduce of the sharded BA")


class ShardedBA(object):
    """This rank's contiguous chunk of the observations of a BA problem.

    exchange: "p2p" fuses the cost all-reduce into the kernel over NVLink peer
    memory (default on the CUDA path when world > 1), "nccl" uses a one-element
    NCCL all-reduce after the kernel (also what the gloo CPU tests drive)."""

    def __init__(self, rvecs, tvecs, K, points, obs_cam, obs_pt, obs_uv, group=None,
                 problem_factory=None, exchange=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        lo, hi = shard_range(len(obs_cam), self.rank, self.world)
        self.range = (lo, hi)
        cuda_path = problem_factory is None
        if problem_factory is None:
            from .ops import BAProblemDevice as problem_factory
        self.problem = problem_factory(rvecs, tvecs, K, points, obs_cam[lo:hi], obs_pt[lo:hi],
                                       obs_uv[lo:hi])
        if exchange is None:
            exchange = "p2p" if (cuda_path and self.world > 1) else "nccl"
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.exchange = PeerCostExchange(group) if (exchange == "p2p" and self.world > 1) else None

    def evaluate(self, graph=False):
        """Evaluates the local residuals/Jacobians (they stay on the owning
        rank) and returns the global cost tensor (1,) float64, all-reduced.
        graph=True (CUDA path) replays the evaluation from a captured CUDA graph."""
        if self.exchange is not None:
            self.exchange.check()                              # a timeout of an earlier call
            if graph:
                self.problem.evaluate_graphed(exchange=self.exchange)
            else:
                self.problem.evaluate(exchange=self.exchange)  # cost is already global
            return self.problem.cost
        if graph and self.world == 1:
            self.problem.evaluate_graphed()
            return self.problem.cost
        self.problem.evaluate()
        cost = self.problem.cost
        if self.world > 1:
            cost = cost.clone()
            self.dist.all_reduce(cost, op=self.dist.ReduceOp.SUM, group=self.group)
        return cost

    def cost_host(self, graph=False):
        """evaluate() -> float; raises RuntimeError if a peer timed out in this call."""
        v = float(self.evaluate(graph=graph).item())
        if self.exchange is not None:
            self.exchange.check()
        return v
