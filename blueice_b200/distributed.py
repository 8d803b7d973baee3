"""Multi-GPU evaluation: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch).

The reference has no distributed evaluation path at all (its only parallelism, blueice/parallel.py,
farms out anchor-model construction).  The hot path shards in two natural ways:

  point / toy sharding   rank r evaluates a contiguous slice of the P parameter points on a replicated
                         dataset; NO collective inside the evaluation, one all_gather of the P float64
                         results at the end.                                   -> PointShardedLikelihood
  event sharding         rank r holds a contiguous, superblock-aligned slice of the events; every rank
                         evaluates sum_{i in shard} log f_i for all P points; the ONE exchange step is an
                         all_gather of P doubles per rank followed by a sum in FIXED RANK ORDER (so the
                         result does not depend on timing or collective algorithm); -sum(mu) and the
                         priors are added once.                                -> EventShardedLikelihood

  toy sharding           rank r generates (Model.simulate_toys with first_toy = its slice start: counter-based
                         random numbers keyed by the GLOBAL toy id, so the toys do not depend on the number of
                         ranks) and evaluates a contiguous slice of the T toys; no collective inside the
                         evaluation, one all_gather of the T results.         -> ToyShardedLikelihood

Communication tensors live on the GPU for the nccl backend and on the host for gloo (CPU tests).
"""
import numpy as np

from . import _cabi


def shard_bounds(n, world_size, align=1):
    """Contiguous split of range(n) into world_size slices whose starts are multiples of `align`.

    Returns a list of (start, stop).  Slices differ by at most one `align` block; trailing ranks may
    be empty when n is small."""
    n_blocks = -(-n // align) if n > 0 else 0
    base, extra = divmod(n_blocks, world_size)
    bounds, start = [], 0
    for r in range(world_size):
        blocks = base + (1 if r < extra else 0)
        stop = min(start + blocks * align, n)
        bounds.append((start, stop))
        start = stop
    return bounds


def _dist():
    import torch.distributed as dist
    return dist


def _comm_device(group=None):
    import torch
    backend = _dist().get_backend(group)
    if backend == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def all_gather_rows(local, group=None):
    """all_gather of equally shaped float64 arrays -> array [world, *local.shape] (same on every rank)."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(local, dtype=np.float64))).to(_comm_device(group))
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.stack([o.cpu().numpy() for o in out])


def gather_concat(local, counts, group=None):
    """Concatenate per-rank 1-d float64 arrays of the given lengths (all ranks get the full array)."""
    width = max(max(counts), 1)
    padded = np.zeros(width)
    padded[:len(local)] = local
    rows = all_gather_rows(padded, group)
    return np.concatenate([rows[r, :c] for r, c in enumerate(counts)])


def rank_ordered_sum(local, group=None):
    """sum over ranks of equally shaped arrays, accumulated as ((r0 + r1) + r2) + ... on every rank."""
    rows = all_gather_rows(local, group)
    acc = rows[0].copy()
    for r in range(1, len(rows)):
        acc = acc + rows[r]
    return acc


class PeerGather(object):
    """The exchange step of a sharded evaluation over NVLink peer memory (bi_peer_exchange): ONE kernel launch stores this
    rank's n float64 values into peer-mapped buffers of every rank (torch symmetric memory), releases its flag on every
    rank, waits for every rank's flag and then either copies the gathered rows out (gather) or adds them up in rank order
    (reduce).  No collective launch, no host round trip; the epoch counter lives on the device, so the launch can be
    captured in a CUDA graph together with the evaluation it follows.  Device tensors in, device tensors out, all on the
    current stream.  Falls back to NCCL all_gather_into_tensor when symmetric memory is unavailable."""

    def __init__(self, n, group=None):
        import ctypes
        import torch
        dist = _dist()
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n = max(int(n), 1)
        self.device = torch.device('cuda', torch.cuda.current_device())
        self.lib = _cabi.load()
        self.fallback = None
        self.words = int(self.lib.bi_peer_exchange_words(self.world, self.n))
        try:
            import torch.distributed._symmetric_memory as symm_mem
            self.buffer = symm_mem.empty(self.words, dtype=torch.float64, device=self.device)
            self.buffer.zero_()
            self.handle = symm_mem.rendezvous(self.buffer, self.group)
            self.peer_ptrs = np.ascontiguousarray(np.array([int(p) for p in self.handle.buffer_ptrs], dtype=np.uint64))
            self.handle.barrier(channel=0)                         # every rank's flags are zero before anybody signals
        except Exception as exc:                                    # no peer mapping on this system: NCCL
            self.fallback = repr(exc)
            self.buffer = None
        self.out = torch.zeros(self.world * self.n, dtype=torch.float64, device=self.device)      # gathered rows
        self.total = torch.zeros(self.n, dtype=torch.float64, device=self.device)                 # reduced values
        self.launches = 0
        self._parity = 0
        self._ctypes = ctypes

    def _stream(self):
        import torch
        return self._ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def barrier(self):
        """Cross-GPU barrier on the current stream (signal pads; NCCL barrier in the fallback)."""
        if self.fallback is not None:
            _dist().barrier(group=self.group)
        else:
            self.handle.barrier(channel=0)

    def error(self):
        """0, or the epoch at which a peer failed to arrive in time (synchronises the stream)."""
        if self.fallback is not None:
            return 0
        import torch
        words = self.buffer.view(torch.int64)
        return int(words[2 * self.world * self.n + self.world + 3].item())

    def _check(self, local):
        if local.numel() > self.n or (self.fallback is not None and local.numel() != self.n):
            raise ValueError("PeerGather was built for %d values per rank, got %d" % (self.n, local.numel()))

    def gather(self, local, out=None, status=None, status_out=None):
        """local: device tensor of <= n float64 -> tensor [world, n] holding every rank's values (rows of ranks that
        passed fewer than n values keep stale tails).  out: where the kernel writes the rows (default: a device buffer);
        a pinned host tensor of world * n float64 makes the kernel deliver them to the host itself, without a copy.
        status / status_out (int32 tensors of len(local)): the same launch copies this rank's point status words, e.g. to
        pinned host memory."""
        self._check(local)
        if self.fallback is not None:
            _dist().all_gather_into_tensor(self.out, local.contiguous(), group=self.group)
            if status_out is not None:
                status_out.copy_(status, non_blocking=True)
            if out is not None:
                out.copy_(self.out, non_blocking=True)
                return out.view(self.world, self.n)
            return self.out.view(self.world, self.n)
        dst = self.out if out is None else out
        if status_out is not None:
            _cabi.check(self.lib.bi_peer_gather_status(_cabi.dev_ptr(local), local.numel(), self.n,
                                                       _cabi.host_ptr(self.peer_ptrs), self.world, self.rank,
                                                       _cabi.dev_ptr(status), _cabi.dev_ptr(status_out), _cabi.dev_ptr(dst),
                                                       self._stream()), "bi_peer_gather_status")
        else:
            _cabi.check(self.lib.bi_peer_exchange(_cabi.dev_ptr(local), local.numel(), self.n, _cabi.host_ptr(self.peer_ptrs),
                                                  self.world, self.rank, 0, None, None, _cabi.dev_ptr(dst), self._stream()),
                        "bi_peer_exchange")
        self.launches += 1
        return dst.view(self.world, self.n)

    def reduce(self, local, musum=None, status=None, out=None):
        """Sum over ranks, accumulated in rank order ((r0 + r1) + r2) + ... on every rank -> device tensor [len(local)];
        with musum / status (device tensors of the same length): -musum + total, -inf where status != 0
        (the event-sharded log likelihood without priors)."""
        import torch
        self._check(local)
        m = local.numel()
        if self.fallback is not None:
            if m != self.n:
                raise ValueError("NCCL fallback needs exactly %d values" % self.n)
            _dist().all_gather_into_tensor(self.out, local.contiguous(), group=self.group)
            rows = self.out.view(self.world, self.n)
            acc = rows[0].clone()
            for r in range(1, self.world):
                acc = acc + rows[r]
            if musum is not None:
                acc = -musum + acc
            if status is not None:
                acc = torch.where(status != 0, torch.full_like(acc, -float('inf')), acc)
            self.total.copy_(acc)
            if out is not None:
                out[:m].copy_(self.total[:m], non_blocking=True)
                return out[:m]
            return self.total[:m]
        dst = self.total if out is None else out                   # out: e.g. a pinned host tensor (no copy afterwards)
        _cabi.check(self.lib.bi_peer_exchange(_cabi.dev_ptr(local), m, self.n, _cabi.host_ptr(self.peer_ptrs),
                                              self.world, self.rank, 1, _cabi.dev_ptr(musum), _cabi.dev_ptr(status),
                                              _cabi.dev_ptr(dst), self._stream()), "bi_peer_exchange")
        self.launches += 1
        return dst[:m]

    def broadcast(self, local):
        """Stores only (bi_peer_broadcast): this rank's values go to slot (parity, rank) of every rank without waiting for
        anybody; the rows are complete after the next barrier().  Host-tracked parity: do NOT mix with gather() / reduce()
        on the same object, and let at most one un-waited broadcast overtake a reader.  Returns the [world, n] view the
        rows arrive in."""
        self._check(local)
        if self.fallback is not None:
            return self.gather(local)
        self._parity ^= 1
        half = self.world * self.n
        _cabi.check(self.lib.bi_peer_broadcast(_cabi.dev_ptr(local), local.numel(), _cabi.host_ptr(self.peer_ptrs),
                                               self.world, self._parity * half + self.rank * self.n, self._stream()),
                    "bi_peer_broadcast")
        self.launches += 1
        return self.buffer[self._parity * half:(self._parity + 1) * half].view(self.world, self.n)


class PointShardedLikelihood(object):
    """Shard the points of ll.batch over the ranks of `group`; every rank returns the full result.

    `ll` is a prepared likelihood with the SAME data set on every rank (replicated templates / anchor
    tensor).  Any object with a `.batch(params, names, livetime_days=None)` method works."""

    def __init__(self, ll, group=None):
        self.ll = ll
        self.group = group
        self._gathers = {}
        self._splits = {}
        self._nccl = None

    def _device_gather_engine(self, n_rows):
        """The ll's fused unbinned engine with a PeerGather for n_rows rows per rank, or None (binned likelihoods,
        gloo, engines without the fused path): then the results are gathered on the host."""
        if self._nccl is None:
            self._nccl = _dist().get_backend(self.group) == 'nccl'
        if not self._nccl:
            return None
        engine = getattr(self.ll, '_engine', None)
        if engine is None or not hasattr(engine, 'evaluate_fused') or not engine.uses_mma():
            return None
        pg = self._gathers.get(n_rows)
        if pg is None:
            pg = self._gathers[n_rows] = PeerGather(n_rows, self.group)
        engine.peer_gather, engine.peer_mode = pg, 'gather'
        return engine

    def _split(self, n_points):
        """(lo, hi, counts, n_rows) of this rank for a table of n_points rows (cached per table length)."""
        split = self._splits.get(n_points)
        if split is None:
            dist = _dist()
            rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
            bounds = shard_bounds(n_points, world)
            counts = [b - a for a, b in bounds]
            if len(self._splits) > 64:
                self._splits.clear()
            split = self._splits[n_points] = (bounds[rank][0], bounds[rank][1], counts, max(counts) if counts else 0)
        return split

    def batch(self, params, names=None, livetime_days=None):
        params = np.asarray(params, dtype=np.float64)
        lo, hi, counts, n_rows = self._split(len(params))
        engine = self._device_gather_engine(n_rows) if n_rows else None
        if engine is None:
            local = self.ll.batch(params[lo:hi], names, livetime_days=livetime_days) if hi > lo else np.zeros(0)
            return gather_concat(np.asarray(local, dtype=np.float64), counts, self.group)
        # device-side gather: every rank evaluates n_rows rows (short shards repeat their last row, or row 0 of the
        # table when they are empty); the logl rows (without priors) of all ranks arrive over NVLink before the D2H
        rows = params[lo:hi] if hi > lo else params[:1]
        if len(rows) < n_rows:
            rows = np.vstack([rows, np.repeat(rows[-1:], n_rows - len(rows), axis=0)])
        try:
            self.ll.batch(rows, names, livetime_days=livetime_days)        # local checks ('error' mode) + the gather
            gathered = engine.last_gathered
        finally:
            engine.peer_gather = None
        # (gathered is either a landing buffer handed over to this call -- no copy -- or a view of the engine's scratch
        # buffer, which the next call overwrites)
        owned = getattr(engine, 'last_gathered_owned', False)
        device_ll = (gathered.reshape(-1) if owned else gathered.reshape(-1).copy()) if min(counts) == n_rows else \
            np.concatenate([gathered[r, :c] for r, c in enumerate(counts)])
        has_priors = any(p is not None for _, p, _ in self.ll.shape_parameters.values()) or \
            any(p is not None for p in self.ll.rate_parameters.values())
        if not has_priors:
            return device_ll
        zs, mult = self.ll._rows_from_params(params, names)
        priors = self.ll._prior_sum(zs, mult)
        return np.where(np.isneginf(device_ll), -np.inf, priors + device_ll)


class ToyShardedLikelihood(object):
    """Toy Monte Carlos sharded over the ranks of `group` (BASELINE config 4: 1e6 toys x 1e3 events over 8 GPUs).

    `ll` is a prepared UnbinnedLogLikelihood (same model on every rank).  simulate(n_toys, ...) lets every rank
    generate and load ITS slice of the toys on its own GPU; batch_toys(params[T, k]) evaluates toy t at params[t]
    on the rank that owns it and returns all T results on every rank."""

    def __init__(self, ll, group=None):
        self.ll = ll
        self.group = group
        self.n_toys = 0
        self.bounds = None
        self._gather = None

    def _rank_world(self):
        dist = _dist()
        return dist.get_rank(self.group), dist.get_world_size(self.group)

    def simulate(self, n_toys, rate_multipliers=None, livetime_days=None, seed=0, mus=None):
        """Generate toys 0 .. n_toys - 1 across the ranks (rank r: its contiguous slice) and load them."""
        rank, world = self._rank_world()
        self.n_toys = int(n_toys)
        self.bounds = shard_bounds(self.n_toys, world)
        lo, hi = self.bounds[rank]
        if mus is not None and np.ndim(mus) == 2:
            mus = np.asarray(mus)[lo:hi]
        toys = self.ll.base_model.simulate_toys(hi - lo, rate_multipliers, livetime_days, seed=seed, first_toy=lo,
                                                mus=mus)
        self.ll.set_toy_data(toys)
        return toys

    def set_local_toys(self, n_toys_total, datasets, offsets=None):
        """Load this rank's slice of externally produced toys (slice = shard_bounds(n_toys_total, world)[rank])."""
        rank, world = self._rank_world()
        self.n_toys = int(n_toys_total)
        self.bounds = shard_bounds(self.n_toys, world)
        self.ll.set_toy_data(datasets, offsets)

    def batch_toys(self, params, names=None, livetime_days=None):
        rank, _ = self._rank_world()
        params = np.asarray(params, dtype=np.float64)
        if len(params) != self.n_toys:
            raise ValueError("need one parameter point per toy: got %d for %d toys" % (len(params), self.n_toys))
        lo, hi = self.bounds[rank]
        counts = [b - a for a, b in self.bounds]
        engine = getattr(self.ll, '_toy_engine', None)
        device_gather = (_dist().get_backend(self.group) == 'nccl' and engine is not None and min(counts) > 0
                         and hasattr(engine, 'peer_gather'))
        if not device_gather:
            local = self.ll.batch_toys(params[lo:hi], names, livetime_days=livetime_days) if hi > lo else np.zeros(0)
            return gather_concat(np.asarray(local, dtype=np.float64), counts, self.group)
        # the logl rows of all ranks (without priors) arrive over NVLink before the D2H (PeerGather)
        if self._gather is None or self._gather.n != max(counts):
            self._gather = PeerGather(max(counts), self.group)
        if self._gather.fallback is not None and min(counts) != max(counts):
            local = self.ll.batch_toys(params[lo:hi], names, livetime_days=livetime_days)
            return gather_concat(np.asarray(local, dtype=np.float64), counts, self.group)
        engine.peer_gather, engine.peer_mode = self._gather, 'gather'
        try:
            self.ll.batch_toys(params[lo:hi], names, livetime_days=livetime_days)
            gathered = engine.last_gathered
        finally:
            engine.peer_gather = None
        # (gathered is a fresh copy of the pinned landing buffer: equal shards need no second copy)
        device_ll = gathered.reshape(-1) if min(counts) == gathered.shape[1] else \
            np.concatenate([gathered[r, :c] for r, c in enumerate(counts)])
        has_priors = any(p is not None for _, p, _ in self.ll.shape_parameters.values()) or \
            any(p is not None for p in self.ll.rate_parameters.values())
        if not has_priors:
            return device_ll
        zs, mult = self.ll._rows_from_params(params, names)
        return np.where(np.isneginf(device_ll), -np.inf, self.ll._prior_sum(zs, mult) + device_ll)


def shard_events(d, group=None, rank=None, world_size=None):
    """This rank's contiguous slice of dataset d, aligned to the canonical 512-event superblock."""
    if rank is None:
        dist = _dist()
        rank, world_size = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(len(d), world_size, align=_cabi.SUPERBLOCK)[rank]
    return d[lo:hi]


class EventShardedLikelihood(object):
    """Unbinned likelihood whose EVENTS are sharded over the ranks of `group`.

    `ll` is a prepared UnbinnedLogLikelihood on which set_data(shard_events(d)) was called with this
    rank's slice.  Each evaluation does one all_gather of P doubles per rank."""

    def __init__(self, ll, group=None):
        self.ll = ll
        self.group = group
        self._gathers = {}

    def combine(self, logsum, musum, status, priors):
        """-musum + (rank-ordered sum of the shards' log sums) + priors; status != 0 -> -inf."""
        total = rank_ordered_sum(np.where(status != 0, 0.0, logsum), self.group)
        return np.where(status != 0, -np.inf, priors + (-musum + total))

    def batch(self, params, names=None, livetime_days=None):
        engine = getattr(self.ll, '_engine', None)
        n_points = len(np.asarray(params, dtype=np.float64).reshape(-1, max(len(names or self.ll.parameter_names()), 1)))
        device_gather = (_dist().get_backend(self.group) == 'nccl' and engine is not None
                         and hasattr(engine, 'peer_gather') and n_points > 0
                         and (not hasattr(engine, 'uses_mma') or engine.uses_mma()))
        if not device_gather:
            logsum, musum, status, priors = self.ll.batch_parts(params, names, livetime_days=livetime_days)
            return self.combine(logsum, musum, status, priors)
        # ONE launch after the evaluation (bi_peer_exchange, inside the evaluation's CUDA graph): the shards' log sums go
        # to every rank over NVLink peer memory and are added up in rank order on the device, so all ranks hold
        # bit-identical results; -sum(mu) is added there too, the priors (Python callables) here
        pg = self._gathers.get(n_points)
        if pg is None:
            pg = self._gathers[n_points] = PeerGather(n_points, self.group)
        engine.peer_gather, engine.peer_mode = pg, 'sum'
        try:
            logsum, musum, status, priors = self.ll.batch_parts(params, names, livetime_days=livetime_days)
            total = engine.last_total
        finally:
            engine.peer_gather = None
        return np.where(status != 0, -np.inf, priors + total)

    def __call__(self, livetime_days=None, **kwargs):
        names = list(kwargs.keys())
        row = np.array([[kwargs[n] for n in names]], dtype=np.float64).reshape(1, len(names))
        return self.batch(row, names, livetime_days=livetime_days)[0]
