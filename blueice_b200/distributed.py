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


class PointShardedLikelihood(object):
    """Shard the points of ll.batch over the ranks of `group`; every rank returns the full result.

    `ll` is a prepared likelihood with the SAME data set on every rank (replicated templates / anchor
    tensor).  Any object with a `.batch(params, names, livetime_days=None)` method works."""

    def __init__(self, ll, group=None):
        self.ll = ll
        self.group = group

    def batch(self, params, names=None, livetime_days=None):
        dist = _dist()
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        params = np.asarray(params, dtype=np.float64)
        bounds = shard_bounds(len(params), world)
        lo, hi = bounds[rank]
        local = self.ll.batch(params[lo:hi], names, livetime_days=livetime_days) if hi > lo else np.zeros(0)
        return gather_concat(np.asarray(local, dtype=np.float64), [b - a for a, b in bounds], self.group)


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
        local = self.ll.batch_toys(params[lo:hi], names, livetime_days=livetime_days) if hi > lo else np.zeros(0)
        return gather_concat(np.asarray(local, dtype=np.float64), [b - a for a, b in self.bounds], self.group)


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

    def combine(self, logsum, musum, status, priors):
        """-musum + (rank-ordered sum of the shards' log sums) + priors; status != 0 -> -inf."""
        total = rank_ordered_sum(np.where(status != 0, 0.0, logsum), self.group)
        return np.where(status != 0, -np.inf, priors + (-musum + total))

    def batch(self, params, names=None, livetime_days=None):
        logsum, musum, status, priors = self.ll.batch_parts(params, names, livetime_days=livetime_days)
        return self.combine(logsum, musum, status, priors)

    def __call__(self, livetime_days=None, **kwargs):
        names = list(kwargs.keys())
        row = np.array([[kwargs[n] for n in names]], dtype=np.float64).reshape(1, len(names))
        return self.batch(row, names, livetime_days=livetime_days)[0]
