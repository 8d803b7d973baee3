"""On-device toy Monte Carlo datasets: Model.simulate for MANY toys at once.

`simulate_toys(model, n_toys, ...)` does what `[model.simulate(...) for _ in range(n_toys)]` does in the
reference (model.py:69-91 -> source.py:248-264 -> multihist Histdd.get_random), on the device, for models whose
sources are stock histogram templates: a Poisson number of events per source and toy, each event a template bin
drawn from the source's pmf and a uniform position inside it.  The events stay in HBM as a `ToyData` that
`UnbinnedLogLikelihood.set_toy_data` consumes directly.  Counter-based random numbers (Philox, keyed by seed and
GLOBAL toy id): toys [a, b) are the same events whether generated in one call, in pieces, or on different GPUs.
Parity with the reference is distributional; the event stage is restated bit for bit in oracle/toys.py."""
import numpy as np

from . import _cabi
from .source import HistogramPdfSource


class ToyData(object):
    """T toy datasets on the device: coords [n_space, N] (torch float64), source [N] (torch int32),
    offsets [T + 1] (NumPy int64: first event of every toy), counts [T, S] (NumPy int32)."""

    def __init__(self, dims, coords, source, offsets, counts, first_toy, seed):
        self.dims, self.coords, self.source = list(dims), coords, source
        self.offsets, self.counts, self.first_toy, self.seed = offsets, counts, int(first_toy), int(seed)

    def __len__(self):
        return len(self.offsets) - 1

    @property
    def n_events(self):
        return int(self.offsets[-1])

    def to_records(self, toy=None):
        """Events as the record array Model.simulate returns ('source' + one field per analysis dimension):
        of one toy, or of all toys back to back (use with .offsets)."""
        lo, hi = (0, self.n_events) if toy is None else (int(self.offsets[toy]), int(self.offsets[toy + 1]))
        d = np.zeros(hi - lo, dtype=[('source', int)] + [(name, float) for name in self.dims])
        if hi > lo:
            host = self.coords[:, lo:hi].cpu().numpy()
            for k, name in enumerate(self.dims):
                d[name] = host[k]
            d['source'] = self.source[lo:hi].cpu().numpy()
        return d


def source_tables(model):
    """(edges list, cdf [S, B]) of a model whose sources are stock histogram templates on shared bin edges;
    cdf rows are Histdd.get_random's table cumsum(pmf.ravel()) / sum."""
    edges0, rows = None, []
    for source in model.sources:
        if not (isinstance(source, HistogramPdfSource) and type(source).simulate is HistogramPdfSource.simulate):
            raise NotImplementedError("simulate_toys needs stock HistogramPdfSource sources; %r is not" % (source.name,))
        _, edges, _ = source.template()
        edges = [np.ascontiguousarray(np.asarray(e, dtype=np.float64)) for e in edges]
        if edges0 is None:
            edges0 = edges
        elif len(edges) != len(edges0) or any(not np.array_equal(a, b) for a, b in zip(edges, edges0)):
            raise NotImplementedError("simulate_toys needs all sources on the same bin edges")
        flat = np.asarray(source.get_pmf_grid()[0], dtype=np.float64).ravel()
        cdf = np.cumsum(flat)
        rows.append(cdf / cdf[-1])
    return edges0, np.ascontiguousarray(np.vstack(rows))


def toy_means(model, rate_multipliers=None, livetime_days=None):
    """Poisson means per source, exactly as Model.simulate forms them (model.py:80-84)."""
    rate_multipliers = rate_multipliers or {}
    mus = []
    for source in model.sources:
        mu = model.expected_events(source) * rate_multipliers.get(source.name, 1) / source.fraction_in_range
        if livetime_days is not None:
            mu *= livetime_days / model.config['livetime_days']
        mus.append(mu)
    return np.asarray(mus, dtype=np.float64)


def generate(edges, cdf, mus, n_toys, seed=0, first_toy=0, device=None):
    """Array-level generator.  mus: [S] (all toys) or [n_toys, S].  Returns (coords, source, offsets, counts)."""
    from .engine import require_cuda
    import ctypes
    torch = require_cuda()
    lib = _cabi.load()
    device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    mus = np.ascontiguousarray(np.asarray(mus, dtype=np.float64))
    S = cdf.shape[0]
    per_toy = mus.ndim == 2
    if mus.shape != ((n_toys, S) if per_toy else (S,)):
        raise ValueError("mus must have shape [n_sources] or [n_toys, n_sources]")
    if np.any(mus >= 2.0 ** 30):
        raise ValueError("toy means must stay below 2^30 events per source")
    n_space = len(edges)
    n_bins = _cabi.as_i32([len(e) - 1 for e in edges])
    edges_concat = _cabi.as_f64(np.concatenate(edges))
    mus_d = torch.from_numpy(mus).to(device)
    counts_d = torch.empty((max(n_toys, 1), S), dtype=torch.int32, device=device)
    _cabi.check(lib.bi_toy_counts(S, n_toys, first_toy, _cabi.dev_ptr(mus_d), int(per_toy), seed,
                                  _cabi.dev_ptr(counts_d), stream), "bi_toy_counts")
    counts = counts_d[:n_toys].cpu().numpy()
    offsets = np.zeros(n_toys + 1, dtype=np.int64)
    np.cumsum(counts.sum(axis=1, dtype=np.int64), out=offsets[1:])
    n = int(offsets[-1])
    coords = torch.empty((n_space, max(n, 1)), dtype=torch.float64, device=device)
    source = torch.empty(max(n, 1), dtype=torch.int32, device=device)
    if n:
        cdf_d = torch.from_numpy(np.ascontiguousarray(cdf)).to(device)
        offsets_d = torch.from_numpy(offsets).to(device)
        _cabi.check(lib.bi_toy_events(n_space, _cabi.host_ptr(n_bins), _cabi.host_ptr(edges_concat), S,
                                      _cabi.dev_ptr(cdf_d), n_toys, first_toy, _cabi.dev_ptr(counts_d),
                                      _cabi.dev_ptr(offsets_d), n, seed, _cabi.dev_ptr(coords), coords.shape[1],
                                      _cabi.dev_ptr(source), stream), "bi_toy_events")
        torch.cuda.current_stream(device).synchronize()
    return coords[:, :n], source[:n], offsets, counts


def simulate_toys(model, n_toys, rate_multipliers=None, livetime_days=None, seed=0, first_toy=0, mus=None,
                  device=None):
    """n_toys toy datasets of `model` (ids first_toy .. first_toy + n_toys - 1) as a ToyData.

    rate_multipliers / livetime_days as in Model.simulate; `mus` ([S] or [n_toys, S]) overrides the Poisson means
    (e.g. one hypothesis per toy)."""
    edges, cdf = source_tables(model)
    if mus is None:
        mus = toy_means(model, rate_multipliers, livetime_days)
    coords, source, offsets, counts = generate(edges, cdf, mus, int(n_toys), seed, first_toy, device)
    dims = [dim[0] for dim in model.config['analysis_space']]
    return ToyData(dims, coords, source, offsets, counts, first_toy, seed)
