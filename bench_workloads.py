"""Seeded synthetic workloads of BASELINE.json's configs.

Array-level generators (pure NumPy; shared by the product benchmark, the oracle baseline and the
tests, so both arms see byte-identical inputs) plus builders that express the same workload through
the blueice_b200 API.  Nothing here touches /root/reference.

  config 1  conf_for_test() Gaussian source, shape parameter mu (3 anchors), ~1k events
  config 2  2 sources, 2-D 100x100 histogram templates, 2 shape nuisances x 5 anchors, ~100k events,
            4096-point profile scan                                          <- the bench headline
  config 3  binned + Beeston-Barlow, 3-D bins, 4 sources, 3 shape parameters x 3 anchors
"""
import numpy as np

ANCHORS_5 = (-2., -1., 0., 1., 2.)


# ------------------------------------------------------------------------------------------------
# config 2: 2-D blob templates
# ------------------------------------------------------------------------------------------------
C2_SPACE = (('cs1', np.linspace(0, 100, 101)), ('cs2', np.linspace(0, 4, 101)))
C2_SOURCES = (
    dict(name='bg', centre=(40., 2.5), widths=(15., 0.5), events_per_day=90000.),
    dict(name='sig', centre=(30., 1.5), widths=(15., 0.5), events_per_day=10000.),
    dict(name='src2', centre=(55., 1.0), widths=(10., 0.4), events_per_day=5000.),
    dict(name='src3', centre=(70., 3.0), widths=(12., 0.3), events_per_day=3000.),
    dict(name='src4', centre=(20., 3.2), widths=(8., 0.6), events_per_day=2000.),
    dict(name='src5', centre=(85., 2.0), widths=(9., 0.7), events_per_day=1000.),
)


def blob_density(edges, centre, widths, shifts=(), floor=1e-6):
    """Normalised 2-D Gaussian blob + floor as a density histogram over `edges`.

    shifts[0] moves the centre, shifts[1] widens the blob by (1 + 0.1 * shift), further shifts tilt it."""
    s = list(shifts) + [0.] * (4 - len(shifts))
    cx = centre[0] + 2.0 * s[0] + 1.0 * s[2]
    cy = centre[1] + 0.05 * s[0] - 0.04 * s[3]
    wx = widths[0] * (1 + 0.1 * s[1]) * (1 + 0.05 * s[3])
    wy = widths[1] * (1 + 0.1 * s[1]) * (1 - 0.05 * s[2])
    x = 0.5 * (edges[0][1:] + edges[0][:-1])
    y = 0.5 * (edges[1][1:] + edges[1][:-1])
    gx = np.exp(-0.5 * ((x - cx) / wx) ** 2)
    gy = np.exp(-0.5 * ((y - cy) / wy) ** 2)
    h = np.outer(gx, gy) + floor
    vol = np.outer(np.diff(edges[0]), np.diff(edges[1]))
    return h / np.sum(h * vol)


def c2_arrays(n_sources=2, n_shape=2, anchors=ANCHORS_5, bins=(100, 100)):
    """(axes, edges, templates [n]*n_shape + [S, bx, by], mus_anchor [n]*n_shape + [S])."""
    edges = [np.linspace(0, 100, bins[0] + 1), np.linspace(0, 4, bins[1] + 1)]
    axes = [np.asarray(anchors, dtype=float) for _ in range(n_shape)]
    shape = [len(a) for a in axes]
    templates = np.zeros(shape + [n_sources, bins[0], bins[1]])
    mus = np.zeros(shape + [n_sources])
    for idx in np.ndindex(*shape):
        zs = [axes[d][i] for d, i in enumerate(idx)]
        for s in range(n_sources):
            spec = C2_SOURCES[s]
            # sources react to the nuisances with different strength so that every anchor differs
            scale = 1.0 if s % 2 == 0 else 0.6
            templates[idx + (s,)] = blob_density(edges, spec['centre'], spec['widths'], [scale * z for z in zs])
            mus[idx + (s,)] = spec['events_per_day'] * (1 + 0.02 * zs[0] * (1 if s == 0 else -1))
    return axes, edges, templates, mus


def sample_from_density(density, edges, n, rng):
    """n events from a 2-D density histogram (uniform within bins)."""
    vol = np.outer(np.diff(edges[0]), np.diff(edges[1]))
    p = (density * vol).ravel()
    p = p / p.sum()
    flat = rng.choice(len(p), size=n, p=p)
    ix, iy = np.unravel_index(flat, density.shape)
    x = edges[0][ix] + rng.random(n) * (edges[0][ix + 1] - edges[0][ix])
    y = edges[1][iy] + rng.random(n) * (edges[1][iy + 1] - edges[1][iy])
    return x, y


def c2_events(templates, mus, edges, n_events=None, seed=1):
    """Events from the base-model mixture (centre anchor); Poisson total unless n_events is given."""
    rng = np.random.default_rng(seed)
    d = templates.ndim - 3
    centre = tuple(s // 2 for s in templates.shape[:d])
    base_mus = mus[centre]
    xs, ys = [], []
    for s, mu in enumerate(base_mus):
        n = rng.poisson(mu) if n_events is None else int(round(n_events * mu / base_mus.sum()))
        x, y = sample_from_density(templates[centre + (s,)], edges, n, rng)
        xs.append(x)
        ys.append(y)
    x, y = np.concatenate(xs), np.concatenate(ys)
    perm = rng.permutation(len(x))
    return x[perm], y[perm]


def scan_points(n_points, n_shape, n_sources, seed=2, z_range=(-2., 2.), mult_range=(0.5, 1.5)):
    """Profile-scan points: (zs [P, n_shape], rate multipliers [P, n_sources]), uniform in the box."""
    rng = np.random.default_rng(seed)
    zs = rng.uniform(z_range[0], z_range[1], size=(n_points, n_shape))
    mult = rng.uniform(mult_range[0], mult_range[1], size=(n_points, n_sources))
    return zs, mult


def c2_api(n_sources=2, n_shape=2, anchors=ANCHORS_5, bins=(100, 100), n_events=None, seed=1,
           method='linear', likelihood_config=None):
    """The config-2 workload through the public API.  Returns (ll, data, names) with data set."""
    from blueice_b200 import HistogramPdfSource, UnbinnedLogLikelihood
    from blueice_b200.hist import Histdd

    edges = [np.linspace(0, 100, bins[0] + 1), np.linspace(0, 4, bins[1] + 1)]
    shift_names = ['shift%d' % (i + 1) for i in range(n_shape)]

    class BlobSource(HistogramPdfSource):
        """Histogram template filled analytically (no Monte Carlo) from the source's config."""

        def build_histogram(self):
            c = self.config
            names, space_bins = zip(*c['analysis_space'])
            scale = c.get('shift_scale', 1.0)
            shifts = [scale * c[n] for n in shift_names]
            dens = blob_density([np.asarray(b, dtype=float) for b in space_bins], c['centre'], c['widths'], shifts)
            self._pdf_histogram = Histdd.from_histogram(dens, space_bins, axis_names=names)
            self._bin_volumes = self._pdf_histogram.bin_volumes()
            self._n_events_histogram = Histdd.from_histogram(np.full(dens.shape, np.inf), space_bins, names)
            self.events_per_day = c['base_rate'] * (1 + 0.02 * c[shift_names[0]] * c.get('rate_sign', 1))

    sources = []
    for s in range(n_sources):
        spec = C2_SOURCES[s]
        sources.append(dict(name=spec['name'], centre=spec['centre'], widths=spec['widths'],
                            base_rate=spec['events_per_day'], shift_scale=1.0 if s % 2 == 0 else 0.6,
                            rate_sign=1 if s == 0 else -1))
    config = dict(sources=sources, default_source_class=BlobSource,
                  analysis_space=[['cs1', edges[0]], ['cs2', edges[1]]],
                  pdf_interpolation_method=method, livetime_days=1.,
                  force_recalculation=True, never_save_to_cache=True)
    for n in shift_names:
        config[n] = 0.
    ll = UnbinnedLogLikelihood(config, likelihood_config)
    for spec in sources:
        ll.add_rate_parameter(spec['name'])
    for n in shift_names:
        ll.add_shape_parameter(n, anchors)
    ll.prepare()
    _, _, templates, mus = c2_arrays(n_sources, n_shape, anchors, bins)
    x, y = c2_events(templates, mus, edges, n_events, seed)
    d = np.zeros(len(x), dtype=[('cs1', float), ('cs2', float), ('source', int)])
    d['cs1'], d['cs2'] = x, y
    ll.set_data(d)
    names = [spec['name'] + '_rate_multiplier' for spec in sources] + shift_names
    return ll, d, names


# ------------------------------------------------------------------------------------------------
# config 1: Gaussian source, one shape parameter with 3 anchors
# ------------------------------------------------------------------------------------------------
def c1_arrays(seed=0, n_expected=1000., mu_anchors=(-2., 0., 2.)):
    """(axes, mus_anchor [3, 1], ps_anchor [3, 1, N], x [N]) of the config-1 model at array level."""
    from scipy import stats
    rng = np.random.default_rng(seed)
    n = rng.poisson(n_expected)
    x = rng.normal(0., 1., n)
    x = x[(x >= -10) & (x <= 10)]
    axes = [np.asarray(mu_anchors, dtype=float)]
    ps = np.stack([stats.norm(m, 1.).pdf(x)[np.newaxis, :] for m in mu_anchors])
    mus = np.full((len(mu_anchors), 1), n_expected)
    return axes, mus, ps, x


def c1_api(seed=0):
    """Config 1 through the public API (form of the reference's tests/test_likelihood.py:102)."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import conf_for_test
    ll = UnbinnedLogLikelihood(conf_for_test(n_sources=1, force_recalculation=True, never_save_to_cache=True))
    ll.add_rate_parameter('s0')
    ll.add_shape_parameter('mu', {-2: -2, 0: 0, 2: 2})
    ll.prepare()
    np.random.seed(seed)
    d = ll.base_model.simulate()
    ll.set_data(d)
    return ll, d, ['s0_rate_multiplier', 'mu']


# ------------------------------------------------------------------------------------------------
# config 3: binned, Beeston-Barlow, 3-D bins
# ------------------------------------------------------------------------------------------------
def c3_arrays(bins=(200, 200, 20), n_sources=4, n_shape=3, anchors=(-1., 0., 1.), seed=3, total_events=2.0e5):
    """(axes, mus_anchor, pmf_anchor [3,3,3,S,*bins], n_model_anchor (same shape), observed [*bins])."""
    rng = np.random.default_rng(seed)
    ex = np.linspace(-6, 6, bins[0] + 1)
    ey = np.linspace(-6, 6, bins[1] + 1)
    ez = np.linspace(0, 1, bins[2] + 1)
    cx, cy, cz = [0.5 * (e[1:] + e[:-1]) for e in (ex, ey, ez)]
    axes = [np.asarray(anchors, dtype=float) for _ in range(n_shape)]
    shape = [len(a) for a in axes]
    n_bins = int(np.prod(bins))
    pmf = np.zeros(shape + [n_sources] + list(bins))
    mus = np.zeros(shape + [n_sources])
    centres = [(-1.5, -1.0, 0.3), (1.0, 1.5, 0.6), (0.0, 0.0, 0.5), (2.0, -2.0, 0.8), (-2.5, 2.0, 0.2), (0.5, 2.5, 0.4)]
    for idx in np.ndindex(*shape):
        zs = [axes[d][i] for d, i in enumerate(idx)] + [0., 0., 0.]
        for s in range(n_sources):
            c = centres[s]
            k = 1.0 if s % 2 == 0 else 0.5
            gx = np.exp(-0.5 * ((cx - c[0] - 0.3 * k * zs[0]) / (1.2 * (1 + 0.1 * k * zs[1]))) ** 2) + 1e-4
            gy = np.exp(-0.5 * ((cy - c[1] + 0.2 * k * zs[0]) / (1.0 * (1 + 0.1 * k * zs[1]))) ** 2) + 1e-4
            gz = np.exp(-0.5 * ((cz - c[2] - 0.05 * k * zs[2]) / 0.25) ** 2) + 1e-3
            g = gx[:, None, None] * gy[None, :, None] * gz[None, None, :]
            pmf[idx + (s,)] = g / g.sum()
            mus[idx + (s,)] = total_events * (0.55, 0.25, 0.15, 0.05, 0.03, 0.02)[s] * (1 + 0.03 * zs[0] * (1 - s))
    # calibration events per bin: every bin >= 1, else the reference asserts
    n_model = 1.0 + rng.poisson(50.0 * pmf * n_bins / 10.0).astype(float)
    centre = tuple(s // 2 for s in shape)
    lam = np.tensordot(mus[centre], pmf[centre], axes=(0, 0))
    observed = rng.poisson(lam).astype(float)
    return axes, [ex, ey, ez], mus, pmf, n_model, observed


# ------------------------------------------------------------------------------------------------
# Source classes defined from arrays, usable with EITHER package (the reference, in
# tests/golden/make_golden.py, or blueice_b200 in the tests): the base class and the histogram class
# are passed in, so both sides build byte-identical models.
# ------------------------------------------------------------------------------------------------
def array_source_class(base_class, histdd_class, axes, edges, names, mus, density, n_model, param_names):
    """HistogramPdfSource subclass whose histograms are looked up in dense anchor arrays.

    density / n_model: [n1..nD, S, *bins];  mus: [n1..nD, S];  config keys: param_names + 'source_index'."""
    axes = [np.asarray(a, dtype=float) for a in axes]

    class ArraySource(base_class):
        def build_histogram(self):
            c = self.config
            idx = tuple(int(np.argmin(np.abs(axes[d] - c[param_names[d]]))) for d in range(len(axes)))
            s = int(c['source_index'])
            dens = histdd_class(bins=edges, axis_names=names)
            dens.histogram = np.array(density[idx + (s,)], dtype=float)
            counts = histdd_class(bins=edges, axis_names=names)
            counts.histogram = (np.array(n_model[idx + (s,)], dtype=float) if n_model is not None
                                else np.full(dens.histogram.shape, np.inf))
            vol = np.ones(1)
            for e in edges:
                vol = np.multiply.outer(vol, np.diff(np.asarray(e, dtype=float)))
            self._bin_volumes = vol.reshape(dens.histogram.shape)
            self._pdf_histogram = dens
            self._n_events_histogram = counts
            self.events_per_day = float(mus[idx + (s,)])

    return ArraySource


def array_model_config(source_class, edges, names, n_sources, param_names, method='linear'):
    config = dict(sources=[dict(name='src%d' % s, source_index=s) for s in range(n_sources)],
                  default_source_class=source_class,
                  analysis_space=[[n, np.asarray(e, dtype=float)] for n, e in zip(names, edges)],
                  pdf_interpolation_method=method, livetime_days=1.,
                  force_recalculation=True, never_save_to_cache=True)
    for p in param_names:
        config[p] = 0.
    return config


def events_from_counts(edges, counts, seed=0):
    """One event at a random position inside its bin for every observed count (record-array columns)."""
    rng = np.random.default_rng(seed)
    idx = np.nonzero(counts)
    reps = counts[idx].astype(int)
    cols = []
    for d, e in enumerate(edges):
        e = np.asarray(e, dtype=float)
        lo = np.repeat(e[idx[d]], reps)
        hi = np.repeat(e[idx[d] + 1], reps)
        cols.append(lo + (0.25 + 0.5 * rng.random(len(lo))) * (hi - lo))
    return cols
