"""Generate golden vectors by running the UNMODIFIED reference (JelleAalbers/blueice v1.2.1).

Run in the build container only (the GPU box has no /root/reference):

    cd /tmp/somewhere && python /root/repo/tests/golden/make_golden.py

The reference needs `multihist` and `atomicwrites`, which are not installable here; the two
test-only stand-ins under tests/golden/_shims are put on sys.path for this script alone
(the reference's own 30 tests pass with them).  Outputs: tests/golden/*.npz,
small enough to commit; tests/test_golden_*.py replay them through blueice_b200 and the oracle.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, '_shims'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, REPO)

os.chdir(tempfile.mkdtemp(prefix='blueice_golden_'))     # the reference writes ./pdf_cache

import blueice                                            # noqa: E402  (the reference)
from blueice.likelihood import BinnedLogLikelihood, UnbinnedLogLikelihood   # noqa: E402
from blueice.source import HistogramPdfSource             # noqa: E402
from blueice.test_helpers import conf_for_test            # noqa: E402
from multihist import Histdd                              # noqa: E402  (shim)

import bench_workloads as wl                              # noqa: E402

assert blueice.__version__ == '1.2.1'


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('wrote', path, {k: np.asarray(v).shape for k, v in arrays.items()})


def special_points(zs, mult, axes):
    """Append points on anchors, on the grid boundary, out of range, and with odd rate multipliers."""
    zs, mult = zs.copy(), mult.copy()
    k = 0
    for d, a in enumerate(axes):
        for v in (a[0], a[len(a) // 2], a[-1]):
            zs[k, d] = v
            k += 1
    zs[k, 0] = axes[0][-1] + 0.5; k += 1          # out of range -> -inf
    zs[k, 0] = np.nan; k += 1                     # NaN -> -inf
    mult[k, 0] = 0.0; k += 1                      # zero rate
    mult[k, 0] = -1.0; k += 1                     # unphysical -> -inf
    mult[k, -1] = np.inf; k += 1                  # unphysical -> -inf
    return zs, mult


# ---------------------------------------------------------------------------------------------
# G1: config 1 (Gaussian source, analytic pdf, mu with 3 anchors)
# ---------------------------------------------------------------------------------------------
def golden_c1():
    lf = UnbinnedLogLikelihood(conf_for_test(n_sources=1))
    lf.add_rate_parameter('s0')
    lf.add_shape_parameter('mu', {-2: -2, 0: 0, 2: 2})
    lf.prepare()
    np.random.seed(0)
    d = lf.base_model.simulate()
    lf.set_data(d)
    zs, mult = wl.scan_points(64, 1, 1, seed=1, mult_range=(0.5, 2.0))
    zs, mult = special_points(zs, mult, [np.array([-2., 0., 2.])])
    logl = np.array([lf(mu=float(z[0]), s0_rate_multiplier=float(m[0])) for z, m in zip(zs, mult)])
    full = [lf(mu=float(zs[i, 0]), s0_rate_multiplier=float(mult[i, 0]), full_output=True) for i in (20, 21)]
    save('c1_gaussian', x=d['x'], zs=zs, mult=mult, logl=logl,
         full_index=np.array([20, 21]), full_mus=np.array([f[1] for f in full]),
         full_ps=np.array([f[2] for f in full]))


# ---------------------------------------------------------------------------------------------
# G2: config-2-shaped (histogram templates, 2 shape parameters), linear and piecewise lookup
# ---------------------------------------------------------------------------------------------
def golden_c2(method):
    n_sources, n_shape, anchors, bins = 2, 2, (-1., 0., 1.), (20, 16)
    axes, edges, templates, mus = wl.c2_arrays(n_sources, n_shape, anchors, bins)
    names = ['cs1', 'cs2']
    params = ['shift1', 'shift2']
    cls = wl.array_source_class(HistogramPdfSource, Histdd, axes, edges, names, mus, templates, None, params)
    lf = UnbinnedLogLikelihood(wl.array_model_config(cls, edges, names, n_sources, params, method))
    for s in range(n_sources):
        lf.add_rate_parameter('src%d' % s)
    for p in params:
        lf.add_shape_parameter(p, anchors)
    lf.prepare()
    x, y = wl.c2_events(templates, mus * 4e-3, edges, seed=7)          # ~400 events
    # put some events exactly on bin edges / bin centres / the space boundary
    x[:4] = [0., 100., 50., 52.5]
    y[:4] = [0., 4., 2., 2.125]
    d = np.zeros(len(x), dtype=[('cs1', float), ('cs2', float), ('source', int)])
    d['cs1'], d['cs2'] = x, y
    lf.set_data(d)
    zs, mult = wl.scan_points(48, n_shape, n_sources, seed=2, z_range=(-1., 1.))
    zs, mult = special_points(zs, mult, axes)
    logl = np.array([lf(shift1=float(z[0]), shift2=float(z[1]), src0_rate_multiplier=float(m[0]),
                        src1_rate_multiplier=float(m[1])) for z, m in zip(zs, mult)])
    full = [lf(shift1=float(zs[i, 0]), shift2=float(zs[i, 1]), src0_rate_multiplier=float(mult[i, 0]),
               src1_rate_multiplier=float(mult[i, 1]), full_output=True) for i in (30, 31)]
    # the dense anchor tensor of the reference itself (pdf_morphers.py:59-65)
    ps_anchor = np.array([[lf.anchor_models[(a, b)].score_events(d) for b in anchors] for a in anchors])
    save('c2_templates_' + method, x=x, y=y, zs=zs, mult=mult, logl=logl, ps_anchor=ps_anchor,
         full_index=np.array([30, 31]), full_mus=np.array([f[1] for f in full]),
         full_ps=np.array([f[2] for f in full]))


# ---------------------------------------------------------------------------------------------
# G3: binned likelihood, with and without Beeston-Barlow, 3-D bins, 3 sources, 2 shape parameters
# ---------------------------------------------------------------------------------------------
def golden_binned(bb):
    bins, n_sources, n_shape, anchors = (6, 5, 4), 3, 2, (-1., 0., 1.)
    axes, edges, mus, pmf, n_model, observed = wl.c3_arrays(bins, n_sources, n_shape, anchors, seed=3,
                                                            total_events=600.)
    vol = np.ones(1)
    for e in edges:
        vol = np.multiply.outer(vol, np.diff(e))
    density = pmf / vol.reshape(bins)
    names = ['x', 'y', 'z']
    params = ['za', 'zb']
    cls = wl.array_source_class(HistogramPdfSource, Histdd, axes, edges, names, mus, density, n_model, params)
    likelihood_config = ({'model_statistical_uncertainty_handling': 'bb_single', 'bb_single_source': 0}
                         if bb else None)
    lf = BinnedLogLikelihood(wl.array_model_config(cls, edges, names, n_sources, params), likelihood_config)
    for s in range(n_sources):
        lf.add_rate_parameter('src%d' % s)
    for p in params:
        lf.add_shape_parameter(p, anchors)
    lf.prepare()
    cols = wl.events_from_counts(edges, observed, seed=5)
    d = np.zeros(len(cols[0]), dtype=[('x', float), ('y', float), ('z', float), ('source', int)])
    d['x'], d['y'], d['z'] = cols
    lf.set_data(d)
    assert np.array_equal(lf.data_events_per_bin.histogram, observed)
    zs, mult = wl.scan_points(40, n_shape, n_sources, seed=4, z_range=(-1., 1.), mult_range=(0.6, 1.4))
    zs, mult = special_points(zs, mult, axes)
    logl, raises = [], []
    for z, m in zip(zs, mult):
        try:
            logl.append(lf(za=float(z[0]), zb=float(z[1]), **{'src%d_rate_multiplier' % s: float(m[s])
                                                            for s in range(n_sources)}))
            raises.append(False)
        except AssertionError:        # Beeston-Barlow asserts (likelihood.py:649,655), e.g. zero rate -> 0/0
            logl.append(np.nan)
            raises.append(True)
    full = [lf(za=float(zs[i, 0]), zb=float(zs[i, 1]), full_output=True,
               **{'src%d_rate_multiplier' % s: float(mult[i, s]) for s in range(n_sources)}) for i in (30, 31)]
    save('binned_bb' if bb else 'binned_plain', x=d['x'], y=d['y'], z=d['z'], observed=observed,
         zs=zs, mult=mult, logl=np.array(logl), raises=np.array(raises), full_index=np.array([30, 31]),
         full_mus=np.array([f[1] for f in full]), full_pmfs=np.array([f[2] for f in full]))


# ---------------------------------------------------------------------------------------------
# G4: values of the reference's own test scenarios on seeded data (tests/test_likelihood.py:124-148)
# ---------------------------------------------------------------------------------------------
def golden_multisource():
    lf = UnbinnedLogLikelihood(conf_for_test(n_sources=2))
    lf.add_shape_parameter('some_multiplier', (0.5, 1, 2, 4))
    lf.add_rate_parameter('s0')
    lf.add_rate_parameter('s1')
    lf.prepare()
    np.random.seed(42)
    d = lf.base_model.simulate()
    lf.set_data(d)
    calls = [dict(), dict(s0_rate_multiplier=2.), dict(s1_rate_multiplier=2.), dict(s0_rate_multiplier=4.),
             dict(s0_rate_multiplier=2.5, s1_rate_multiplier=2.5), dict(s0_rate_multiplier=2., s1_rate_multiplier=2.),
             dict(some_multiplier=2.), dict(some_multiplier=3.3, s0_rate_multiplier=0.7), dict(some_multiplier=0.5),
             dict(some_multiplier=4.), dict(some_multiplier=0.75, s1_rate_multiplier=1.2)]
    names = ['s0_rate_multiplier', 's1_rate_multiplier', 'some_multiplier']
    table = np.array([[c.get(n, 1.) for n in names] for c in calls])
    logl = np.array([lf(**c) for c in calls])
    save('multisource', x=d['x'], params=table, logl=logl)


# ---------------------------------------------------------------------------------------------
# G5: source-wise interpolation (likelihood.py:113-145,210-240,534-555): s0 depends on (mu, sigma),
# s1 on sigma only, s2 on no shape parameter
# ---------------------------------------------------------------------------------------------
def sourcewise_config():
    config = conf_for_test(n_sources=3)
    config['sources'][0]['events_per_day'] = 700.
    config['sources'][1].update(events_per_day=250., extra_dont_hash_settings=['mu'])
    config['sources'][2].update(events_per_day=50., extra_dont_hash_settings=['mu', 'sigma'])
    config['source_wise_interpolation'] = True
    return config


def golden_sourcewise():
    lf = UnbinnedLogLikelihood(sourcewise_config())
    for s in range(3):
        lf.add_rate_parameter('s%d' % s)
    lf.add_shape_parameter('mu', {-2: -2, 0: 0, 2: 2})
    lf.add_shape_parameter('sigma', (0.5, 1, 2))
    lf.prepare()
    np.random.seed(3)
    d = lf.base_model.simulate()
    lf.set_data(d)
    axes = [np.array([-2., 0., 2.]), np.array([0.5, 1., 2.])]
    rng = np.random.default_rng(11)
    n = 40
    zs = np.column_stack([rng.uniform(-2, 2, n), rng.uniform(0.5, 2, n)])
    mult = rng.uniform(0.5, 1.5, (n, 3))
    zs, mult = special_points(zs, mult, axes)
    logl = np.array([lf(mu=float(z[0]), sigma=float(z[1]), **{'s%d_rate_multiplier' % s: float(m[s]) for s in range(3)})
                     for z, m in zip(zs, mult)])
    full = [lf(mu=float(zs[i, 0]), sigma=float(zs[i, 1]), full_output=True,
               **{'s%d_rate_multiplier' % s: float(mult[i, s]) for s in range(3)}) for i in (30, 31)]
    # the per-source anchor rows of the reference itself, sources concatenated, sub-anchors in C order
    rows, mus_rows = [], []
    for sn, base_source in zip(lf.source_name_list, lf.base_model.sources):
        if sn in lf.source_morphers:
            for anchor in lf.source_morphers[sn].get_anchor_points(bounds=None):
                src = lf.anchor_sources[sn][anchor]
                rows.append(src.pdf(d['x']))
                mus_rows.append(src.expected_events)
        else:
            rows.append(base_source.pdf(d['x']))
            mus_rows.append(base_source.expected_events)
    save('sourcewise', x=d['x'], zs=zs, mult=mult, logl=logl, rows=np.array(rows), mus_rows=np.array(mus_rows),
         full_index=np.array([30, 31]), full_mus=np.array([f[1] for f in full]),
         full_ps=np.array([f[2] for f in full]))


# ---------------------------------------------------------------------------------------------
# G6: LogLikelihoodReParam (likelihood.py:715-864) on the reference's own fixtures
# (test_helpers.py:70-97, tests/test_likelihood_reparam.py)
# ---------------------------------------------------------------------------------------------
def golden_reparam():
    from copy import deepcopy
    from blueice.likelihood import LogLikelihoodReParam
    from blueice.test_helpers import BASE_CONV_CONFIG, conf_for_reparam_test
    # events_per_day as a FLOAT: with the test's integer 1 the reference's mus array is an integer array and
    # `mus[s] *= mult` (likelihood.py:368) truncates non-integer products -- invisible in the reference's own test
    # (integer points only), and a quirk blueice_b200 does not reproduce
    lf_old = UnbinnedLogLikelihood(conf_for_reparam_test(events_per_day=1.))
    for name in ('op0', 'op1', 'op2'):
        lf_old.add_rate_parameter(name)
    lf_old.prepare()
    lf = LogLikelihoodReParam(lf_old, deepcopy(BASE_CONV_CONFIG))
    rng = np.random.default_rng(6)
    x = rng.normal(0., 1., 7)
    d = np.zeros(len(x), dtype=[('x', float), ('source', int)])
    d['x'] = x
    lf.set_data(d)
    pts = np.column_stack([rng.uniform(0.3, 4., 24), rng.uniform(0.3, 4., 24)])
    pts[0] = [1., 1.]
    pts[1] = [2., 1.]
    pts[2] = [1., 2.]
    logl = np.array([lf(np0=float(a), np1=float(b)) for a, b in pts])
    only_np0 = np.array([lf(np0=float(a)) for a, _ in pts])
    converted = np.array([[lf._parameter_converter(np0=float(a), np1=float(b))[k]
                           for k in ('op0_rate_multiplier', 'op1_rate_multiplier', 'op2_rate_multiplier')]
                          for a, b in pts])
    no_suffix = lf._parameter_converter(with_suffix=False, np0=2., op1=3.)
    save('reparam', x=x, points=pts, logl=logl, logl_only_np0=only_np0, converted=converted,
         default=np.array([lf()]), bounds_np0=np.array(lf.get_bounds('np0')),
         bounds_all=np.array(lf.get_bounds(), dtype=float),
         no_suffix_keys=np.array(list(no_suffix.keys())), no_suffix_values=np.array(list(no_suffix.values()), dtype=float),
         rate_parameters=np.array(list(lf.rate_parameters.keys()), dtype=str),
         shape_parameters=np.array(list(lf.shape_parameters.keys())))


# ---------------------------------------------------------------------------------------------
# G7: a model with a long contraction -- 3 shape parameters (8 corners per hypercube cell) x 5 sources = 40 terms per
# point-event, the regime of the K-chunk form of K2 (more than 32 terms)
# ---------------------------------------------------------------------------------------------
def golden_long_contraction():
    lf = UnbinnedLogLikelihood(conf_for_test(n_sources=5, events_per_day=300.))
    lf.add_shape_parameter('mu', (-0.5, 0., 0.5))
    lf.add_shape_parameter('sigma', (0.8, 1., 1.3))
    lf.add_shape_parameter('some_multiplier', (0.5, 1., 2.))
    for i in range(5):
        lf.add_rate_parameter('s%d' % i)
    lf.prepare()
    np.random.seed(7)
    d = lf.base_model.simulate()
    lf.set_data(d)
    names = ['s%d_rate_multiplier' % i for i in range(5)] + ['mu', 'sigma', 'some_multiplier']
    rng = np.random.default_rng(8)
    P = 40
    table = np.column_stack([rng.uniform(0.5, 2., (P, 5)), rng.uniform(-0.5, 0.5, P), rng.uniform(0.8, 1.3, P),
                             rng.uniform(0.5, 2., P)])
    table[0] = [1.] * 5 + [0., 1., 1.]                       # the base model
    table[1, 5:] = [-0.5, 0.8, 0.5]                          # anchors
    table[2, 5:] = [0.5, 1.3, 2.]
    table[3, 5] = 0.75                                       # out of range -> -inf
    table[4, 0] = -1.                                        # unphysical -> -inf
    table[5, :5] = 0.                                        # every event an outlier
    logl = np.array([lf(**dict(zip(names, [float(v) for v in row]))) for row in table])
    save('long_contraction', x=d['x'], params=table, logl=logl, names=np.array(names))


if __name__ == '__main__':
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            globals()['golden_' + name]()
        sys.exit(0)
    golden_c1()
    golden_c2('linear')
    golden_c2('piecewise')
    golden_binned(False)
    golden_binned(True)
    golden_multisource()
    golden_sourcewise()
    golden_reparam()
    golden_long_contraction()
