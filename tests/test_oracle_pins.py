"""Pin the oracle restatement against the third-party numerics the reference calls (SciPy / NumPy)
and against the known answers of the reference's own tests.  CPU only."""
import itertools

import numpy as np
import pytest
from scipy import stats
from scipy.interpolate import RegularGridInterpolator

from oracle import binned, hist, morph, unbinned


def test_find_cell_matches_scipy_find_indices():
    rng = np.random.default_rng(0)
    for n in (2, 3, 5, 17):
        axis = np.sort(rng.uniform(-3, 3, n))
        itp = RegularGridInterpolator((axis,), np.arange(float(n)))
        zs = np.concatenate([rng.uniform(axis[0], axis[-1], 2000), axis])
        idx, frac = itp._find_indices(zs[np.newaxis, :])
        mine_i, mine_y = morph.find_cells(axis, zs)
        assert np.array_equal(idx[0], mine_i)
        assert np.array_equal(frac[0], mine_y)
        for z in zs[:50]:
            i, y = morph.find_cell(axis, z)
            k = np.searchsorted(zs, z)
            assert (i, y) == (mine_i[np.nonzero(zs == z)[0][0]], mine_y[np.nonzero(zs == z)[0][0]])


def test_anchor_cell_choice_side_right():
    axis = np.array([-2., -1., 0.5, 1., 2.])
    assert morph.find_cell(axis, -2.) == (0, 0.0)
    assert morph.find_cell(axis, -1.) == (1, 0.0)          # interior anchor: upper cell, y == 0
    assert morph.find_cell(axis, 2.) == (3, 1.0)           # last anchor: cell n-2, y == 1
    assert morph.find_cell(np.array([1.0]), 1.0) == (-1, 0.0)


@pytest.mark.parametrize("d", [1, 2, 3, 4])
def test_morph_explicit_is_bitwise_scipy(d):
    rng = np.random.default_rng(d)
    axes = [np.sort(rng.uniform(-3, 3, rng.integers(2, 6))) for _ in range(d)]
    values = rng.random([len(a) for a in axes] + [3, 7])
    itp = morph.morph_rgi(axes, values)
    for t in range(300):
        zs = np.array([rng.uniform(a[0], a[-1]) for a in axes])
        if t % 7 == 0:
            k = rng.integers(d)
            zs[k] = axes[k][rng.integers(len(axes[k]))]
        assert np.array_equal(itp(zs), morph.morph_explicit(axes, values, zs))


def test_anchor_points_c_order():
    axes = morph.anchor_axes([[2, -2, 0], [1, 0]])
    pts = morph.anchor_points(axes)
    assert pts[0] == (-2.0, 0.0) and pts[1] == (-2.0, 1.0) and pts[-1] == (2.0, 1.0)
    assert len(pts) == 6


@pytest.mark.parametrize("shape", [(30,), (11, 7), (5, 4, 3)])
def test_lookup_linear_explicit_is_bitwise_scipy(shape):
    rng = np.random.default_rng(len(shape))
    edges = [np.sort(rng.uniform(0, 10, n + 1)) for n in shape]
    h = rng.random(shape)
    n = 2000
    coords = [rng.uniform(e[0] - 0.5, e[-1] + 0.5, n) for e in edges]
    for c, e in zip(coords, edges):                      # events on edges and centres
        c[:len(e)] = e
        c[len(e):2 * len(e) - 1] = hist.bin_centers(e)
    a = hist.lookup_linear(h, edges, coords)
    b = hist.lookup_linear_explicit(h, edges, coords)
    assert np.array_equal(a, b)


def test_histogramdd_indices_match_numpy():
    rng = np.random.default_rng(5)
    edges = [np.linspace(0, 1, 6), np.array([0., 0.5, 2., 3.])]
    n = 5000
    x = rng.uniform(-0.2, 1.2, n)
    y = rng.uniform(-0.5, 3.5, n)
    x[:6] = edges[0]
    y[:4] = edges[1]
    x[10] = np.nan
    ref = hist.histogramdd(edges, [x, y])
    idx = hist.histogramdd_indices(edges, [x, y])
    mine = np.bincount(idx[idx >= 0], minlength=15).reshape(5, 3).astype(float)
    assert np.array_equal(ref, mine)
    assert idx[10] == -1


def test_poisson_logpmf_matches_scipy():
    lams = np.array([-1.0, 0.0, np.nan, np.inf, 1.5, 1e-320, 1e300, 7.25, 1000.])
    ks = np.array([-1.0, 0.0, 1.0, 1.5, np.nan, 3.0, np.inf, 250.])
    for lam in lams:
        with np.errstate(all='ignore'):
            ref = stats.poisson(lam).logpmf(ks)
        mine = binned.poisson_logpmf(ks, lam)
        assert np.array_equal(ref, mine, equal_nan=True), (lam, ref, mine)


def test_beeston_barlow_known_answers():
    # tests/test_BeestonBarlow.py:32 of the reference
    r = binned.beeston_barlow_root2(np.array([32]), 0.2, np.array([1]), np.array([2]))[0]
    assert abs((28.0814209 - r) / 28.0814209) <= 1e-6       # the reference's almost_equal
    # :68-71
    a = binned.beeston_barlow_root2(np.array([16, 30, 32, 27]), 0.2, np.array([0.]), np.array([3, 5, 2, 7]))
    np.testing.assert_almost_equal(a, [15.833, 29.166, 28.333, 28.333], decimal=2)
    # :120-123
    a = binned.beeston_barlow_root2(np.array([16, 30, 32, 27]), 0.2, np.array([5, 7, 1, 3]), np.array([3, 5, 2, 7]))
    np.testing.assert_almost_equal(a, [14.24, 26.8070, 28.08, 26.21], decimal=2)


def test_extended_loglikelihood_known_answers():
    # tests/test_likelihood.py:17-18 of the reference: one event at x = 0
    ps = np.array([[stats.norm.pdf(0)]])
    assert unbinned.extended_loglikelihood(np.array([1.]), ps, 1e-12) == -1 + stats.norm.logpdf(0)
    assert unbinned.extended_loglikelihood(np.array([2.]), ps, 1e-12) == -2 + np.log(2 * stats.norm.pdf(0))


def test_extended_loglikelihood_nan_and_outlier_semantics():
    # mu=[0,2], ps=[[inf,1,nan,0],[.5,0,1,0]] -> p=[1,0,2,0]
    mu = np.array([0., 2.])
    ps = np.array([[np.inf, 1, np.nan, 0], [.5, 0, 1, 0]])
    expect = -2 + np.log(1) + np.log(1e-12) + np.log(2) + np.log(1e-12)
    assert unbinned.extended_loglikelihood(mu, ps, 1e-12) == expect
    assert unbinned.extended_loglikelihood(mu, ps, 0.0) == -np.inf


def test_numpy_small_sum_order():
    """np.sum over S <= 128 float64: sequential for S < 8, eight interleaved lanes otherwise."""
    rng = np.random.default_rng(1)

    def pw(a):
        n = len(a)
        if n < 8:
            r = 0.0
            for v in a:
                r = r + v
            return r
        r = list(a[:8])
        i = 8
        while i < n - (n % 8):
            for k in range(8):
                r[k] = r[k] + a[i + k]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
        return res

    for n in (1, 2, 3, 7, 8, 9, 16, 17, 31, 64):
        for _ in range(200):
            a = rng.random(n) * 10.0 ** rng.integers(-3, 3, n)
            assert np.sum(a) == pw(a)


def test_rates_unphysical_rule():
    assert not unbinned.rates_unphysical([0., 1.])
    assert unbinned.rates_unphysical([-1., 1.])
    assert unbinned.rates_unphysical([np.inf, 1.])
    assert unbinned.rates_unphysical([np.nan, 1.])
    assert not unbinned.rates_unphysical([-1., 2.], [True, False])
    assert unbinned.rates_unphysical([-1., 2.], [False, True])
    assert unbinned.rates_unphysical([-3., 2.], [True, False])          # sum < 0


def test_philox4x32_10_known_answers():
    """Random123 known-answer vectors (kat_vectors: philox4x32 10) pin the oracle's counter-based generator,
    which in turn pins the device generator bit for bit (tests/test_gpu_toys.py)."""
    from oracle.toys import philox4x32_10, uniform53
    cases = [
        ([0, 0, 0, 0], (0, 0), [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, (0xffffffff, 0xffffffff), [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0),
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in cases:
        got = philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert [int(v) for v in got] == want
    # NumPy's 53-bit construction: in [0, 1), top value just below 1
    u = uniform53(np.array([0, 0xffffffff], dtype=np.uint32), np.array([0, 0xffffffff], dtype=np.uint32))
    assert u[0] == 0.0 and u[1] == 1.0 - 2.0 ** -53


def test_toy_event_stage_matches_histdd_get_random_rule():
    """Given the uniforms, oracle.toys.toy_events places events exactly like Histdd.get_random (cdf search on the
    flattened pmf, uniform inside the bin) and groups them by source like Model.simulate."""
    from oracle import toys
    rng = np.random.default_rng(0)
    edges = [np.linspace(0., 10., 6), np.array([0., 1., 4., 9.])]
    pmf = rng.uniform(0.1, 1., size=(2, 5, 3))
    cdf = np.vstack([np.cumsum(p.ravel()) / p.sum() for p in pmf])
    counts = np.array([[3, 2], [0, 4], [0, 0], [5, 0]])
    coords, source, offsets = toys.toy_events(edges, cdf, counts, seed=7, first_toy=100)
    assert list(offsets) == [0, 5, 9, 9, 14]
    assert list(source) == [0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0]
    assert np.all((coords[0] >= 0) & (coords[0] < 10) & (coords[1] >= 0) & (coords[1] < 9))
    # event (toy 101, index 2) by hand
    u_bin, u0 = toys.toy_uniforms(7, [101], [2], 1)
    u1, _ = toys.toy_uniforms(7, [101], [2], 2)
    flat = min(int(np.searchsorted(cdf[1], u_bin[0])), 14)
    ix, iy = divmod(flat, 3)
    e = 5 + 2
    assert coords[0, e] == edges[0][ix] + u0[0] * (edges[0][ix + 1] - edges[0][ix])
    assert coords[1, e] == edges[1][iy] + u1[0] * (edges[1][iy + 1] - edges[1][iy])
    # sharding invariance: toys 2..3 generated alone are the same events
    c2, s2, _ = toys.toy_events(edges, cdf, counts[2:], seed=7, first_toy=102)
    assert np.array_equal(c2, coords[:, 9:]) and np.array_equal(s2, source[9:])
