"""Parity of the sm_100a kernels (through the C-ABI, driven by blueice_b200.engine) against the oracle.

Bars (BASELINE.json north_star): bit-exact for cells / fractions / corner indices / weights / bin
indices / mus / per-bin and per-event morphed values; |dlogL| <= 1e-9 * N_events for reduced
quantities (the tests also assert the much tighter bound the canonical summation actually delivers).
"""
import itertools

import numpy as np
import pytest

import bench_workloads as wl
from oracle import binned as obinned
from oracle import hist as ohist
from oracle import morph as omorph
from oracle import unbinned as ounbinned
from oracle.pipeline import BinnedOracle, SourcewiseUnbinnedOracle, UnbinnedOracle

pytestmark = pytest.mark.gpu

TOL_PER_EVENT = 1e-9          # the contract
TIGHT_REL = 2e-13             # what we actually expect relative to |logL| + N


def _engine_mod():
    from blueice_b200 import engine
    return engine


def random_axes(rng, d, one_point_axis=False):
    axes = [np.sort(rng.uniform(-3, 3, rng.integers(2, 6))) for _ in range(d)]
    if one_point_axis and d:
        axes[-1] = np.array([0.7])
    return axes


def random_points(rng, axes, p):
    zs = np.column_stack([rng.uniform(a[0], a[-1], p) for a in axes]) if axes else np.zeros((p, 0))
    for k in range(min(p, 3 * len(axes))):        # hits on anchors (first, middle, last)
        d = k % len(axes)
        zs[k, d] = axes[d][[0, len(axes[d]) // 2, -1][k // len(axes)]]
    return zs


def assert_logl_close(got, ref, n_events, what=""):
    got, ref = np.asarray(got, dtype=float), np.asarray(ref, dtype=float)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), what
    inf = np.isinf(ref)
    assert np.array_equal(got[inf], ref[inf]), what
    fin = np.isfinite(ref)
    diff = np.abs(got[fin] - ref[fin])
    assert np.all(diff <= TOL_PER_EVENT * max(n_events, 1)), (what, diff.max())
    assert np.all(diff <= TIGHT_REL * (np.abs(ref[fin]) + max(n_events, 1))), (what, diff.max())


# ------------------------------------------------------------------------------------------------
# K1: point setup
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,one_point", [(0, False), (1, False), (2, False), (3, False), (4, False), (2, True)])
def test_point_setup_is_bit_exact(d, one_point):
    engine = _engine_mod()
    rng = np.random.default_rng(10 + d)
    axes = random_axes(rng, d, one_point)
    s = 3
    grid = engine.MorphGrid(axes)
    mus_anchor = rng.uniform(0.1, 50, [len(a) for a in axes] + [s])
    p = 257
    zs = random_points(rng, axes, p)
    mult = rng.uniform(0.2, 3, (p, s))
    mult[5, 1] = 0.0
    mult[6, 0] = -1.0
    mult[7, 2] = np.inf
    mult[8, 1] = np.nan
    scale = rng.uniform(0.5, 2, p)
    eff = rng.uniform(0.5, 1, (p, s))
    if d:
        zs[9, 0] = axes[0][-1] + 1.0
        zs[10, 0] = np.nan
        zs[11, d - 1] = axes[d - 1][0] - 1e-9
    eng = engine.UnbinnedEngine(grid, mus_anchor.reshape(grid.n_anchors, s))
    out = eng.point_setup_host(zs, mult, scale, eff)
    for i in range(p):
        in_range = all(a[0] <= z <= a[-1] for a, z in zip(axes, zs[i]))
        assert bool(out["status"][i] & 1) == (not in_range), i
        if not in_range:
            continue
        corners, weights = omorph.corner_table(axes, zs[i]) if d else (np.zeros((1, 0), dtype=int), np.ones(1))
        cells = [omorph.find_cell(a, z) for a, z in zip(axes, zs[i])]
        assert [c[0] for c in cells] == out["cell"][i].tolist(), i
        assert np.array_equal(np.array([c[1] for c in cells]), out["frac"][i]), i
        flat = [int(np.ravel_multi_index(tuple(int(x) % n for x, n in zip(c, grid.shape)), grid.shape)) if d else 0
                for c in corners]
        assert flat == out["corner"][i].tolist(), i
        assert np.array_equal(weights, out["weight"][i]), i
        mus = omorph.morph_explicit(axes, mus_anchor, zs[i]) if d else mus_anchor.copy()
        mus = ounbinned.scale_mus(mus, mult[i], scale[i], np.ones(s, dtype=bool), eff[i])
        assert np.array_equal(mus, out["mus"][i], equal_nan=True), i
        with np.errstate(all='ignore'):
            assert np.array_equal(np.sum(mus), out["musum"][i], equal_nan=True), i
        assert bool(out["status"][i] & 2) == ounbinned.rates_unphysical(mus), i


def test_point_setup_allow_negative_and_many_sources():
    engine = _engine_mod()
    rng = np.random.default_rng(3)
    s = 19                                               # exercises numpy's 8-lane pairwise sum order
    grid = engine.MorphGrid([np.array([0., 1., 2.])])
    mus_anchor = rng.uniform(0.1, 50, (3, s))
    allow = [False] * s
    allow[2] = True
    eng = engine.UnbinnedEngine(grid, mus_anchor, allow_negative=allow)
    p = 64
    zs = rng.uniform(0, 2, (p, 1))
    mult = rng.uniform(0.2, 3, (p, s))
    mult[0, 2] = -0.5          # allowed negative source
    mult[1, 3] = -0.5          # not allowed
    mult[2, 2] = -1e6          # allowed, but the sum goes negative
    mult[3, :] = np.inf
    out = eng.point_setup_host(zs, mult)
    for i in range(p):
        mus = ounbinned.scale_mus(omorph.morph_explicit(grid.axes, mus_anchor, zs[i]), mult[i])
        assert np.array_equal(mus, out["mus"][i]), i
        assert np.sum(mus) == out["musum"][i] or (np.isnan(np.sum(mus)) and np.isnan(out["musum"][i])), i
        assert bool(out["status"][i] & 2) == ounbinned.rates_unphysical(mus, allow), i


# ------------------------------------------------------------------------------------------------
# K2: fused unbinned likelihood
# ------------------------------------------------------------------------------------------------
def make_case(rng, d, s, n, anchors_per_dim=3):
    axes = [np.sort(rng.uniform(-2, 2, anchors_per_dim)) for _ in range(d)]
    shape = [len(a) for a in axes]
    mus_anchor = rng.uniform(5, 500, shape + [s])
    ps_anchor = np.exp(rng.normal(-4, 2, shape + [s, n]))
    return axes, mus_anchor, ps_anchor


def build_engine(axes, mus_anchor, ps_anchor, outlier=1e-12, allow_negative=None):
    engine = _engine_mod()
    grid = engine.MorphGrid(axes)
    s = mus_anchor.shape[-1]
    eng = engine.UnbinnedEngine(grid, mus_anchor.reshape(grid.n_anchors, s), outlier, allow_negative)
    eng.set_ps_anchor(ps_anchor)
    return eng


@pytest.mark.parametrize("d,s,n", [(0, 1, 1), (0, 3, 1000), (1, 1, 1004), (1, 2, 31), (2, 2, 4097), (2, 3, 513),
                                   (3, 3, 1000), (4, 6, 700), (5, 2, 300), (2, 9, 2000), (1, 1, 0)])
def test_unbinned_stream_matches_oracle(d, s, n):
    rng = np.random.default_rng(100 * d + s + n)
    axes, mus_anchor, ps_anchor = make_case(rng, d, s, n)
    eng = build_engine(axes, mus_anchor, ps_anchor)
    p = 9
    zs = random_points(rng, axes, p)
    mult = rng.uniform(0.5, 2, (p, s))
    if d:
        zs[8, 0] = axes[0][-1] + 0.1
    mult[7, 0] = -1
    orc = UnbinnedOracle(axes, mus_anchor).set_ps(ps_anchor)
    ref = orc.batch(zs, mult)
    for mode in (None, 'stream'):                     # None: the engine's own choice (DMMA kernel when C*S <= 32)
        eng.force_kernel = mode
        got = eng.evaluate(zs, mult)
        assert_logl_close(got, ref, n, "%s d=%d s=%d n=%d" % (mode, d, s, n))


@pytest.mark.parametrize("d,s,n", [(0, 1, 1000), (1, 1, 1), (1, 2, 33), (2, 2, 5000), (2, 2, 511), (2, 2, 512),
                                   (2, 2, 513), (3, 3, 2048), (4, 2, 1500), (4, 8, 640), (2, 5, 100000),
                                   # more than 128 contraction terms: the K-chunk kernel (k_unbinned_mma_wide)
                                   (5, 5, 3000), (4, 12, 1111), (3, 40, 513), (5, 8, 64), (4, 9, 1), (5, 33, 130)])
def test_mma_kernel_against_stream_kernel_and_oracle(d, s, n):
    rng = np.random.default_rng(7 * d + s + n)
    axes, mus_anchor, ps_anchor = make_case(rng, d, s, n)
    eng = build_engine(axes, mus_anchor, ps_anchor)
    p = 700
    zs = random_points(rng, axes, p)
    if d:                                             # crowd most points into two cells, some exactly on anchors
        lo, hi = axes[0][0], axes[0][1]
        zs[50:500, 0] = rng.uniform(lo, hi, 450)
    mult = rng.uniform(0.5, 2, (p, s))
    mult[3, 0] = -1.0                                 # unphysical
    res = {}
    for mode in ('stream', None):
        eng.force_kernel = mode
        res[mode] = eng.evaluate(zs, mult)
    plan = eng.plan(zs)
    assert plan.kernel == 'mma'                       # auto mode really exercised the DMMA kernels
    assert_logl_close(res[None], res['stream'], n, "mma vs streaming kernel")
    orc = UnbinnedOracle(axes, mus_anchor).set_ps(ps_anchor)
    check = rng.choice(p, size=12 if n > 20000 else 40, replace=False)
    ref = orc.batch(zs[check], mult[check])
    assert_logl_close(res[None][check], ref, n, "scan d=%d s=%d n=%d" % (d, s, n))


def test_results_do_not_depend_on_batch_shape_or_order():
    rng = np.random.default_rng(5)
    axes, mus_anchor, ps_anchor = make_case(rng, 2, 2, 3000, anchors_per_dim=4)
    eng = build_engine(axes, mus_anchor, ps_anchor)
    p = 900
    zs = random_points(rng, axes, p)
    mult = rng.uniform(0.5, 2, (p, 2))
    full = eng.evaluate(zs, mult)
    for i in (0, 1, 17, 400, 899):
        alone = eng.evaluate(zs[i:i + 1], mult[i:i + 1])
        assert alone[0] == full[i]
    perm = rng.permutation(p)
    assert np.array_equal(eng.evaluate(zs[perm], mult[perm]), full[perm])
    assert np.array_equal(eng.evaluate(zs[:37], mult[:37]), full[:37])
    assert np.array_equal(eng.evaluate(zs[100:613], mult[100:613]), full[100:613])
    assert eng.evaluate(zs[:0], mult[:0]).shape == (0,)


def test_six_shape_parameters_against_the_oracle():
    """6 shape parameters: 64 corners per hypercube cell (x 2 sources = 128 terms, x 3 = 192) -- beyond the streaming kernel's
    32 corners, so the only cross-check is the oracle (scipy's RegularGridInterpolator over the 6-D anchor grid)."""
    for s_count, n in ((2, 700), (3, 300)):
        rng = np.random.default_rng(66 + s_count)
        axes, mus_anchor, ps_anchor = make_case(rng, 6, s_count, n, anchors_per_dim=2)
        eng = build_engine(axes, mus_anchor, ps_anchor)
        assert eng.n_terms == 64 * s_count and eng.uses_mma()
        p = 90
        zs = random_points(rng, axes, p)
        mult = rng.uniform(0.5, 2, (p, s_count))
        got = eng.evaluate(zs, mult)
        orc = UnbinnedOracle(axes, mus_anchor).set_ps(ps_anchor)
        assert_logl_close(got, orc.batch(zs, mult), n, "6 shape parameters, %d sources" % s_count)
        assert eng.evaluate(zs[5:6], mult[5:6])[0] == got[5]


def test_wide_contraction_results_do_not_depend_on_batch_shape_or_order():
    """160 contraction terms (5 shape parameters x 5 sources): the K-chunk kernel, whose units are CTA-wide groups of 64
    points -- a point alone, in a permuted batch or in a slice of the batch gives the same bits."""
    rng = np.random.default_rng(55)
    axes, mus_anchor, ps_anchor = make_case(rng, 5, 5, 1500)
    eng = build_engine(axes, mus_anchor, ps_anchor)
    assert eng.n_terms == 160 and eng.uses_mma()
    p = 600
    zs = random_points(rng, axes, p)
    zs[200:500, 0] = rng.uniform(axes[0][0], axes[0][1], 300)
    zs[200:500, 1:] = [0.5 * (a[0] + a[1]) for a in axes[1:]]
    mult = rng.uniform(0.5, 2, (p, 5))
    full = eng.evaluate(zs, mult)
    assert np.all(np.isfinite(full))
    for i in (0, 1, 17, 250, 599):
        alone = eng.evaluate(zs[i:i + 1], mult[i:i + 1])
        assert alone[0] == full[i]
    perm = rng.permutation(p)
    assert np.array_equal(eng.evaluate(zs[perm], mult[perm]), full[perm])
    assert np.array_equal(eng.evaluate(zs[190:403], mult[190:403]), full[190:403])
    orc = UnbinnedOracle(axes, mus_anchor).set_ps(ps_anchor)
    check = np.r_[0:5, 245:255]
    assert_logl_close(full[check], orc.batch(zs[check], mult[check]), 1500, "wide contraction vs oracle")


@pytest.mark.parametrize("outlier", [1e-12, 0.0])
def test_wide_contraction_nan_inf_zero_and_negative_densities(outlier):
    """The rare path of the K-chunk kernel (rows re-read from global memory) against likelihood.py:686-689."""
    rng = np.random.default_rng(12)
    axes, mus_anchor, ps_anchor = make_case(rng, 4, 9, 777)                   # 144 terms
    specials = [0.0, np.nan, np.inf, -1.0, -np.inf, 1e-320, 1e305]
    for k, v in enumerate(specials):
        ps_anchor[..., 0, 10 + 3 * k] = v                                     # whole column special in source 0
        ps_anchor[1, 0, 2, 1, 4, 40 + 3 * k] = v                              # one anchor, one source only
    ps_anchor[..., 100] = 0.0                                                 # density exactly 0 -> outlier
    ps_anchor[..., 101] = np.nan                                              # all terms NaN -> nansum 0 -> outlier
    ps_anchor[..., 776] = 0.0                                                 # in the last, partial group
    eng = build_engine(axes, mus_anchor, ps_anchor, outlier=outlier)
    orc = UnbinnedOracle(axes, mus_anchor, outlier_likelihood=outlier).set_ps(ps_anchor)
    p = 70
    zs = random_points(rng, axes, p)
    mult = rng.uniform(0.5, 2, (p, 9))
    mult[5, 0] = 0.0
    mult[6, :] = 0.0
    ref = orc.batch(zs, mult)
    for mode in (None, 'stream'):
        eng.force_kernel = mode
        assert_logl_close(eng.evaluate(zs, mult), ref, 777, "wide special values, outlier=%g, %s" % (outlier, mode))


_WIDE_AB_SCRIPT = """
import sys, numpy as np
sys.path.insert(0, %r)
from blueice_b200 import engine
out = {}
for d, s, n in %r:
    rng = np.random.default_rng(1000 + 7 * d + s + n)
    axes = [np.sort(rng.uniform(-2, 2, 3)) for _ in range(d)]
    mus_anchor = rng.uniform(5, 500, [3] * d + [s])
    ps_anchor = np.exp(rng.normal(-4, 2, [3] * d + [s, n]))
    ps_anchor[..., n // 2] = 0.0
    grid = engine.MorphGrid(axes)
    eng = engine.UnbinnedEngine(grid, mus_anchor.reshape(grid.n_anchors, s), 1e-12, None)
    eng.set_ps_anchor(ps_anchor)
    zs = np.column_stack([rng.uniform(a[0], a[-1], 333) for a in axes]) if d else np.zeros((333, 0))
    mult = rng.uniform(0.5, 2, (333, s))
    out['%%d_%%d_%%d' %% (d, s, n)] = eng.evaluate(zs, mult)
np.savez(sys.argv[1], **out)
"""


def test_k_chunk_kernel_is_bitwise_identical_to_the_per_warp_kernel(tmp_path):
    """BI_MMA_WIDE_MIN_TERMS=1 sends every contraction to the K-chunk kernel, BI_MMA_WIDE_MIN_POINTS=huge every contraction of
    up to 128 terms to the per-warp-ring kernel: same fma chain over k, same trees -> same bits (the two runs are separate
    processes: the thresholds are read once)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shapes = [(0, 1, 1000), (2, 2, 5000), (3, 3, 2048), (4, 8, 640), (1, 2, 33), (2, 9, 513), (3, 7, 1200), (4, 5, 700)]
    script = _WIDE_AB_SCRIPT % (root, shapes)
    res = {}
    for tag, val in (('narrow', None), ('wide', '1')):
        env = dict(os.environ)
        env.pop('BI_MMA_WIDE_MIN_TERMS', None)
        env.pop('BI_MMA_WIDE_MIN_POINTS', None)
        if val:
            env['BI_MMA_WIDE_MIN_TERMS'] = val
        else:
            env['BI_MMA_WIDE_MIN_POINTS'] = '1000000000'
        env['BI_SMALL'] = '0'                       # keep the four-launch path for every batch size
        path = str(tmp_path / (tag + '.npz'))
        run = subprocess.run([sys.executable, '-c', script, path], env=env, cwd=root, capture_output=True, text=True)
        assert run.returncode == 0, run.stderr[-2000:]
        res[tag] = dict(np.load(path))
    for key in res['narrow']:
        assert np.array_equal(res['narrow'][key], res['wide'][key], equal_nan=True), key


def test_anchor_hit_equals_unmorphed_tensor():
    """z exactly on anchors -> weights are exactly {0, 1}: same bits as evaluating that anchor without morphing."""
    rng = np.random.default_rng(8)
    axes, mus_anchor, ps_anchor = make_case(rng, 2, 3, 1500, anchors_per_dim=3)
    eng = build_engine(axes, mus_anchor, ps_anchor)
    mult = rng.uniform(0.5, 2, (1, 3))
    for i, j in itertools.product(range(3), range(3)):
        got = eng.evaluate(np.array([[axes[0][i], axes[1][j]]]), mult)
        flat = build_engine([], mus_anchor[i, j][np.newaxis], ps_anchor[i, j][np.newaxis])
        assert got[0] == flat.evaluate(np.zeros((1, 0)), mult)[0]


@pytest.mark.parametrize("outlier", [1e-12, 0.0, 3.5])
def test_nan_inf_zero_and_negative_densities(outlier):
    """likelihood.py:686-689: NaN products dropped, then non-positive / NaN densities -> outlier (if != 0)."""
    rng = np.random.default_rng(11)
    axes, mus_anchor, ps_anchor = make_case(rng, 1, 2, 777)
    specials = [0.0, np.nan, np.inf, -1.0, -np.inf, 1e-320, 1e305]
    for k, v in enumerate(specials):
        ps_anchor[:, 0, 10 + 3 * k] = v               # whole column special in source 0
        ps_anchor[1, 1, 40 + 3 * k] = v               # one anchor only
    ps_anchor[:, :, 100] = 0.0                        # density exactly 0 -> outlier
    ps_anchor[:, :, 101] = np.nan                     # all terms NaN -> nansum 0 -> outlier
    eng = build_engine(axes, mus_anchor, ps_anchor, outlier=outlier)
    orc = UnbinnedOracle(axes, mus_anchor, outlier_likelihood=outlier).set_ps(ps_anchor)
    p = 40
    zs = random_points(rng, axes, p)
    mult = rng.uniform(0.5, 2, (p, 2))
    mult[5, 0] = 0.0                                  # 0 * inf = NaN term, dropped
    mult[6, :] = 0.0                                  # all-zero rates: every event is an outlier
    for mode in (None, 'stream'):
        eng.force_kernel = mode
        got = eng.evaluate(zs, mult)
        ref = orc.batch(zs, mult)
        assert_logl_close(got, ref, 777, "special values, outlier=%g, %s" % (outlier, mode))


@pytest.mark.parametrize("source_dims,n", [([[0, 1], [1], []], 1000), ([[0], [1], [0, 1, 2], [2]], 777),
                                           ([[], []], 100), ([[0, 1, 2, 3]], 513), ([[1], [0], [1], [0], [1]], 2049),
                                           # 45 and 135 contraction terms (odd counts): the K-chunk kernel on arbitrary row lists
                                           ([[0, 1, 2]] * 5 + [[1]] * 2 + [[]], 700),
                                           ([[0, 1, 2, 3]] * 8 + [[0, 2]] + [[3]] + [[]], 300)])
def test_sourcewise_engine_matches_oracle(source_dims, n):
    """bi_point_setup_sourcewise + K2 against one RegularGridInterpolator per source (likelihood.py:210-240,534-555)."""
    engine = _engine_mod()
    rng = np.random.default_rng(31 + n)
    d = 1 + max([max(dims) for dims in source_dims if dims] + [0])
    axes = random_axes(rng, d)
    shapes = [tuple(len(axes[k]) for k in dims) for dims in source_dims]
    mus_sub = [rng.uniform(50, 500, sh) for sh in shapes]
    ps_sub = [rng.uniform(1e-4, 1e-1, sh + (n,)) for sh in shapes]
    ps_sub[0][..., 7] = 0.0                       # source 0 contributes nothing to event 7
    ps_sub[-1][..., 11] = np.nan                  # NaN term dropped by nansum
    grid = engine.MorphGrid(axes)
    eng = engine.SourcewiseUnbinnedEngine(grid, source_dims, np.concatenate([m.reshape(-1) for m in mus_sub]))
    eng.set_ps_anchor(np.concatenate([p.reshape(-1, n) for p in ps_sub]))
    orc = SourcewiseUnbinnedOracle(axes, source_dims, mus_sub).set_ps(ps_sub)
    p = 70
    zs = random_points(rng, axes, p)
    mult = rng.uniform(0.5, 2, (p, len(source_dims)))
    zs[60, 0] = axes[0][-1] + 1.0                 # out of range
    mult[61, 0] = -1.0                            # unphysical
    mult[62, :] = 0.0                             # every event an outlier
    got = eng.evaluate(zs, mult)
    ref = orc.batch(zs, mult)
    assert_logl_close(got, ref, n, "source-wise %s" % (source_dims,))
    for i in (0, 1, 33):
        assert eng.evaluate(zs[i:i + 1], mult[i:i + 1])[0] == got[i]
        ll, mus, ps = orc(zs[i], mult[i], full_output=True)
        dev_mus, dev_ps = eng.ps(zs[i], mult[i])
        assert np.array_equal(dev_mus, mus)                           # bit-exact per-source morph of the rates
        assert np.array_equal(dev_ps, ps, equal_nan=True)             # bit-exact per-source morph of the pdf values


def test_zero_rates_give_n_log_outlier():
    # rate_multiplier = 0 is legal, every event gets outlier_likelihood
    axes, mus_anchor, ps_anchor, x = wl.c1_arrays(seed=2)
    eng = build_engine(axes, mus_anchor, ps_anchor)
    got = eng.evaluate(np.array([[0.3]]), np.array([[0.0]]))
    assert abs(got[0] - len(x) * np.log(1e-12)) <= 1e-9 * len(x)


def test_ps_output_is_bit_exact():
    rng = np.random.default_rng(12)
    for d in (0, 1, 2, 3):
        axes, mus_anchor, ps_anchor = make_case(rng, d, 3, 333)
        eng = build_engine(axes, mus_anchor, ps_anchor)
        zs = random_points(rng, axes, 4)
        for z in zs:
            mus, ps = eng.ps(z, np.ones(3))
            ref = omorph.morph_explicit(axes, ps_anchor, z) if d else ps_anchor
            assert np.array_equal(ps, ref)
            ref_mu = omorph.morph_explicit(axes, mus_anchor, z) if d else mus_anchor
            assert np.array_equal(mus, ref_mu)


def test_parts_output_recombines():
    rng = np.random.default_rng(13)
    axes, mus_anchor, ps_anchor = make_case(rng, 2, 2, 1234)
    eng = build_engine(axes, mus_anchor, ps_anchor)
    zs = random_points(rng, axes, 50)
    mult = rng.uniform(0.5, 2, (50, 2))
    ll = eng.evaluate(zs, mult)
    logsum, musum, status = eng.evaluate(zs, mult, return_parts=True)
    assert np.array_equal(-musum + logsum, ll)
    # two event shards, aligned to the 512-event superblock: partial sums add up (rank-order sum)
    a = build_engine(axes, mus_anchor, ps_anchor[..., :1024]).evaluate(zs, mult, return_parts=True)[0]
    b = build_engine(axes, mus_anchor, ps_anchor[..., 1024:]).evaluate(zs, mult, return_parts=True)[0]
    np.testing.assert_allclose(a + b, logsum, rtol=1e-14)


# ------------------------------------------------------------------------------------------------
# K3: template lookup and binning
# ------------------------------------------------------------------------------------------------
def lookup_inputs(rng, shape, n):
    edges = [np.sort(rng.uniform(0, 10, b + 1)) for b in shape]
    coords = [rng.uniform(e[0] - 0.5, e[-1] + 0.5, n) for e in edges]
    for c, e in zip(coords, edges):                   # events exactly on edges, centres and far outside
        c[:len(e)] = e
        c[len(e):2 * len(e) - 1] = ohist.bin_centers(e)
        c[2 * len(e)] = e[0] - 100
        c[2 * len(e) + 1] = e[-1] + 100
    return edges, coords


@pytest.mark.parametrize("shape", [(30,), (1,), (11, 7), (100, 100), (5, 4, 3), (3, 1, 4), (3, 2, 4, 3)])
def test_hist_lookup_linear_is_bit_exact(shape):
    from blueice_b200 import device_ops
    rng = np.random.default_rng(len(shape) * 31 + shape[0])
    n = 3001
    edges, coords = lookup_inputs(rng, shape, max(n, 2 * max(shape) + 10))
    templates = rng.random((5,) + shape)
    got, idx = device_ops.hist_lookup(templates, edges, coords, 'linear', return_bin_index=True)
    if all(b >= 2 for b in shape):
        for t in range(5):
            assert np.array_equal(got[t], ohist.lookup_linear(templates[t], edges, coords)), t
    for t in range(5):
        assert np.array_equal(got[t], ohist.lookup_linear_explicit(templates[t], edges, coords)), t
    # bit-exact lower-corner cell index
    cells = [omorph.find_cells(ohist.bin_centers(e), np.clip(c, ohist.bin_centers(e).min(), ohist.bin_centers(e).max()))[0]
             for e, c in zip(edges, coords)]
    cells = [np.where(c < 0, b - 1, c) for c, b in zip(cells, shape)]
    assert np.array_equal(idx, np.ravel_multi_index(cells, shape))


@pytest.mark.parametrize("shape", [(30,), (11, 7), (5, 4, 3)])
def test_hist_lookup_piecewise_matches_oracle(shape):
    from blueice_b200 import device_ops
    rng = np.random.default_rng(len(shape) * 17)
    edges, coords = lookup_inputs(rng, shape, 2500)
    templates = rng.random((3,) + shape)
    got, idx = device_ops.hist_lookup(templates, edges, coords, 'piecewise', return_bin_index=True)
    ref_idx = ohist.lookup_piecewise_indices(edges, coords)
    assert np.array_equal(idx, np.ravel_multi_index(ref_idx, shape))
    for t in range(3):
        assert np.array_equal(got[t], ohist.lookup_piecewise(templates[t], edges, coords))


def test_hist_lookup_rejects_nan_for_linear_like_scipy():
    from blueice_b200 import device_ops
    edges = [np.linspace(0, 1, 5)]
    with pytest.raises(ValueError):
        device_ops.hist_lookup(np.ones((1, 4)), edges, [np.array([0.5, np.nan])], 'linear')
    assert device_ops.hist_lookup(np.ones((1, 4)), edges, [np.zeros(0)], 'linear').shape == (1, 0)


@pytest.mark.parametrize("shape", [(7,), (11, 7), (5, 4, 3)])
def test_histogramdd_matches_numpy(shape):
    from blueice_b200 import device_ops
    rng = np.random.default_rng(len(shape) * 13)
    edges, coords = lookup_inputs(rng, shape, 20000)
    coords[0][50] = np.nan
    got, idx = device_ops.histogramdd(edges, coords, return_bin_index=True)
    assert np.array_equal(got, ohist.histogramdd(edges, coords))
    assert np.array_equal(idx, ohist.histogramdd_indices(edges, coords))
    assert got.sum() < 20000 and idx[50] == -1


# ------------------------------------------------------------------------------------------------
# K4: binned Poisson + Beeston-Barlow
# ------------------------------------------------------------------------------------------------
def binned_case(rng, d, s, bins, total=3000.):
    axes = [np.sort(rng.uniform(-2, 2, 3)) for _ in range(d)]
    shape = [len(a) for a in axes]
    n_bins = int(np.prod(bins))
    pmf = rng.random(shape + [s] + list(bins)) + 0.05
    pmf /= pmf.reshape(shape + [s, n_bins]).sum(axis=-1).reshape(shape + [s] + [1] * len(bins))
    mus = rng.uniform(0.1, 1, shape + [s]) * total
    n_model = 1.0 + rng.poisson(30.0 * pmf * n_bins).astype(float)
    centre = tuple(k // 2 for k in shape)
    lam = np.tensordot(mus[centre], pmf[centre], axes=(0, 0))
    observed = rng.poisson(lam).astype(float)
    return axes, mus, pmf, n_model, observed


@pytest.mark.parametrize("d,s,bins,bb", [(0, 1, (1,), None), (0, 2, (4,), 0), (1, 2, (40,), None), (2, 3, (6, 5, 4), 0),
                                         (2, 4, (30, 20), 2), (3, 4, (20, 20, 5), 0), (1, 3, (1100,), None)])
def test_binned_matches_oracle(d, s, bins, bb):
    engine = _engine_mod()
    rng = np.random.default_rng(1000 + 10 * d + s)
    axes, mus, pmf, n_model, observed = binned_case(rng, d, s, bins)
    grid = engine.MorphGrid(axes)
    eng = engine.BinnedEngine(grid, mus.reshape(grid.n_anchors, s), pmf, n_model if bb is not None else None, bb)
    eng.set_observed(observed)
    orc = BinnedOracle(axes, mus, pmf, n_model if bb is not None else None, bb).set_observed(observed)
    p = 23
    zs = random_points(rng, axes, p)
    mult = rng.uniform(0.5, 1.5, (p, s))
    mult[4, 0] = -1.0
    if d:
        zs[5, 0] = axes[0][0] - 1
    got, status, flags = eng.evaluate(zs, mult, return_status=True)
    assert np.all(flags == 0)
    ref = orc.batch(zs, mult)
    n_bins = int(np.prod(bins))
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))
    fin = np.isfinite(ref)
    assert np.all(np.abs(got[fin] - ref[fin]) <= 1e-9 * max(observed.sum(), n_bins))
    assert np.all(np.abs(got[fin] - ref[fin]) <= 1e-12 * (np.abs(ref[fin]) + n_bins))
    # adjusted mus and pmfs of one point (full_output)
    i = 9
    ll_i, mus_adj, pmfs_adj, fl = eng.pmfs(zs[i], mult[i])
    ref_ll, ref_mus, ref_pmfs = orc(zs[i], mult[i], full_output=True)
    assert ll_i == got[i]
    if bb is None:
        assert np.array_equal(pmfs_adj, ref_pmfs) and np.array_equal(mus_adj, ref_mus)
    else:
        np.testing.assert_allclose(pmfs_adj, ref_pmfs, rtol=1e-11)
        np.testing.assert_allclose(mus_adj, ref_mus, rtol=1e-12)


def test_binned_poisson_edge_semantics():
    """scipy.stats.poisson(lam).logpmf(k): lam = 0 & k = 0 -> 0; lam = 0 < k -> -inf; lam < 0 or NaN -> NaN."""
    engine = _engine_mod()
    pmf = np.array([[[0.0, 0.5, 0.5, 0.0]]])                  # [G=1, S=1, B=4]
    eng = engine.BinnedEngine(engine.MorphGrid([]), np.array([[10.0]]), pmf)
    with np.errstate(all='ignore'):
        for observed in ([0., 4., 6., 0.], [1., 4., 6., 0.], [0., 0., 0., 0.]):
            eng.set_observed(np.array(observed))
            got = eng.evaluate(np.zeros((3, 0)), np.array([[1.0], [0.0], [2.5]]))
            ref = [obinned.binned_loglikelihood([10.0 * m], pmf[0], np.array(observed)) for m in (1.0, 0.0, 2.5)]
            assert np.array_equal(np.isneginf(got), np.isneginf(ref))
            fin = np.isfinite(ref)
            np.testing.assert_allclose(got[fin], np.asarray(ref)[fin], rtol=1e-14)
    pmf_nan = np.array([[[np.nan, 0.5, 0.5, 0.0]]])
    eng = engine.BinnedEngine(engine.MorphGrid([]), np.array([[10.0]]), pmf_nan).set_observed(np.array([0., 4., 6., 0.]))
    assert np.isnan(eng.evaluate(np.zeros((1, 0)), np.array([[1.0]]))[0])


def test_beeston_barlow_flags_follow_the_reference_asserts():
    engine = _engine_mod()
    rng = np.random.default_rng(2)
    axes, mus, pmf, n_model, observed = binned_case(rng, 0, 2, (12,))
    n_model[0, 3] = 0.0                                     # bin without calibration events: 0/0 -> assert
    eng = engine.BinnedEngine(engine.MorphGrid([]), mus.reshape(1, 2), pmf, n_model, 0).set_observed(observed)
    _, status, flags = eng.evaluate(np.zeros((1, 0)), np.ones((1, 2)), return_status=True)
    assert flags[0] != 0
    orc = BinnedOracle([], mus, pmf, n_model, 0).set_observed(observed)
    with pytest.raises(AssertionError):
        orc(np.zeros(0), np.ones(2))


def test_beeston_barlow_known_answers_through_the_kernel():
    """tests/test_BeestonBarlow.py:68-76 of the reference: A = root2([16,30,32,27], 0.2, 0, [3,5,2,7])."""
    engine = _engine_mod()
    from scipy import stats
    a = np.array([16., 30., 32., 27.])
    d = np.array([3., 5., 2., 7.])
    pmf = (a / a.sum())[np.newaxis, np.newaxis, :]
    eng = engine.BinnedEngine(engine.MorphGrid([]), np.array([[21.0]]), pmf, a[np.newaxis, np.newaxis, :], 0)
    eng.set_observed(d)
    got = eng.evaluate(np.zeros((1, 0)), np.ones((1, 1)))[0]
    A = obinned.beeston_barlow_root2(a, 0.2, np.array([0.]), d)
    np.testing.assert_almost_equal(A, [15.833, 29.166, 28.333, 28.333], decimal=2)
    assert abs((got - np.sum(stats.poisson(0.2 * A).logpmf(d))) / got) <= 1e-6


@pytest.mark.parametrize("d,s,bins,bb,p", [(0, 2, (5,), 0, 3), (1, 3, (1100,), None, 70), (2, 3, (37, 29), 1, 300),
                                           (3, 4, (20, 20, 7), 0, 90), (2, 2, (9, 7), 0, 1500), (4, 2, (25, 13), None, 40)])
def test_binned_tiled_kernel_is_bitwise_identical_to_the_gather_kernel(d, s, bins, bb, p, monkeypatch):
    """k_binned_tile (anchor rows of a 256-bin tile staged in shared memory by TMA, points grouped by hypercube cell,
    t_b kept between the Beeston-Barlow passes) against k_binned_pass (BI_BINNED_LEGACY=1): same bits, for bin counts
    with ragged tiles and blocks, groups of more than 32 points per cell, several point passes, status != 0 points."""
    engine = _engine_mod()
    rng = np.random.default_rng(7000 + 10 * d + s)
    axes, mus, pmf, n_model, observed = binned_case(rng, d, s, bins)
    grid = engine.MorphGrid(axes)
    eng = engine.BinnedEngine(grid, mus.reshape(grid.n_anchors, s), pmf, n_model if bb is not None else None, bb)
    eng.set_observed(observed)
    zs = random_points(rng, axes, p)
    mult = rng.uniform(0.5, 1.5, (p, s))
    mult[min(4, p - 1), 0] = -1.0                             # unphysical -> -inf, no work
    if d:
        zs[min(5, p - 1), 0] = axes[0][0] - 1                 # out of range
        zs[0] = [a[-1] for a in axes]                         # on the upper boundary
    monkeypatch.setenv("BI_BINNED_LEGACY", "1")
    want, st_w, fl_w = eng.evaluate(zs, mult, return_status=True)
    want_full = eng.pmfs(zs[0], mult[0])
    monkeypatch.setenv("BI_BINNED_LEGACY", "0")
    got, st_g, fl_g = eng.evaluate(zs, mult, return_status=True)
    got_full = eng.pmfs(zs[0], mult[0])
    assert np.array_equal(got, want) and np.array_equal(st_g, st_w) and np.array_equal(fl_g, fl_w)
    assert got_full[0] == want_full[0] and np.array_equal(got_full[1], want_full[1])
    assert np.array_equal(got_full[2], want_full[2])
    assert np.isfinite(got).sum() >= p - 2
    # one observed histogram per point (binned toys)
    rows = rng.poisson(np.maximum(observed.reshape(1, -1), 0.3), size=(p, observed.size)).astype(float)
    eng.set_observed_rows(rows)
    monkeypatch.setenv("BI_BINNED_LEGACY", "1")
    want_t = eng.evaluate_toys(zs, mult)
    monkeypatch.setenv("BI_BINNED_LEGACY", "0")
    got_t = eng.evaluate_toys(zs, mult)
    assert np.array_equal(got_t, want_t)


@pytest.mark.parametrize("d,s,n,p", [(1, 1, 1012, 1), (0, 2, 1, 3), (1, 2, 31, 5), (2, 2, 32, 1), (2, 3, 33, 9),
                                     (2, 2, 511, 4), (2, 2, 512, 4), (2, 2, 513, 60), (3, 3, 5000, 17), (4, 6, 700, 2),
                                     (2, 5, 8192, 100), (5, 4, 2000, 3)])
def test_single_launch_path_is_bitwise_identical_to_the_four_launches(d, s, n, p, monkeypatch):
    """bi_unbinned_ll_small (K1 + K2 + finalize in one launch, inputs and results in pinned host memory) against the
    K1 -> schedule -> k_unbinned_mma -> finalize sequence (BI_SMALL=0): same bits for logL, log sum, mu sum and status,
    including special densities (the reference-semantics fallback), out-of-range and unphysical points, live-time
    scaling; and the lean scalar runner returns the batch's numbers."""
    rng = np.random.default_rng(9000 + 100 * d + s + n)
    axes, mus_anchor, ps_anchor = make_case(rng, d, s, n)
    if n > 40:
        for k, v in enumerate([0.0, np.nan, np.inf, -1.0, 1e-320, 1e305]):
            ps_anchor[..., 0, 3 + 5 * k] = v
        ps_anchor[..., :, 37] = 0.0
    eng = build_engine(axes, mus_anchor, ps_anchor)
    zs = random_points(rng, axes, p)
    mult = rng.uniform(0.5, 2, (p, s))
    if p > 3:
        mult[2, 0] = -1.0
        if d:
            zs[3, 0] = axes[0][-1] + 0.1
    scale = rng.uniform(0.5, 2.0, p)
    assert eng._small_ok(p)
    res = {}
    for small in ("0", "1"):
        monkeypatch.setenv("BI_SMALL", small)
        assert eng._small_ok(p) == (small == "1")
        launches = eng.launches
        res[small] = (eng.evaluate(zs, mult, return_status=True), eng.evaluate(zs, mult, scale=scale, return_parts=True),
                      eng.evaluate(zs, mult, return_status=True))          # the third call replays the cached state
        assert eng.launches - launches == (3 if small == "1" else 12)
    for a, b in zip(res["0"], res["1"]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y, equal_nan=True)
    pin, run = eng.scalar_runner(False)
    pin[:d] = zs[0]
    pin[d:d + s] = mult[0]
    logl, status = run()
    assert logl == res["1"][0][0][0] and status == res["1"][0][1][0]
    orc = UnbinnedOracle(axes, mus_anchor).set_ps(ps_anchor)
    assert_logl_close(res["1"][0][0], orc.batch(zs, mult), n, "single launch vs oracle")
