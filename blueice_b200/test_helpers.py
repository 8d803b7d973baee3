"""Fixtures shared by the tests and by bench.py's config-1 workload.

Same names and behaviour as blueice/test_helpers.py (GaussianSource :22-37, GaussianMCSource :40-43,
FixedSampleSource :46-52, BASE_CONFIG :55-66, BASE_CONV_CONFIG :70-76, conf_for_test :79-84,
conf_for_reparam_test :86-97, make_data :103-126,
almost_equal :99-100), so that the parity tests read like the reference's own tests.
"""
from copy import deepcopy

import numpy as np
from scipy import stats

from .source import DensityEstimatingSource, MonteCarloSource, Source
from .utils import combine_dicts


class GaussianSourceBase(Source):
    """1-d Gaussian in 'x' with mean config['mu'] and width config['sigma']: the event generator."""

    def _norm(self):
        return stats.norm(self.config['mu'], self.config['sigma'])

    def simulate(self, n_events):
        events = np.zeros(n_events, dtype=[('x', float), ('source', int)])
        events['x'] = self._norm().rvs(n_events)
        return events


class GaussianSource(GaussianSourceBase):
    """Analytic pdf (evaluated on the host by scipy; its rows are uploaded to the device in set_data)."""

    def compute_pdf(self):
        # two settings that only change the rate: a numeric one and a non-numeric one (length of a string)
        self.events_per_day *= self.config.get('some_multiplier', 1)
        self.events_per_day *= len(self.config.get('strlen_multiplier', 'x'))
        super().compute_pdf()

    def pdf(self, *coordinates):
        if not self.pdf_has_been_computed:
            raise RuntimeError("Trying to call a PDF that hasn't been computed!")
        return self._norm().pdf(coordinates[0])


class GaussianMCSource(GaussianSourceBase, MonteCarloSource):
    """Same Gaussian, but with the pdf estimated from its own Monte Carlo sample."""


class FixedSampleSource(DensityEstimatingSource):
    """Density estimated from the sample passed as config['data']."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.events_per_day *= len(self.config.get('strlen_multiplier', 'x'))

    def get_events_for_density_estimate(self):
        sample = self.config['data']
        return sample, len(sample)


BASE_CONFIG = dict(
    sources=[{'name': 's0', 'events_per_day': 1000.}],
    mu=0,
    strlen_multiplier='q',
    events_per_day=1000.,
    n_events_for_pdf=int(1e6),
    sigma=1,
    default_source_class=GaussianSource,
    some_multiplier=1,
    force_pdf_recalculation=True,     # (sic) the reference's fixture uses this non-existent key
    analysis_space=[['x', np.linspace(-10, 10, 100)]],
)


def conf_for_test(n_sources=1, mc=False, **kwargs):
    conf = deepcopy(BASE_CONFIG)
    conf['sources'] = [{'name': 's%d' % i} for i in range(n_sources)]
    if mc:
        conf['default_source_class'] = GaussianMCSource
    return combine_dicts(conf, kwargs)


# conv_config of the re-parameterisation tests (blueice/test_helpers.py:70-76): two new parameters np0, np1 drive the
# rates of the three sources op0, op1, op2 of conf_for_reparam_test
BASE_CONV_CONFIG = dict(
    np0=(np.linspace(1e-12, 10, 2), None, None),
    np1=(np.linspace(1e-12, 10, 2), None, None),
    op0_rate_multiplier=dict(params=["np0"], func=lambda np0: np0**2),
    op1_rate_multiplier=dict(params=["np1"], func=lambda np1: np1**2),
    op2_rate_multiplier=dict(params=["np0", "np1"], func=lambda np0, np1: np0*np1),
)


def conf_for_reparam_test(n_source=1, mc=False, **kwargs):
    """conf_for_test with the sources op0..op2 and the new parameters np0 = np1 = 1 (blueice/test_helpers.py:86-97)."""
    conf = conf_for_test(n_source, mc, **kwargs)
    conf["sources"] = [dict(name="op0"), dict(name="op1"), dict(name="op2")]
    conf["np0"] = 1
    conf["np1"] = 1
    return conf


def almost_equal(a, b, fraction=1e-6):
    return abs((a - b) / a) <= fraction


def make_data(instructions):
    """make_data([dict(n_events=24, x=0.5), dict(n_events=56, x=1.5)]) -> (record array, total count)."""
    total = sum(item['n_events'] for item in instructions)
    d = np.zeros(total, dtype=[('source', int), ('x', float), ('y', float)])
    start = 0
    for item in instructions:
        stop = start + item['n_events']
        for field, value in item.items():
            if field != 'n_events':
                d[field][start:stop] = value
        start = stop
    return d, total
