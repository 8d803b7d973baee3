"""Event sources: the objects that provide pdf values, pmf grids, rates and toy events.

Host-side mirror of blueice/source.py (Source :33-189, HistogramPdfSource :192-267,
DensityEstimatingSource :270-317, MonteCarloSource :320-348).  Model construction, hashing and the
on-disk pdf cache are the cold path and stay ordinary Python.  What changes for the hot path:

  * `HistogramPdfSource.template()` exposes (densities, bin edges, lookup method) so that
    UnbinnedLogLikelihood.set_data can evaluate all anchors x sources with ONE device gather
    (bi_hist_lookup, K3) instead of G*S host interpolations (likelihood.py:557-560).
  * `HistogramPdfSource.pdf(*coords)` itself -- the public per-source call of the reference -- runs the
    same K3 kernel for a single template.

Config keys, defaults, cache attributes and error behaviour follow the reference.
"""
import inspect
import os

import numpy as np

from . import utils
from .data_reading import read_files_in
from .exceptions import PDFNotComputedException
from .hist import Histdd

__all__ = ['Source', 'HistogramPdfSource', 'DensityEstimatingSource', 'MonteCarloSource']

_SOURCE_DEFAULTS = dict(
    name='unnamed_source', label='Unnamed source', color='black',
    events_per_day=0,            # all events this source produces per day, detected or not
    rate_multiplier=1,           # independent of the likelihood's <name>_rate_multiplier
    fraction_in_range=1,         # fraction of simulated events inside the analysis space
    cache_attributes=[],         # attributes stored in / restored from the pdf cache
    delay_pdf_computation=False,
    dont_hash_settings=[], extra_dont_hash_settings=[],
    force_recalculation=False,   # never read the cache (still writes it)
    never_save_to_cache=False,   # never write the cache (still reads it)
    cache_dir='pdf_cache', task_dir='pdf_tasks')

_ALWAYS_CACHED = ['fraction_in_range', 'events_per_day', 'pdf_has_been_computed']
_NEVER_HASHED = ['hash', 'rate_multiplier', 'force_recalculation', 'never_save_to_cache',
                 'dont_hash_settings', 'label', 'color', 'extra_dont_hash_settings',
                 'delay_pdf_computation', 'cache_dir', 'task_dir']


class Source(object):
    """Base class: config defaults, identity hash, cache load/save, expected event count."""

    _data_cache = dict()     # hash -> cached attribute dict, shared by all sources of this process

    def __repr__(self):
        return "%s[%s]" % (self.name, getattr(self, 'hash', 'nohashknown'))

    def __init__(self, config, *args, **kwargs):
        c = utils.combine_dicts(_SOURCE_DEFAULTS, config)
        c['cache_attributes'] = list(c['cache_attributes']) + _ALWAYS_CACHED
        c['dont_hash_settings'] = (list(c['dont_hash_settings']) + _NEVER_HASHED
                                   + list(c.pop('extra_dont_hash_settings')))
        self.name = c.pop('name')

        if hasattr(self, 'events_per_day'):
            raise ValueError("events_per_day defaults should be set via config!")
        self.events_per_day = c['events_per_day']
        self.fraction_in_range = c['fraction_in_range']
        self.pdf_has_been_computed = False

        if 'hash' in c:
            self.hash = c['hash']
        else:
            self.hash = c['hash'] = utils.deterministic_hash(
                utils.combine_dicts(c, exclude=c['dont_hash_settings']))

        os.makedirs(c['cache_dir'], exist_ok=True)
        self._cache_filename = os.path.join(c['cache_dir'], self.hash)

        self.from_cache = (not c['force_recalculation']) and os.path.exists(self._cache_filename)
        if self.from_cache:
            if self.hash not in Source._data_cache:
                Source._data_cache[self.hash] = utils.read_pickle(self._cache_filename)
            for key, value in Source._data_cache[self.hash].items():
                if key not in c['cache_attributes']:
                    raise ValueError("%s found in cached file, but you only wanted %s from cache. Old cache?"
                                     % (key, c['cache_attributes']))
                setattr(self, key, value)

        self.config = read_files_in(c, config['data_dirs'])

        if self.from_cache:
            assert self.pdf_has_been_computed
        elif self.config['delay_pdf_computation']:
            self.prepare_task()
        else:
            self.compute_pdf()

    # -- cold path --------------------------------------------------------------------------------
    def compute_pdf(self):
        """Subclasses build their pdf first and call this last: marks the pdf computed and caches it."""
        if self.pdf_has_been_computed:
            raise RuntimeError("compute_pdf called twice on a source!")
        self.pdf_has_been_computed = True
        self.save_to_cache()

    def save_to_cache(self):
        if not self.from_cache and not self.config['never_save_to_cache']:
            utils.save_pickle({k: getattr(self, k) for k in self.config['cache_attributes']},
                              self._cache_filename)
        return self._cache_filename

    def prepare_task(self):
        utils.save_pickle((self.__class__, self.config), os.path.join(self.config['task_dir'], self.hash))

    # -- interface used by the likelihood ---------------------------------------------------------
    def pdf(self, *coordinate_arrays):
        raise NotImplementedError

    def get_pmf_grid(self, *args):
        """(pmf per analysis-space bin, calibration events per bin or inf)."""
        raise NotImplementedError

    def simulate(self, n_events):
        raise NotImplementedError

    @property
    def expected_events(self):
        return (self.events_per_day * self.config['livetime_days']
                * self.fraction_in_range * self.config['rate_multiplier'])


class HistogramPdfSource(Source):
    """A source whose pdf is a density histogram over the analysis space."""
    _pdf_histogram = None
    _bin_volumes = None
    _n_events_histogram = None

    def __init__(self, config, *args, **kwargs):
        config = utils.combine_dicts(dict(pdf_sampling_multiplier=1, pdf_interpolation_method='linear'), config)
        config['cache_attributes'] = list(config.get('cache_attributes', [])) + \
            ['_pdf_histogram', '_n_events_histogram', '_bin_volumes']
        Source.__init__(self, config, *args, **kwargs)

    def build_histogram(self):
        """Set _pdf_histogram (density), _n_events_histogram and _bin_volumes."""
        raise NotImplementedError

    def compute_pdf(self):
        self.build_histogram()
        Source.compute_pdf(self)

    def template(self):
        """(density array [*bins], list of bin-edge arrays, lookup method) for the device gather."""
        if not self.pdf_has_been_computed:
            raise PDFNotComputedException("%s: Attempt to call a PDF that has not been computed" % self)
        method = self.config['pdf_interpolation_method']
        if method not in ('linear', 'piecewise'):
            raise NotImplementedError("PDF Interpolation method %s not implemented" % method)
        return self._pdf_histogram.histogram, self._pdf_histogram.bin_edges, method

    def pdf(self, *coordinate_arrays):
        """pdf values at the given coordinates, evaluated by the K3 device kernel."""
        from . import device_ops
        hist, edges, method = self.template()
        return device_ops.hist_lookup(hist[np.newaxis], edges, coordinate_arrays, method)[0]

    def simulate(self, n_events):
        if not self.pdf_has_been_computed:
            raise PDFNotComputedException(
                "%s: Attempt to simulate events from a PDF that has not been computed" % self)
        positions = (self._pdf_histogram * self._bin_volumes).get_random(n_events)
        space = self.config['analysis_space']
        d = np.zeros(n_events, dtype=[('source', int)] + [(dim[0], float) for dim in space])
        for i, dim in enumerate(space):
            d[dim[0]] = positions[:, i]
        return d

    def get_pmf_grid(self):
        return self._pdf_histogram.histogram * self._bin_volumes, self._n_events_histogram.histogram


class DensityEstimatingSource(HistogramPdfSource):
    """Estimates its density histogram from events returned by get_events_for_density_estimate."""

    def __init__(self, config, *args, **kwargs):
        config = utils.combine_dicts(dict(n_events_for_pdf=1e6), config)
        config['cache_attributes'] = list(config.get('cache_attributes', []))
        HistogramPdfSource.__init__(self, config, *args, **kwargs)

    def build_histogram(self):
        space = self.config['analysis_space']
        names, bins = zip(*space)
        counts = Histdd(bins=bins, axis_names=names)

        provider = self.get_events_for_density_estimate
        batches = provider() if inspect.isgeneratorfunction(provider) else [provider()]
        n_simulated = 0
        # Template construction is model building (prepare()), not the likelihood path: with a CUDA device the 1e6+
        # samples per anchor model and source are binned by bi_histogramdd (np.histogramdd semantics, bit-identical
        # counts), without one by NumPy -- host-only model building stays possible.
        import torch
        on_device = torch.cuda.is_available()
        if on_device:
            from . import device_ops
        for events, n in batches:
            n_simulated += n
            coords = utils._events_to_analysis_dimensions(events, space)
            if on_device and len(events):
                counts.histogram = counts.histogram + device_ops.histogramdd(bins, coords)
            else:
                counts.add(*coords)

        self.fraction_in_range = counts.n / n_simulated
        # density = counts / (events in range) / bin volume
        self._bin_volumes = counts.bin_volumes()
        density = counts.similar_blank_hist()
        density.histogram = counts.histogram.astype(float) / counts.n
        density.histogram /= self._bin_volumes
        self._pdf_histogram = density
        self._n_events_histogram = counts
        return counts

    def get_events_for_density_estimate(self):
        """Return (events, number simulated) or yield such pairs in batches."""
        raise NotImplementedError


class MonteCarloSource(DensityEstimatingSource):
    """Density-estimating source that draws its sample from its own simulate()."""

    def __init__(self, config, *args, **kwargs):
        config = utils.combine_dicts(dict(n_events_for_pdf=1e6, pdf_sampling_multiplier=1,
                                          pdf_sampling_batch_size=1e6), config)
        config['dont_hash_settings'] = list(config.get('dont_hash_settings', [])) + ['pdf_sampling_batch_size']
        DensityEstimatingSource.__init__(self, config, *args, **kwargs)

    def get_events_for_density_estimate(self):
        wanted = self.config['n_events_for_pdf'] * self.config['pdf_sampling_multiplier']
        batch = self.config['pdf_sampling_batch_size']
        if wanted <= batch:
            batch = wanted
        for _ in range(int(wanted // batch)):
            yield self.simulate(n_events=batch), batch
