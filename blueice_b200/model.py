"""Model: a set of sources sharing one config (host-side mirror of blueice/model.py:9-144)."""
import numpy as np

from . import utils

__all__ = ['Model']

_MODEL_DEFAULTS = dict(livetime_days=1, data_dirs=1,
                       nohash_settings=['data_dirs', 'pdf_sampling_batch_size', 'force_recalculation'])
_NOT_FOR_SOURCES = ['sources', 'default_source_class', 'class']
_RATE_SUFFIX = '_rate_multiplier'


class Model(object):
    """Collects the Sources that do the actual work; used for dataset simulation and analysis."""

    def __init__(self, config, **kwargs):
        self.config = utils.combine_dicts(_MODEL_DEFAULTS, config, kwargs, deep_copy=True)
        if 'rate_multiplier' in self.config:
            raise ValueError("Don't put a setting named rate_multiplier in the model config please...")

        self.sources = []
        for entry in self.config['sources']:
            source_class = entry['class'] if 'class' in entry else self.config['default_source_class']
            # each source sees the whole model config plus its own entry (model.py:34-36)
            conf = utils.combine_dicts(self.config, entry, exclude=_NOT_FOR_SOURCES)
            name = conf.get('name', 'WHAAAAAA_YOUDIDNOTNAMEYOURSOURCETHIS')
            own_multiplier = conf.get(name + _RATE_SUFFIX, 1)
            conf = {k: v for k, v in conf.items() if not k.endswith(_RATE_SUFFIX)}
            conf['rate_multiplier'] = own_multiplier
            self.sources.append(source_class(conf))
        del self.config['sources']

    # -- source access ----------------------------------------------------------------------------
    def get_source_i(self, source_id):
        if isinstance(source_id, (int, float)):
            return int(source_id)
        for i, s in enumerate(self.sources):
            if source_id in s.name:
                return i
        raise ValueError("Unknown source %s" % source_id)

    def get_source(self, source_id):
        return self.sources[self.get_source_i(source_id)]

    # -- datasets ---------------------------------------------------------------------------------
    def range_cut(self, d):
        """Events of d inside the (closed) bounds of the analysis space."""
        keep = np.ones(len(d), dtype=bool)
        for name, edges in self.config['analysis_space']:
            keep &= (d[name] >= edges[0]) & (d[name] <= edges[-1])
        return d[keep]

    def simulate(self, rate_multipliers=None, livetime_days=None):
        """Toy dataset: Poisson number of events per source, drawn with Source.simulate."""
        rate_multipliers = rate_multipliers or {}
        parts = []
        for i, source in enumerate(self.sources):
            mu = self.expected_events(source) * rate_multipliers.get(source.name, 1) / source.fraction_in_range
            if livetime_days is not None:
                mu *= livetime_days / self.config['livetime_days']
            events = source.simulate(np.random.poisson(mu))
            events['source'] = i
            parts.append(events)
        return self.range_cut(np.concatenate(parts))

    def simulate_toys(self, n_toys, rate_multipliers=None, livetime_days=None, seed=0, first_toy=0, mus=None):
        """n_toys toy datasets generated on the device (blueice_b200.toys): what n_toys calls of simulate() give,
        as one device-resident ToyData for UnbinnedLogLikelihood.set_toy_data.  Not in the reference API."""
        from .toys import simulate_toys
        return simulate_toys(self, n_toys, rate_multipliers, livetime_days, seed, first_toy, mus)

    def to_analysis_dimensions(self, d):
        """List of coordinate arrays of the events, one per analysis dimension."""
        return utils._events_to_analysis_dimensions(d, self.config['analysis_space'])

    # -- quantities the likelihood interpolates ---------------------------------------------------
    def score_events(self, d):
        """(n_sources, n_events) pdf values (model.py:97-99)."""
        coords = self.to_analysis_dimensions(d)
        return np.vstack([s.pdf(*coords) for s in self.sources])

    def pmf_grids(self):
        """((n_sources, *bins) pmf grids, (n_sources, *bins) calibration events per bin) (model.py:101-104)."""
        grids = [s.get_pmf_grid() for s in self.sources]
        return np.stack([g[0] for g in grids]), np.stack([g[1] for g in grids])

    def expected_events(self, s=None):
        """Expected events in range for source s, or the array over all sources.

        Always float64: the reference returns an int64 array for all-integer configs, which makes
        `mus[i] *= mult` truncate -- a deliberate deviation."""
        if s is None:
            return np.array([self.expected_events(src) for src in self.sources], dtype=np.float64)
        return s.expected_events

    def show(self, d, ax=None, dims=None, **kwargs):
        """Scatter plot of the events of d per source (needs matplotlib)."""
        import matplotlib.pyplot as plt
        kwargs.setdefault('s', 5)
        names, bins = zip(*self.config['analysis_space'])
        if dims is None:
            dims = (0,) if len(bins) == 1 else (0, 1)
        ax = ax or plt.gca()
        for i, src in enumerate(self.sources):
            coords = self.to_analysis_dimensions(d[d['source'] == i])
            y = coords[dims[1]] if len(dims) > 1 else np.zeros(len(coords[dims[0]]))
            ax.scatter(coords[dims[0]], y, color=src.config['color'], label=src.config['label'], **kwargs)
        ax.set_xlabel(names[dims[0]])
        ax.set_xlim(bins[dims[0]][0], bins[dims[0]][-1])
        if len(dims) > 1:
            ax.set_ylabel(names[dims[1]])
            ax.set_ylim(bins[dims[1]][0], bins[dims[1]][-1])
