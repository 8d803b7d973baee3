"""Log-likelihood constructors with the reference's API, evaluated by sm_100a kernels.

Drop-in surface (blueice/likelihood.py):
    UnbinnedLogLikelihood / BinnedLogLikelihood(pdf_base_config, likelihood_config=None, **overrides)
    add_rate_parameter, add_shape_parameter, add_rate_uncertainty, add_shape_uncertainty,
    prepare(), set_data(d), ll(**params), get_bounds(), the inference methods, and the NEW
    ll.batch(params_array, names=None, livetime_days=None) -> float64[P]
    (row p is bit-identical to ll(**dict(zip(names, params_array[p])))).

What runs where
    host (this file): parameter validation, defaults, bounds -> -inf, priors (arbitrary Python
        callables), livetime/efficiency factors, 'error'-mode exceptions -- the control flow of
        LogLikelihoodBase.__call__ (likelihood.py:318-427), vectorised over the batch.
    device (engine.py -> csrc/): anchor-grid morphing of mus and of the per-event pdf / per-bin pmf
        tensors, rate scaling, unphysical-rate test, mixture density, log, reduction; binned Poisson
        with Beeston-Barlow; template lookup and event binning in set_data.
"""
import os
from collections import OrderedDict
from copy import deepcopy
from functools import wraps

import numpy as np
from scipy import stats

from . import _cabi
from . import inference
from .engine import BinnedEngine, MorphGrid, SourcewiseUnbinnedEngine, TemplateUnbinnedEngine, UnbinnedEngine
from .exceptions import InvalidParameter, InvalidParameterSpecification, NotPreparedException
from .hist import Histdd
from .model import Model
from .pdf_morphers import MORPHERS
from .source import HistogramPdfSource
from .utils import combine_dicts, inherit_docstring_from

__all__ = ['LogLikelihoodBase', 'BinnedLogLikelihood', 'UnbinnedLogLikelihood', 'LogLikelihoodSum',
           'LogAncillaryLikelihood', 'LogLikelihoodReParam', 'extended_loglikelihood',
           'beeston_barlow_root1', 'beeston_barlow_root2', 'beeston_barlow_roots']

_RATE_SUFFIX = '_rate_multiplier'
_NEG_INF = -float('inf')


def _needs_preparation(method):
    """Auto-prepare when there are no shape parameters, otherwise insist on prepare() (likelihood.py:30-41)."""
    @wraps(method)
    def wrapped(self, *args, **kwargs):
        if not self.is_prepared:
            if len(self.shape_parameters):
                raise NotPreparedException("%s requires you to first prepare the likelihood function using prepare()"
                                           % method.__name__)
            self.prepare()
        return method(self, *args, **kwargs)
    return wrapped


def _needs_data(method):
    @wraps(method)
    def wrapped(self, *args, **kwargs):
        if not self.is_data_set:
            raise NotPreparedException("%s requires you to first set the data using set_data()" % method.__name__)
        return method(self, *args, **kwargs)
    return wrapped


def _is_number(x):
    # the reference's test (likelihood.py:285,465,472): np.float64 passes, np.int64 does not
    return isinstance(x, (float, int))


class LogLikelihoodBase(object):
    """Log likelihood function with several rate and/or shape parameters.

    likelihood_config options: morpher, morpher_config, unphysical_behaviour ('error' or None),
    outlier_likelihood (default 1e-12), model_statistical_uncertainty_handling, bb_single_source.
    """

    def __init__(self, pdf_base_config, likelihood_config=None, **kwargs):
        self.pdf_base_config = combine_dicts(pdf_base_config, kwargs, deep_copy=True)
        self.config = {} if likelihood_config is None else likelihood_config   # kept by reference (likelihood.py:74)
        self.config.setdefault('morpher', 'GridInterpolator')
        self.source_wise_interpolation = self.pdf_base_config.get('source_wise_interpolation', False)

        self.base_model = Model(self.pdf_base_config)
        sources = self.base_model.sources
        self.source_name_list = [s.name for s in sources]
        self.source_allowed_negative = [s.config.get("allow_negative", False) for s in sources]
        self.source_apply_efficiency = np.array([s.config.get("apply_efficiency", False) for s in sources])
        self.source_efficiency_names = np.array([s.config.get("efficiency_name", "efficiency") for s in sources])

        self.rate_parameters = OrderedDict()     # source name -> log prior (or None)
        self.shape_parameters = OrderedDict()    # setting name -> (anchors {z: setting}, log prior, base z)
        self.is_prepared = False
        self.is_data_set = False
        self._has_non_numeric = False

        self.ps = None                           # no shape parameters: pdf values / pmf grid of the base model
        self.anchor_models = OrderedDict()       # anchor z tuple -> Model
        self.anchor_sources = OrderedDict()
        self.morpher = None
        self.mus_interpolator = None
        self.ps_interpolator = None
        self.n_model_events_interpolator = lambda x: None
        self.n_model_events = None
        self._engine = None                      # device engine, built by set_data (unbinned) / prepare (binned)
        self._scalar_plans = None                # lean single-evaluation closures per keyword-name tuple (unbinned)
        self._grid = None
        self._mus_anchor = None

    # ------------------------------------------------------------------------------------------
    # parameter bookkeeping
    # ------------------------------------------------------------------------------------------
    def add_rate_parameter(self, source_name, log_prior=None):
        """Add <source_name>_rate_multiplier, multiplying the expected events of that source."""
        self.rate_parameters[source_name] = log_prior

    def add_shape_parameter(self, setting_name, anchors, log_prior=None, base_value=None):
        """Add a shape parameter: `anchors` is a sequence of numeric setting values, or a dict
        {representative number z: setting} for non-numeric settings (then base_value is required)."""
        numeric = _is_number(self.pdf_base_config.get(setting_name))
        if not isinstance(anchors, dict):
            if not numeric:
                raise InvalidParameterSpecification("When specifying anchors only by setting values, "
                                                    "base setting must have a numerical default.")
            anchors = {z: z for z in anchors}
        if not numeric:
            self._has_non_numeric = True
            if base_value is None:
                raise InvalidParameterSpecification("For non-numeric settings, you must specify what number "
                                                    "will represent the default value (the base model setting)")
        elif base_value is not None:
            raise InvalidParameterSpecification("For numeric settings, base_value is an unnecessary argument.")
        self.shape_parameters[setting_name] = (anchors, log_prior, base_value)

    def add_rate_uncertainty(self, source_name, fractional_uncertainty):
        """Rate parameter with a Gaussian prior of width fractional_uncertainty around 1."""
        self.add_rate_parameter(source_name, log_prior=stats.norm(1, fractional_uncertainty).logpdf)

    def add_shape_uncertainty(self, setting_name, fractional_uncertainty, anchor_zs=(-2, -1, 0, 1, 2), base_value=None):
        """Shape parameter with a Gaussian prior around its default value (likelihood.py:492-504)."""
        self.add_shape_parameter(setting_name, anchor_zs, base_value=base_value)
        anchors, _, base_value = self.shape_parameters[setting_name]
        prior = stats.norm(base_value, base_value * fractional_uncertainty).logpdf
        self.shape_parameters[setting_name] = (anchors, prior, base_value)

    def get_bounds(self, parameter_name=None):
        """Bounds of one parameter, or the list of bounds of all shape parameters."""
        if parameter_name is None:
            return [self.get_bounds(p) for p in self.shape_parameters.keys()]
        if parameter_name in self.shape_parameters:
            zs = list(self.shape_parameters[parameter_name][0].keys())
            return min(zs), max(zs)
        if parameter_name.endswith(_RATE_SUFFIX):
            for name, negative_ok in zip(self.source_name_list, self.source_allowed_negative):
                if parameter_name.startswith(name) and negative_ok == True:    # noqa: E712 (reference semantics)
                    return float('-inf'), float('inf')
            return 0, float('inf')
        raise InvalidParameter("Non-existing parameter %s" % parameter_name)

    @property
    def source_shape_parameters(self):
        """source name -> OrderedDict of the shape parameters that source depends on (likelihood.py:113-130)."""
        out = OrderedDict()
        for name, source, use_eff, eff_name in zip(self.source_name_list, self.base_model.sources,
                                                   self.source_apply_efficiency, self.source_efficiency_names):
            ignored = set(source.config['dont_hash_settings'])
            if use_eff:
                ignored.discard(eff_name)
            mine = OrderedDict((k, v) for k, v in self.shape_parameters.items() if k not in ignored)
            if mine:
                out[name] = mine
        return out

    def _default_z(self, setting_name, base_value):
        base_setting = self.pdf_base_config.get(setting_name)
        if _is_number(base_setting):
            assert base_value is None
            return base_setting
        return base_value

    def _kwargs_to_settings(self, **kwargs):
        """(rate multipliers per source, {shape setting: z}) from call keywords (likelihood.py:443-481)."""
        for key in kwargs:
            if key in self.shape_parameters:
                continue
            if key.endswith(_RATE_SUFFIX) and key[:-len(_RATE_SUFFIX)] in self.source_name_list:
                continue
            raise InvalidParameter("%s is not a known shape or rate parameter!" % key)
        settings = dict()
        for name, (_, _, base_value) in self.shape_parameters.items():
            z = kwargs.get(name)
            if z is None:
                z = self._default_z(name, base_value)
            if not _is_number(z):
                raise ValueError("Arguments to likelihood function must be numeric, not %s" % type(z))
            settings[name] = z
        multipliers = [kwargs.get(name + _RATE_SUFFIX, 1) for name in self.source_name_list]
        return multipliers, settings

    # ------------------------------------------------------------------------------------------
    # prepare: anchor models (cold path, host)
    # ------------------------------------------------------------------------------------------
    def prepare(self, n_cores=1, ipp_client=None):
        """Compute the model at every anchor point of the shape-parameter grid (likelihood.py:147-254).

        Model construction is user Python (simulators, files) and stays on the host, serially;
        n_cores / ipp_client are accepted for signature compatibility (the reference's process-pool /
        ipyparallel farm, blueice/parallel.py, is out of scope -- SURVEY.md section 2 row 12)."""
        if len(self.shape_parameters):
            if self.source_wise_interpolation:
                self._prepare_source_wise()
            else:
                self.morpher = MORPHERS[self.config['morpher']](self.config.get('morpher_config', {}),
                                                                self.shape_parameters)
                zs_list = self.morpher.get_anchor_points(bounds=self.get_bounds())
                for zs in zs_list:
                    config = deepcopy(self.pdf_base_config)
                    for i, (setting_name, (anchors, _, _)) in enumerate(self.shape_parameters.items()):
                        if zs[i] is not None:
                            config[setting_name] = anchors[zs[i]]
                    self.anchor_models[tuple(zs)] = Model(config)
                self._grid = MorphGrid(self.morpher.anchor_z_arrays)
                self._mus_anchor = self.morpher.anchor_tensor(lambda m: m.expected_events(),
                                                              [len(self.source_name_list)], self.anchor_models)
                self.mus_interpolator = _LazyInterpolator(self.morpher, self._mus_anchor,
                                                          [len(self.source_name_list)])
        else:
            self._grid = MorphGrid([])
            self._mus_anchor = np.asarray(self.base_model.expected_events(), dtype=np.float64)[np.newaxis, :]
        self.is_data_set = False
        self.is_prepared = True

    # -- source-wise interpolation (likelihood.py:113-145,152-169,210-240) ---------------------------
    @property
    def source_shape_parameters(self):
        """source name -> OrderedDict of the shape parameters that source depends on (likelihood.py:113-131)."""
        result = OrderedDict()
        for sn, source, apply_eff, eff_name in zip(self.source_name_list, self.base_model.sources,
                                                   self.source_apply_efficiency, self.source_efficiency_names):
            ignore = set(source.config['dont_hash_settings'])
            if apply_eff:
                ignore.discard(eff_name)      # not hashed, but it must reach the morpher
            own = OrderedDict((k, v) for k, v in self.shape_parameters.items() if k not in ignore)
            if own:
                result[sn] = own
        return result

    def _get_shape_indices(self, source_name):
        """Indices (into self.shape_parameters) of the shape parameters used by the source."""
        keys = self.source_shape_parameters[source_name].keys()
        return [i for i, k in enumerate(self.shape_parameters.keys()) if k in keys]

    def _get_model_anchor(self, anchor, source_name):
        """Anchor of the full model for a source's own anchor; unused parameters are None."""
        model_anchor = [None] * len(self.shape_parameters)
        for i, idx in enumerate(self._get_shape_indices(source_name)):
            model_anchor[idx] = anchor[i]
        return tuple(model_anchor)

    def _prepare_source_wise(self):
        """One morpher per source over the parameters it depends on; models only at the union of the
        sources' anchors (likelihood.py:152-169,210-240).  The full grid is kept for the bounds test and
        for bucketing points; the device tables are per (source, sub-anchor) rows."""
        ssp = self.source_shape_parameters
        self.source_morphers = OrderedDict(
            (sn, MORPHERS[self.config['morpher']](self.config.get('morpher_config', {}), sp)) for sn, sp in ssp.items())
        zs_list = []
        for sn, morpher in self.source_morphers.items():
            for anchor in morpher.get_anchor_points(bounds=None):
                zs = self._get_model_anchor(anchor, sn)
                if zs not in zs_list:
                    zs_list.append(zs)
        models = []
        for zs in zs_list:
            config = deepcopy(self.pdf_base_config)
            for i, (setting_name, (anchors, _, _)) in enumerate(self.shape_parameters.items()):
                if zs[i] is not None:
                    config[setting_name] = anchors[zs[i]]
            models.append(Model(config))
        self.anchor_sources = OrderedDict()
        for sn, morpher in self.source_morphers.items():
            source_index = self.source_name_list.index(sn)
            self.anchor_sources[sn] = OrderedDict()
            for anchor in morpher.get_anchor_points(bounds=None):
                model = models[zs_list.index(self._get_model_anchor(anchor, sn))]
                self.anchor_sources[sn][anchor] = model.sources[source_index]
        full = MORPHERS[self.config['morpher']](self.config.get('morpher_config', {}), self.shape_parameters)
        self._grid = MorphGrid(full.anchor_z_arrays)
        self._sw_source_dims = [self._get_shape_indices(sn) if sn in ssp else [] for sn in self.source_name_list]
        self._sw_row_sources = []             # (source index, Source) per row, sources concatenated, C order
        for s, (sn, base_source) in enumerate(zip(self.source_name_list, self.base_model.sources)):
            if sn in self.source_morphers:
                for anchor in self.source_morphers[sn].get_anchor_points(bounds=None):
                    self._sw_row_sources.append((s, self.anchor_sources[sn][anchor]))
            else:
                self._sw_row_sources.append((s, base_source))
        self._sw_mus_rows = np.array([src.expected_events for _, src in self._sw_row_sources], dtype=np.float64)
        self._mus_anchor = None

        def mus_interpolator(zs):
            from .engine import SourcewiseUnbinnedEngine
            engine = SourcewiseUnbinnedEngine(self._grid, self._sw_source_dims, self._sw_mus_rows,
                                              allow_negative=self.source_allowed_negative)
            engine.allocate_ps_anchor(0)
            z = np.asarray(zs, dtype=np.float64).reshape(1, -1)
            return engine.point_setup_host(z, np.ones((1, len(self.source_name_list))))["mus"][0]
        self.mus_interpolator = mus_interpolator

    @_needs_preparation
    def set_data(self, d):
        """Prepare dataset d (anything indexable by analysis-dimension name) for evaluation."""
        self._data = d
        self.is_data_set = True

    # ------------------------------------------------------------------------------------------
    # evaluation
    # ------------------------------------------------------------------------------------------
    def parameter_names(self):
        """Default column order of batch(): rate parameters (insertion order), then shape parameters
        (the order make_objective uses, inference.py:79-102)."""
        return [s + _RATE_SUFFIX for s in self.rate_parameters.keys()] + list(self.shape_parameters.keys())

    @_needs_data
    def __call__(self, livetime_days=None, compute_pdf=False, full_output=False, **kwargs):
        """Evaluate the log likelihood; parameters not passed take their base values.

        :param livetime_days: exposure to evaluate at (scales the rates of all sources)
        :param compute_pdf: build a new model at the requested parameters instead of interpolating
        :param full_output: also return the adjusted mus and the pdf values / pmf grid
        """
        multipliers, settings = self._kwargs_to_settings(**kwargs)
        if len(self.shape_parameters) and compute_pdf:
            if self._has_non_numeric:
                raise NotImplementedError("compute_pdf only works for numerical values")
            return self._call_computed_pdf(multipliers, settings, livetime_days, full_output, kwargs)
        zs = np.array([[settings[name] for name in self.shape_parameters]], dtype=np.float64).reshape(1, -1)
        mult = np.array([multipliers], dtype=np.float64)
        result = self._evaluate_rows(self._engine, zs, mult, livetime_days, scalar=True)
        if full_output and result != _NEG_INF:
            scale, _ = self._livetime_scale(livetime_days)
            eff = self._efficiencies(zs)
            return self._full_output(self._engine, zs[0], mult[0], scale, None if eff is None else eff[0], result)
        return result

    @_needs_data
    def batch(self, params, names=None, livetime_days=None):
        """Evaluate the log likelihood at P parameter points with one device pass.

        :param params: array [P, k]; column j holds parameter names[j]
        :param names: parameter names of the columns; default self.parameter_names()
        :returns: float64 [P]; row p equals self(**dict(zip(names, params[p])), livetime_days=...)
        """
        zs, mult = self._rows_from_params(params, names)
        return self._evaluate_rows(self._engine, zs, mult, livetime_days, scalar=False)

    def _rows_from_params(self, params, names):
        """(zs [P, D], rate multipliers [P, S]) from a parameter table, defaults filled in like __call__."""
        names = self.parameter_names() if names is None else list(names)
        params = np.asarray(params, dtype=np.float64)
        if params.ndim == 1:
            params = params.reshape(-1, max(len(names), 1))
        if params.ndim != 2 or params.shape[1] != len(names):
            raise ValueError("params must have shape [n_points, %d]" % len(names))
        P = params.shape[0]
        # the column map of a name list is resolved (and validated like __call__ validates its kwargs) once per list
        key = (tuple(names), len(self.rate_parameters),
               tuple(repr(self._default_z(n, bv)) for n, (_, _, bv) in self.shape_parameters.items()))
        cache = self.__dict__.setdefault('_column_maps', {})
        cmap = cache.get(key)
        if cmap is None:
            self._kwargs_to_settings(**{n: 1.0 for n in names})      # same name validation as __call__
            defaults_mult, defaults_settings = self._kwargs_to_settings()
            z_cols = [(names.index(name) if name in names else -1, defaults_settings[name]) for name in self.shape_parameters]
            m_cols = [(names.index(s + _RATE_SUFFIX) if s + _RATE_SUFFIX in names else -1, defaults_mult[j])
                      for j, s in enumerate(self.source_name_list)]
            numeric = all(c >= 0 or _is_number(v) for c, v in z_cols)
            cmap = (z_cols, m_cols)
            if numeric and len(cache) < 64:
                cache[key] = cmap
        z_cols, m_cols = cmap
        zs = np.empty((P, len(z_cols)), dtype=np.float64)
        for j, (c, default) in enumerate(z_cols):
            zs[:, j] = params[:, c] if c >= 0 else default
        mult = np.empty((P, len(m_cols)), dtype=np.float64)
        for j, (c, default) in enumerate(m_cols):
            mult[:, j] = params[:, c] if c >= 0 else default
        return zs, mult

    def _livetime_scale(self, livetime_days):
        """Factor applied to all mus (likelihood.py:374-382); None when no scaling happens."""
        if livetime_days is None:
            return None, False
        if 'livetime_days' not in self.pdf_base_config:
            raise ValueError("Cannot scale live-time, base value absent")
        base = self.pdf_base_config['livetime_days']
        if base == 0:
            if livetime_days != 0:
                raise ValueError("Cannot scale from 0 to non-0 livetime")
            return None, True
        return livetime_days / base, False

    def _prior_sum(self, zs, mult):
        """Sum of log priors per point, accumulated in the reference's order (likelihood.py:349-350,369-371)."""
        P = len(mult)
        total = np.zeros(P)
        columns = [(prior, zs[:, j]) for j, (_, prior, _) in enumerate(self.shape_parameters.values())]
        columns += [(self.rate_parameters.get(name), mult[:, j]) for j, name in enumerate(self.source_name_list)]
        for prior, values in columns:
            if prior is None:
                continue
            vec = None
            if P > 1:
                try:
                    vec = np.asarray(prior(values), dtype=np.float64)
                    if vec.shape != (P,):
                        vec = None
                except Exception:
                    vec = None
            if vec is None:
                vec = np.array([prior(float(v)) for v in values], dtype=np.float64)
            total = total + vec
        return total

    def _efficiencies(self, zs):
        """[P, S] efficiency multipliers, or None when no source applies one (likelihood.py:385-393)."""
        if True not in self.source_apply_efficiency:
            return None
        names = list(self.shape_parameters.keys())
        eff = np.ones((len(zs), len(self.source_name_list)))
        for s, (use, eff_name) in enumerate(zip(self.source_apply_efficiency, self.source_efficiency_names)):
            if use and eff_name in names:
                eff[:, s] = zs[:, names.index(eff_name)]
        return eff

    def _evaluate_rows(self, engine, zs, mult, livetime_days, scalar):
        scale, zero_base = self._livetime_scale(livetime_days)
        P = len(mult)
        scale_arr = None if scale is None else np.full(P, scale, dtype=np.float64)
        eff = self._efficiencies(zs)
        priors = self._prior_sum(zs, mult)
        ll, status = self._device_loglikelihood(engine, zs, mult, scale_arr, eff)
        unphysical = (status & _cabi.POINT_UNPHYSICAL) != 0
        in_range = (status & _cabi.POINT_OUT_OF_RANGE) == 0
        if zero_base:
            mus = engine.point_setup_host(zs, mult, scale_arr, eff)["mus"]
            assert np.all(mus[in_range] == 0), "Got non-0 mus with 0 livetime?!"
        if self.config.get('unphysical_behaviour') == 'error' and np.any(unphysical & in_range):
            p = int(np.flatnonzero(unphysical & in_range)[0])
            mus = engine.point_setup_host(zs[p:p + 1], mult[p:p + 1],
                                          None if scale_arr is None else scale_arr[p:p + 1],
                                          None if eff is None else eff[p:p + 1])["mus"][0]
            raise ValueError("Unphysical rates: %s" % str(mus))
        result = np.where(status != 0, _NEG_INF, priors + ll)
        if scalar:
            return result[0] if status[0] == 0 else _NEG_INF
        return result

    def _device_loglikelihood(self, engine, zs, mult, scale, eff):
        """(logL without priors [P], status [P]) from the device engine."""
        raise NotImplementedError

    def _full_output(self, engine, z_row, mult_row, scale, eff_row, result):
        raise NotImplementedError

    def _evaluate_fixed_model(self, engine, multipliers, settings, livetime_days, full_output):
        """compute_pdf=True: a freshly computed model instead of interpolation, then rate priors, scaling
        and checks as usual (likelihood.py:331-335,365-427; the reference skips shape priors here)."""
        mult = np.array([multipliers], dtype=np.float64)
        zs_named = np.array([[settings[name] for name in self.shape_parameters]], dtype=np.float64).reshape(1, -1)
        scale, _ = self._livetime_scale(livetime_days)
        scale_arr = None if scale is None else np.full(1, scale)
        eff = self._efficiencies(zs_named)
        prior = 0.
        for j, name in enumerate(self.source_name_list):
            log_prior = self.rate_parameters.get(name)
            if log_prior is not None:
                prior += log_prior(multipliers[j])
        no_zs = np.zeros((1, 0))
        ll, status = self._device_loglikelihood(engine, no_zs, mult, scale_arr, eff)
        if status[0] != 0:
            if self.config.get('unphysical_behaviour') == 'error':
                raise ValueError("Unphysical rates: %s" % str(engine.point_setup_host(no_zs, mult, scale_arr, eff)["mus"][0]))
            return _NEG_INF
        result = prior + ll[0]
        if full_output:
            return self._full_output(engine, no_zs[0], mult[0], scale, None if eff is None else eff[0], result)
        return result

    def _call_computed_pdf(self, multipliers, settings, livetime_days, full_output, kwargs):
        raise NotImplementedError

    def _compute_single_model(self, **kwargs):
        """Model built from the base config with the given shape settings as overrides (likelihood.py:506-512)."""
        _, settings = self._kwargs_to_settings(**kwargs)
        config = combine_dicts(self.pdf_base_config, settings, deep_copy=True)
        config['never_save_to_cache'] = True
        return Model(config, **settings)

    def adjust_expectations(self, mus, ps, n_model_events):
        """Hook of the reference (likelihood.py:429-441); identity unless overridden."""
        return mus, ps

    def _compute_likelihood(self, *args, **kwargs):
        raise NotImplementedError

    def _compute_single_pdf(self, **kwargs):
        raise NotImplementedError


class _LazyInterpolator(object):
    """`mus_interpolator` / `ps_interpolator` attributes of the reference API: a callable zs -> array,
    backed by a device tensor that is only uploaded if somebody actually calls it."""

    def __init__(self, morpher, anchor_tensor, extra_dims):
        self._morpher, self._tensor, self._extra = morpher, anchor_tensor, list(extra_dims)
        self._fn = None

    def __call__(self, zs):
        if self._fn is None:
            from .pdf_morphers import DeviceGridFunction
            self._fn = DeviceGridFunction(self._morpher.anchor_z_arrays, self._tensor, self._extra)
        return self._fn(zs)


class _ToyView(object):
    """Engine facade for _evaluate_rows: point t is evaluated on dataset t."""

    def __init__(self, engine, toy_index=None):
        self.engine = engine
        self.toy_index = toy_index
        self.point_setup_host = engine.point_setup_host

    def evaluate(self, zs, mult, scale=None, eff=None, return_status=False):
        if self.toy_index is not None:
            if len(self.toy_index) != len(mult):
                raise ValueError("toy_index must hold one toy per parameter row")
            return self.engine.evaluate_pairs(self.toy_index, zs, mult, scale, eff, return_status=return_status)
        return self.engine.evaluate_toys(zs, mult, scale, eff, return_status=return_status)


class _BinnedToyView(object):
    """Engine facade for _evaluate_rows: point t is evaluated on the observed histogram of toy t."""

    def __init__(self, ll, engine):
        self.ll, self.engine = ll, engine
        self.point_setup_host = engine.point_setup_host

    def evaluate(self, zs, mult, scale=None, eff=None, return_status=False):
        return self.engine.evaluate_toys(zs, mult, scale, eff, return_status=return_status)


class UnbinnedLogLikelihood(LogLikelihoodBase):

    @inherit_docstring_from(LogLikelihoodBase)
    def set_data(self, d):
        LogLikelihoodBase.set_data(self, d)
        outlier = self.config.get('outlier_likelihood', 1e-12)
        template_mode = self._wants_template_engine(len(d))
        if template_mode:
            engine = self._build_template_engine(template_mode)
            dims = self.base_model.to_analysis_dimensions(d)
            engine.set_datasets(np.asarray([np.asarray(c, dtype=np.float64) for c in dims]).reshape(len(dims), -1))
        elif len(self.shape_parameters) and self.source_wise_interpolation:
            engine = SourcewiseUnbinnedEngine(self._grid, self._sw_source_dims, self._sw_mus_rows,
                                              outlier_likelihood=outlier, allow_negative=self.source_allowed_negative)
            items = [(row, s, source) for row, (s, source) in enumerate(self._sw_row_sources)]
            self._fill_anchor_rows(engine, items, d)
        else:
            engine = UnbinnedEngine(self._grid, self._mus_anchor.reshape(self._grid.n_anchors, -1),
                                    outlier_likelihood=outlier, allow_negative=self.source_allowed_negative)
            self._fill_anchor_rows(engine, self._anchor_items(), d)
        self._engine = engine
        self._scalar_plans = {} if os.environ.get('BI_SCALAR_FAST', '1') != '0' else None
        if len(self.shape_parameters):
            self.ps_interpolator = lambda zs: engine.ps(np.asarray(zs, dtype=np.float64),
                                                        np.ones(engine.n_sources))[1]
        elif not isinstance(engine, TemplateUnbinnedEngine):
            self.ps = engine.ps(np.zeros(0), np.ones(engine.n_sources))[1]

    # ------------------------------------------------------------------------------------------
    # single evaluations: the lean path (a minimiser or an interval search calls ll(**params) thousands of times,
    # inference.py:131-178,332-389; there the Python around the kernels is what costs)
    # ------------------------------------------------------------------------------------------
    def __call__(self, livetime_days=None, compute_pdf=False, full_output=False, **kwargs):
        if not (compute_pdf or full_output) and self.is_data_set and self._scalar_plans is not None:
            key = (tuple(kwargs), livetime_days is None)
            plan = self._scalar_plans.get(key)
            if plan is None:
                plan = self._scalar_plans[key] = self._make_scalar_plan(key[0], livetime_days) or False
            if plan:
                result = plan(kwargs, livetime_days)
                if result is not NotImplemented:
                    return result
        return LogLikelihoodBase.__call__(self, livetime_days=livetime_days, compute_pdf=compute_pdf,
                                          full_output=full_output, **kwargs)

    def batch(self, params, names=None, livetime_days=None):
        plans = self._scalar_plans
        if plans is not None and self.is_data_set:
            params = np.asarray(params, dtype=np.float64)
            if params.ndim == 2 and params.shape[0] > 0:
                pg = getattr(self._engine, 'peer_gather', None)          # sharded evaluations get plans of their own
                key = ('batch', None if names is None else tuple(names), params.shape, livetime_days is None,
                       None if pg is None else (id(pg), self._engine.peer_mode))
                plan = plans.get(key)
                if plan is None:
                    if len(plans) > 64:
                        plans.clear()
                    plan = plans[key] = self._make_batch_plan(names, params.shape, livetime_days) or False
                if plan:
                    result = plan(params, livetime_days)
                    if result is not NotImplemented:
                        return result
        return LogLikelihoodBase.batch(self, params, names, livetime_days=livetime_days)

    def _make_batch_plan(self, names, shape, livetime_days):
        """Closure evaluating ll.batch(params, names) for tables of this shape and column list, or None when the general
        path must be used.  Same defaults, priors, device sequence and result arithmetic as LogLikelihoodBase.batch; the
        columns go straight into the pinned staging buffer and the result is read straight out of the pinned result."""
        engine = self._engine
        if type(engine) is not UnbinnedEngine or not engine.uses_mma() or engine.n_events <= 0:
            return None
        pg = engine.peer_gather
        if True in self.source_apply_efficiency:
            return None
        names = self.parameter_names() if names is None else list(names)
        P, k = shape
        if k != len(names):
            return None
        try:
            self._kwargs_to_settings(**{n: 1.0 for n in names})
            scale, zero_base = self._livetime_scale(livetime_days)
            defaults_mult, defaults_settings = self._kwargs_to_settings()
        except Exception:
            return None
        if zero_base or any(not _is_number(v) for v in defaults_settings.values()):
            return None
        has_scale = scale is not None
        zs_v, mult_v, scale_v, run = engine.batch_runner(P, has_scale)
        z_cols = [(j, names.index(n) if n in names else -1, float(defaults_settings[n]))
                  for j, n in enumerate(self.shape_parameters)]
        m_cols = [(j, names.index(s + _RATE_SUFFIX) if s + _RATE_SUFFIX in names else -1, float(defaults_mult[j]))
                  for j, s in enumerate(self.source_name_list)]
        has_priors = any(p is not None for _, p, _ in self.shape_parameters.values()) or \
            any(p is not None for p in self.rate_parameters.values())
        error_mode = self.config.get('unphysical_behaviour') == 'error'
        base_livetime = self.pdf_base_config.get('livetime_days')

        def plan(params, livetime_days):
            if engine.peer_gather is not pg:
                return NotImplemented
            for j, c, default in z_cols:
                zs_v[:, j] = params[:, c] if c >= 0 else default
            for j, c, default in m_cols:
                mult_v[:, j] = params[:, c] if c >= 0 else default
            if has_scale:
                scale_v[:] = livetime_days / base_livetime
            priors = self._prior_sum(zs_v, mult_v) if has_priors else 0.0
            logl, status = run()
            if status.any():
                if error_mode:
                    return NotImplemented                      # the general path raises the reference's error
                return np.where(status != 0, _NEG_INF, priors + logl)
            return priors + logl
        return plan

    def _make_scalar_plan(self, names, livetime_days):
        """Closure evaluating ll(**kwargs) for this set of keyword names, or None when the general path must be used.
        It reproduces LogLikelihoodBase.__call__ for P = 1 (same defaults, priors in the same order, same device call);
        anything unusual (non-numeric arguments, unknown names, efficiencies, 'error'-mode failures) returns
        NotImplemented so that the general path produces the reference's behaviour."""
        engine = self._engine
        if type(engine) is not UnbinnedEngine or not engine.uses_mma() or engine.n_events <= 0:
            return None
        if True in self.source_apply_efficiency:
            return None
        try:
            self._kwargs_to_settings(**{n: 1.0 for n in names})       # the name validation of __call__
            scale, zero_base = self._livetime_scale(livetime_days)
            defaults_mult, defaults_settings = self._kwargs_to_settings()
        except Exception:
            return None
        if zero_base or any(not _is_number(v) for v in defaults_settings.values()):
            return None
        has_scale = scale is not None
        pin, run = engine.scalar_runner(has_scale)
        shape_names = list(self.shape_parameters.keys())
        D, S = len(shape_names), len(self.source_name_list)
        z_slots = [(j, n, float(defaults_settings[n])) for j, n in enumerate(shape_names)]
        m_slots = [(D + j, s + _RATE_SUFFIX, float(defaults_mult[j])) for j, s in enumerate(self.source_name_list)]
        priors = [(prior, j) for j, (_, prior, _) in enumerate(self.shape_parameters.values()) if prior is not None]
        priors += [(self.rate_parameters.get(s), D + j) for j, s in enumerate(self.source_name_list)
                   if self.rate_parameters.get(s) is not None]
        error_mode = self.config.get('unphysical_behaviour') == 'error'
        base_livetime = self.pdf_base_config.get('livetime_days')
        number = (float, int)
        f64 = np.float64

        def plan(kwargs, livetime_days):
            get = kwargs.get
            for j, name, default in z_slots:
                v = get(name)
                if v is None:
                    v = default
                elif not isinstance(v, number):
                    return NotImplemented
                pin[j] = v
            for j, name, default in m_slots:
                pin[j] = get(name, default)
            if has_scale:
                pin[D + S] = livetime_days / base_livetime
            logl, status = run()
            if status != 0:
                return NotImplemented if error_mode else _NEG_INF
            if priors:
                total = f64(0.0)
                for prior, j in priors:
                    total = total + prior(float(pin[j]))
                return total + logl
            return f64(0.0) + logl
        return plan

    def _anchor_items(self):
        """(anchor index, source index, Source) for every row of the full anchor grid, anchors in C order."""
        if len(self.shape_parameters):
            models = [self.anchor_models[tuple(zs)] for _, zs in self.morpher._anchor_grid_iterator()]
        else:
            models = [self.base_model]
        return [(g, s, source) for g, model in enumerate(models) for s, source in enumerate(model.sources)]

    # dense anchor tensors above this many bytes switch `unbinned_engine: 'auto'` to the template-space engine
    _TEMPLATE_ENGINE_BYTES = 16 << 30

    def _wants_template_engine(self, n_events):
        """Template-engine mode for this dataset, or None for the dense anchor tensor.

        likelihood_config['unbinned_engine']: 'anchor' (dense per-event anchor tensor, K3 + K2);
        'template' (K5: per-event values looked up on the fly, bit-identical to 'anchor'); 'mixture' (K5b: templates
        morphed per point, events looked up in the mixture template -- one lookup per point-event, HBM-bound; equal
        to 'anchor' to ~1e-13 relative); 'auto' (default: the dense tensor unless it would exceed
        _TEMPLATE_ENGINE_BYTES, then 'mixture', or 'template' if some template holds a non-finite value)."""
        kind = self.config.get('unbinned_engine', 'auto')
        if kind == 'anchor':
            return None
        if kind not in ('template', 'mixture', 'auto'):
            raise ValueError("unbinned_engine must be 'anchor', 'template', 'mixture' or 'auto'")
        if kind in ('template', 'mixture'):
            reason = self._template_engine_obstacle()
            if reason:
                raise NotImplementedError("unbinned_engine=%r: %s" % (kind, reason))
            return 'exact' if kind == 'template' else 'mixture'
        dense = 8 * self._grid.n_anchors * len(self.source_name_list) * n_events
        if dense <= self._TEMPLATE_ENGINE_BYTES or self._template_engine_obstacle() is not None:
            return None
        finite = all(np.all(np.isfinite(source.template()[0])) for _, _, source in self._anchor_items())
        return 'mixture' if finite else 'exact'

    def _template_engine_obstacle(self):
        """None if every row of the anchor grid is a stock histogram template on shared bin edges, else why not."""
        if len(self.shape_parameters) and self.source_wise_interpolation:
            return "source-wise interpolation is not supported"
        if len(self.base_model.config['analysis_space']) > _cabi.MAX_SPACE_DIMS:
            return "too many analysis dimensions"
        key = None
        for _, _, source in self._anchor_items():
            if not (isinstance(source, HistogramPdfSource) and type(source).pdf is HistogramPdfSource.pdf):
                return "source %r is not a stock HistogramPdfSource" % (source.name,)
            _, edges, method = source.template()
            k = (tuple(np.asarray(e, dtype=np.float64).tobytes() for e in edges), method)
            if key is None:
                key = k
            elif k != key:
                return "sources differ in bin edges or lookup method"
        if self._grid.n_corners * len(self.source_name_list) > _cabi.TS_MAX_TERMS:
            return "more than %d contraction terms" % _cabi.TS_MAX_TERMS
        return None

    def _build_template_engine(self, mode='exact'):
        """TemplateUnbinnedEngine over the templates of all anchor models (HBM / L2 resident)."""
        items = self._anchor_items()
        _, edges, method = items[0][2].template()
        templates = np.stack([source.template()[0] for _, _, source in items])
        return TemplateUnbinnedEngine(self._grid, self._mus_anchor.reshape(self._grid.n_anchors, -1), templates,
                                      edges, method, outlier_likelihood=self.config.get('outlier_likelihood', 1e-12),
                                      allow_negative=self.source_allowed_negative, mode=mode)

    # ------------------------------------------------------------------------------------------
    # many datasets, one parameter point each (toy Monte Carlos; not in the reference API)
    # ------------------------------------------------------------------------------------------
    @_needs_preparation
    def set_toy_data(self, datasets, offsets=None):
        """Load T datasets for batch_toys.

        datasets: a sequence of T datasets (each anything set_data accepts), ONE dataset holding the events of
        all toys back to back together with offsets [T + 1] (first event of every toy), or a device-resident
        blueice_b200.toys.ToyData (Model.simulate_toys).  The toys live on the
        device next to the anchor templates; nothing per event and anchor is stored (template-space engine), so
        1e6 toys x 1e3 events need 36 bytes per event."""
        reason = self._template_engine_obstacle()
        if reason:
            raise NotImplementedError("set_toy_data: " + reason)
        from .toys import ToyData
        if isinstance(datasets, ToyData):
            dims = [dim[0] for dim in self.base_model.config['analysis_space']]
            if datasets.dims != dims:
                raise ValueError("toy data has dimensions %s, the model %s" % (datasets.dims, dims))
            coords, offsets = datasets.coords, datasets.offsets
        elif offsets is None:
            sizes = [len(d) for d in datasets]
            offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
            per_toy = [self.base_model.to_analysis_dimensions(d) for d in datasets]
            n_space = len(self.base_model.config['analysis_space'])
            coords = np.empty((n_space, int(offsets[-1])), dtype=np.float64)
            for t, dims in enumerate(per_toy):
                for k in range(n_space):
                    coords[k, offsets[t]:offsets[t + 1]] = dims[k]
        else:
            dims = self.base_model.to_analysis_dimensions(datasets)
            coords = np.asarray([np.asarray(c, dtype=np.float64) for c in dims])
        engine = self._build_template_engine()
        engine.set_datasets(coords, offsets)
        self._toy_engine = engine
        return self

    def batch_toys(self, params, names=None, livetime_days=None, toy_index=None):
        """Log likelihood of toy t at parameter point params[t] for all T toys of set_toy_data, one device pass.

        Element t equals `self.set_data(toy_t); self(**dict(zip(names, params[t])))` bit for bit.
        With toy_index [Q], row q of params is evaluated on toy toy_index[q] instead (any number of points per
        toy: the finite-difference batches and line searches of many toy fits in lock step)."""
        engine = getattr(self, '_toy_engine', None)
        if engine is None:
            raise NotPreparedException("set_toy_data must be called before batch_toys")
        zs, mult = self._rows_from_params(params, names)
        view = _ToyView(engine, None if toy_index is None else np.asarray(toy_index, dtype=np.int64))
        return self._evaluate_rows(view, zs, mult, livetime_days, scalar=False)

    @property
    def n_toys(self):
        engine = getattr(self, '_toy_engine', None)
        return 0 if engine is None else engine.n_datasets

    def _fill_anchor_rows(self, engine, items, d):
        """Build the per-event pdf rows in HBM (likelihood.py:557-562 -> model.py:97-99; source-wise :534-549).

        items: (anchor index, source index, Source) per row to fill (source-wise: anchor index = absolute row).
        Histogram-backed sources that share bin edges and lookup method are evaluated for ALL their rows
        with one gather kernel (K3); any other Source.pdf is called on the host and its row uploaded."""
        torch = engine.torch
        dims = self.base_model.to_analysis_dimensions(d)
        coords = [np.asarray(c, dtype=np.float64) for c in dims]
        n = len(coords[0]) if len(coords) else len(d)
        engine.allocate_ps_anchor(n)
        if n == 0:
            return
        groups = OrderedDict()          # (edges bytes, method) -> [(anchor, source, histogram)]
        for g, s, source in items:
            on_device = (isinstance(source, HistogramPdfSource)
                         and type(source).pdf is HistogramPdfSource.pdf
                         and len(dims) <= _cabi.MAX_SPACE_DIMS)
            if on_device:
                hist, edges, method = source.template()
                key = (tuple(np.asarray(e, dtype=np.float64).tobytes() for e in edges), method)
                groups.setdefault(key, (edges, method, []))[2].append((g, s, hist))
            else:
                engine.set_rows(g, s, source.pdf(*dims))
        if groups:
            host = np.ascontiguousarray(np.asarray(coords, dtype=np.float64))
            if np.isnan(host).any() and any(m == 'linear' for _, m, _ in groups.values()):
                raise ValueError("One of the requested xi is out of bounds in dimension 0")
            coords_dev = torch.from_numpy(host).to(engine.device)
            for edges, method, items in groups.values():
                rows = [(g, s) for g, s, _ in items]
                templates = np.stack([h for _, _, h in items])
                engine.lookup_rows(rows, templates, edges,
                                   coords_dev, _cabi.LOOKUP_LINEAR if method == 'linear' else _cabi.LOOKUP_PIECEWISE)

    def _device_loglikelihood(self, engine, zs, mult, scale, eff):
        return engine.evaluate(zs, mult, scale, eff, return_status=True)

    @_needs_data
    def batch_parts(self, params, names=None, livetime_days=None):
        """(sum_i log f_i [P], sum_s mu_s [P], status [P], prior sum [P]) of this likelihood's events:
        the terms combined across ranks when events are sharded over GPUs (blueice_b200.distributed)."""
        zs, mult = self._rows_from_params(params, names)
        scale, _ = self._livetime_scale(livetime_days)
        scale_arr = None if scale is None else np.full(len(mult), scale, dtype=np.float64)
        logsum, musum, status = self._engine.evaluate(zs, mult, scale_arr, self._efficiencies(zs),
                                                      return_parts=True)
        return logsum, musum, status, self._prior_sum(zs, mult)

    def _full_output(self, engine, z_row, mult_row, scale, eff_row, result):
        mus, ps = engine.ps(z_row, mult_row, scale, eff_row)
        return result, mus, ps

    @inherit_docstring_from(LogLikelihoodBase)
    def _compute_single_pdf(self, **kwargs):
        model = self._compute_single_model(**kwargs)
        return model.expected_events(), model.score_events(self._data), None

    def _call_computed_pdf(self, multipliers, settings, livetime_days, full_output, kwargs):
        mus, ps, _ = self._compute_single_pdf(**kwargs)
        outlier = self.config.get('outlier_likelihood', 1e-12)
        engine = UnbinnedEngine(MorphGrid([]), np.asarray(mus, dtype=np.float64)[np.newaxis, :],
                                outlier_likelihood=outlier, allow_negative=self.source_allowed_negative)
        engine.set_ps_anchor(np.asarray(ps, dtype=np.float64)[np.newaxis])
        return self._evaluate_fixed_model(engine, multipliers, settings, livetime_days, full_output)

    def _compute_likelihood(self, mus, pdf_values_at_events):
        """Extended unbinned log likelihood of explicit (mus, ps) arrays on the device (likelihood.py:571-573)."""
        return extended_loglikelihood(mus, pdf_values_at_events,
                                      outlier_likelihood=self.config.get('outlier_likelihood', 1e-12))


class BinnedLogLikelihood(LogLikelihoodBase):

    def __init__(self, pdf_base_config, likelihood_config=None, **kwargs):
        LogLikelihoodBase.__init__(self, pdf_base_config, likelihood_config, **kwargs)
        pdf_base_config['pdf_interpolation_method'] = 'piecewise'     # (sic) reference likelihood.py:580
        self.model_statistical_uncertainty_handling = self.config.get('model_statistical_uncertainty_handling')

    @inherit_docstring_from(LogLikelihoodBase)
    def prepare(self, *args):
        LogLikelihoodBase.prepare(self, *args)
        self.ps, self.n_model_events = self.base_model.pmf_grids()
        n_sources = len(self.source_name_list)
        use_bb = self.model_statistical_uncertainty_handling is not None
        if len(self.shape_parameters):
            if self.source_wise_interpolation:
                raise NotImplementedError("Source-wise interpolation not implemented for binned likelihoods")
            shape = list(self.ps.shape)
            pmf_anchor = self.morpher.anchor_tensor(lambda m: m.pmf_grids()[0], shape, self.anchor_models)
            self.ps_interpolator = _LazyInterpolator(self.morpher, pmf_anchor, shape)
            nm_anchor = None
            if use_bb:
                nm_anchor = self.morpher.anchor_tensor(lambda m: m.pmf_grids()[1], shape, self.anchor_models)
                self.n_model_events_interpolator = _LazyInterpolator(self.morpher, nm_anchor, shape)
        else:
            pmf_anchor = self.ps[np.newaxis]
            nm_anchor = self.n_model_events[np.newaxis] if use_bb else None
        self._pmf_anchor, self._nm_anchor = pmf_anchor, nm_anchor
        self._engine = None

    def _bb_source_index(self):
        if self.model_statistical_uncertainty_handling != 'bb_single':
            return None
        source_i = self.config.get('bb_single_source')
        if source_i is None:
            raise ValueError("You need to specify bb_single_source to use bb_single_source expectation adjustment")
        return self.base_model.get_source_i(source_i)

    def _build_engine(self):
        bb = self._bb_source_index()
        G = self._grid.n_anchors
        S = len(self.source_name_list)
        pmf = np.asarray(self._pmf_anchor, dtype=np.float64).reshape((G, S) + tuple(self.ps.shape[1:]))
        nm = None
        if bb is not None:
            nm = np.asarray(self._nm_anchor, dtype=np.float64).reshape(pmf.shape)
        return BinnedEngine(self._grid, self._mus_anchor.reshape(G, -1), pmf, nm, bb, bin_shape=self.ps.shape[1:])

    @inherit_docstring_from(LogLikelihoodBase)
    def set_data(self, d):
        LogLikelihoodBase.set_data(self, d)
        from . import device_ops
        dimnames, bins = zip(*self.base_model.config['analysis_space'])
        coords = self.base_model.to_analysis_dimensions(d)
        counts = device_ops.histogramdd(bins, coords) if len(d) else np.zeros([len(b) - 1 for b in bins])
        self.data_events_per_bin = Histdd.from_histogram(counts, bins, axis_names=dimnames)
        self._observed_dirty = True

    def _ensure_engine(self):
        if self._engine is None:
            self._engine = self._build_engine()
            self._observed_dirty = True
        if self._observed_dirty:
            self._engine.set_observed(self.data_events_per_bin.histogram)
            self._observed_dirty = False
        return self._engine

    @_needs_data
    def __call__(self, *args, **kwargs):
        self._ensure_engine()
        return LogLikelihoodBase.__call__(self, *args, **kwargs)

    @_needs_data
    def batch(self, *args, **kwargs):
        self._ensure_engine()
        return LogLikelihoodBase.batch(self, *args, **kwargs)

    # -- many datasets, one parameter point each (binned toys; not in the reference API) -------------------
    @_needs_preparation
    def set_toy_data(self, datasets, offsets=None):
        """Load T datasets for batch_toys: every toy is binned like set_data bins one dataset (likelihood.py:604-609).

        datasets: a device-resident blueice_b200.toys.ToyData (Model.simulate_toys), a sequence of T datasets, or ONE
        dataset holding all toys back to back together with offsets [T + 1]."""
        from .toys import ToyData
        dimnames, bins = zip(*self.base_model.config['analysis_space'])
        if isinstance(datasets, ToyData):
            coords, offsets = datasets.coords, datasets.offsets
        elif offsets is None:
            sizes = [len(d) for d in datasets]
            offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
            coords = np.empty((len(bins), int(offsets[-1])), dtype=np.float64)
            for t, d in enumerate(datasets):
                for k, c in enumerate(self.base_model.to_analysis_dimensions(d)):
                    coords[k, offsets[t]:offsets[t + 1]] = c
        else:
            coords = np.asarray([np.asarray(c, dtype=np.float64) for c in self.base_model.to_analysis_dimensions(datasets)])
        if self._engine is None:
            self._engine = self._build_engine()
            self._observed_dirty = True
        self._engine.set_observed_toys(bins, coords, offsets)
        return self

    @property
    def n_toys(self):
        return 0 if self._engine is None else getattr(self._engine, 'n_toys', 0)

    def batch_toys(self, params, names=None, livetime_days=None):
        """Log likelihood of toy t at params[t] for all T toys of set_toy_data, one device pass; element t equals
        `self.set_data(toy_t); self(**dict(zip(names, params[t])))` bit for bit."""
        if self._engine is None or not getattr(self._engine, 'n_toys', 0):
            raise NotPreparedException("set_toy_data must be called before batch_toys")
        zs, mult = self._rows_from_params(params, names)
        return self._evaluate_rows(_BinnedToyView(self, self._engine), zs, mult, livetime_days, scalar=False)

    @staticmethod
    def _raise_bb_flags(flags):
        if np.any(flags & _cabi.BB_ROOT1_POSITIVE):
            raise AssertionError("Beeston-Barlow: first root is not <= 0 in every bin "
                                 "(e.g. a bin without calibration events)")
        if np.any(flags & _cabi.BB_NEGATIVE_A):
            raise AssertionError("Beeston-Barlow: adjusted expectation is not >= 0 in every bin")

    def _device_loglikelihood(self, engine, zs, mult, scale, eff):
        ll, status, flags = engine.evaluate(zs, mult, scale, eff, return_status=True)
        self._raise_bb_flags(flags[status == 0])
        return ll, status

    def _full_output(self, engine, z_row, mult_row, scale, eff_row, result):
        _, mus, pmfs, flags = engine.pmfs(z_row, mult_row, scale, eff_row)
        return result, mus, pmfs

    @inherit_docstring_from(LogLikelihoodBase)
    def _compute_single_pdf(self, **kwargs):
        model = self._compute_single_model(**kwargs)
        ps, n_model_events = model.pmf_grids()
        return model.expected_events(), ps, n_model_events

    def _call_computed_pdf(self, multipliers, settings, livetime_days, full_output, kwargs):
        mus, ps, n_model_events = self._compute_single_pdf(**kwargs)
        bb = self._bb_source_index()
        engine = BinnedEngine(MorphGrid([]), np.asarray(mus, dtype=np.float64)[np.newaxis, :], ps,
                              None if bb is None else n_model_events, bb)
        engine.set_observed(self.data_events_per_bin.histogram)
        return self._evaluate_fixed_model(engine, multipliers, settings, livetime_days, full_output)

    def _fixed_engine(self, mus, pmfs, n_model_events):
        bb = self._bb_source_index()
        engine = BinnedEngine(MorphGrid([]), np.asarray(mus, dtype=np.float64)[np.newaxis, :],
                              np.asarray(pmfs, dtype=np.float64),
                              None if bb is None else np.asarray(n_model_events, dtype=np.float64), bb)
        engine.set_observed(self.data_events_per_bin.histogram)
        return engine

    @_needs_data
    def adjust_expectations(self, mus, pmfs, n_model_events):
        """Beeston-Barlow adjusted (mus, pmfs) of explicit arrays, computed on the device (likelihood.py:618-660)."""
        mus = np.array(mus, dtype=np.float64)
        pmfs = np.array(pmfs, dtype=np.float64)
        if self.model_statistical_uncertainty_handling != 'bb_single':
            return mus, pmfs
        engine = self._fixed_engine(mus, pmfs, n_model_events)
        _, mus_adj, pmfs_adj, flags = engine.pmfs(np.zeros(0), np.ones(len(mus)))
        self._raise_bb_flags(np.array([flags]))
        return mus_adj, pmfs_adj

    def _compute_likelihood(self, mus, pmfs):
        """Binned Poisson log likelihood of explicit (mus, pmfs) arrays on the device (likelihood.py:662-675)."""
        engine = BinnedEngine(MorphGrid([]), np.asarray(mus, dtype=np.float64)[np.newaxis, :],
                              np.asarray(pmfs, dtype=np.float64), None, None)
        engine.set_observed(self.data_events_per_bin.histogram)
        ll, status, _ = engine.evaluate(np.zeros((1, 0)), np.ones((1, len(mus))), return_status=True)
        return ll[0]


def extended_loglikelihood(mu, ps, outlier_likelihood=0.0):
    """Extended unbinned log likelihood -sum(mu) + sum_i log(sum_s mu_s ps[s, i]) on the device.

    Same semantics as blueice/likelihood.py:678-690 (NaN terms dropped, non-positive densities replaced
    by outlier_likelihood when it is non-zero); no unphysical-rate test is applied here."""
    mu = np.asarray(mu, dtype=np.float64)
    ps = np.asarray(ps, dtype=np.float64)
    engine = UnbinnedEngine(MorphGrid([]), np.ones((1, len(mu))), outlier_likelihood=outlier_likelihood,
                            allow_negative=[True] * len(mu))
    engine.set_ps_anchor(ps[np.newaxis])
    # rates enter as multipliers of unit anchors, so any finite mu is passed through unchanged
    ll, status = engine.evaluate(np.zeros((1, 0)), mu[np.newaxis, :], return_status=True)
    if status[0] != 0:
        # the reference has no rate check here; evaluate the definition for the (rare) flagged input
        return float('nan')
    return ll[0]


def _bb_root(a, p, U, d, sign):
    a, p, U, d = [np.asarray(x, dtype=np.float64) for x in (a, p, U, d)]
    with np.errstate(all='ignore'):
        disc = (U**2*p**2 + 2*U**2*p + U**2 + 2*U*a*p**2 + 2*U*a*p - 2*U*d*p**2 - 2*U*d*p
                + a**2*p**2 + 2*a*d*p**2 + d**2*p**2)
        return (-U*p - U + a*p + d*p + sign * np.sqrt(disc)) / (2*p*(p + 1))


def beeston_barlow_root1(a, p, U, d):
    """Closed-form root of the single-source Beeston-Barlow equations that the reference asserts to be
    non-positive (likelihood.py:693-700).  Host helper of the public API; the likelihood evaluates the
    same expression per bin inside the K4 kernel."""
    return _bb_root(a, p, U, d, -1.0)


def beeston_barlow_root2(a, p, U, d):
    """The physical root (likelihood.py:703-708)."""
    return _bb_root(a, p, U, d, +1.0)


def beeston_barlow_roots(a, p, U, d):
    return beeston_barlow_root1(a, p, U, d), beeston_barlow_root2(a, p, U, d)


class LogLikelihoodSum(object):
    """Weighted sum of several likelihoods that share parameters by name (likelihood.py:867-955).

    Pure keyword plumbing over ll(**kw) / ll.batch; adds a batch() that sums the members' batches."""

    def __init__(self, likelihood_list, likelihood_weights=None):
        self.likelihood_list = list(likelihood_list)
        self.likelihood_weights = [1 for _ in self.likelihood_list] if likelihood_weights is None \
            else likelihood_weights
        self.rate_parameters = dict()
        self.shape_parameters = dict()
        self.source_list = []
        self.pdf_base_config = {}
        self.likelihood_parameters = []
        for ll in self.likelihood_list:
            self.rate_parameters.update(ll.rate_parameters)
            self.shape_parameters.update(ll.shape_parameters)
            names = []
            for rate_name in ll.rate_parameters.keys():
                names.append(rate_name + _RATE_SUFFIX)
                self._remember_base(ll, rate_name)
            for shape_name in ll.shape_parameters.keys():
                names.append(shape_name)
                self._remember_base(ll, shape_name)
            self.likelihood_parameters.append(names)

    def _remember_base(self, ll, name):
        value = ll.pdf_base_config.get(name)
        if value is not None:
            self.pdf_base_config[name] = value

    def __call__(self, compute_pdf=False, livetime_days=None, **kwargs):
        total = 0.
        for i, (ll, names, weight) in enumerate(zip(self.likelihood_list, self.likelihood_parameters,
                                                    self.likelihood_weights)):
            mine = {k: v for k, v in kwargs.items() if k in names}
            livetime = livetime_days[i] if isinstance(livetime_days, list) else livetime_days
            total += weight * ll(compute_pdf=compute_pdf, livetime_days=livetime, **mine)
        return total

    def batch(self, params, names, livetime_days=None):
        params = np.asarray(params, dtype=np.float64)
        names = list(names)
        total = np.zeros(len(params))
        for i, (ll, own, weight) in enumerate(zip(self.likelihood_list, self.likelihood_parameters,
                                                  self.likelihood_weights)):
            cols = [j for j, n in enumerate(names) if n in own]
            livetime = livetime_days[i] if isinstance(livetime_days, list) else livetime_days
            total = total + weight * ll.batch(params[:, cols], [names[j] for j in cols], livetime_days=livetime)
        return total

    def split_results(self, result_dict):
        return [{k: v for k, v in result_dict.items() if k in names} for names in self.likelihood_parameters]

    def get_bounds(self, parameter_name=None):
        if parameter_name is None:
            return [self.get_bounds(p) for p in self.shape_parameters]
        if parameter_name in self.shape_parameters.keys():
            bounds = np.array([ll.get_bounds(parameter_name) for ll in self.likelihood_list
                               if parameter_name in ll.shape_parameters.keys()])
            lo, hi = np.max(bounds[:, 0]), np.min(bounds[:, 1])
            if hi <= lo:
                raise InvalidParameterSpecification("lower bound %s higher than upper bound!" % parameter_name)
            return lo, hi
        if parameter_name.endswith(_RATE_SUFFIX):
            return 0, float('inf')
        raise InvalidParameter("Non-existing parameter %s" % parameter_name)


class LogAncillaryLikelihood(object):
    """Analytic constraint term func({parameter: value}, **func_kwargs) (likelihood.py:958-1001)."""

    def __init__(self, func, parameter_list, config=None, func_kwargs=None):
        self.rate_parameters = dict()
        self.shape_parameters = OrderedDict((name, (None, None, None)) for name in parameter_list)
        self.source_list = []
        self.pdf_base_config = dict() if config is None else config
        self.func = func
        self.func_kwargs = dict() if func_kwargs is None else func_kwargs

    def get_bounds(self, parameter_name=None):
        if parameter_name is None:
            return [self.get_bounds(p) for p in self.shape_parameters]
        if parameter_name in self.shape_parameters.keys():
            return -np.inf, np.inf
        raise InvalidParameter("Non-existing parameter %s" % parameter_name)

    def __call__(self, **kwargs):
        values = OrderedDict((name, self.pdf_base_config[name]) for name in self.shape_parameters)
        values.update(kwargs)
        return self.func(values, **self.func_kwargs)


class LogLikelihoodReParam(object):
    """A likelihood seen through new parameters (likelihood.py:715-864).

    conv_config maps
        '<source>_rate_multiplier' -> dict(params=[new parameter names], func=callable)   the rate multiplier of the
                                      wrapped likelihood becomes func(*new) / func(*base values of new)
        '<new parameter>'          -> (anchor values, log prior, base value)             declared as a shape parameter
    New parameters must have a (truthy) base value in the model config.  Pure keyword plumbing over the wrapped
    likelihood; `batch` converts whole columns and makes ONE ll.batch call."""

    def __init__(self, likelihood, conv_config):
        self._wrapped = likelihood
        self.conv_config = conv_config
        self.check_conv_config()
        self.pdf_base_config = likelihood.pdf_base_config

    # -- bookkeeping ------------------------------------------------------------------------------
    def _conversions(self):
        """[(rate multiplier name, new parameter names, func)] in conv_config order."""
        return [(k, list(v["params"]), v["func"]) for k, v in self.conv_config.items() if k.endswith(_RATE_SUFFIX)]

    def check_conv_config(self):
        """The new parameters declared and the ones the conversions use must be the same set, and each needs a base
        value in the model config (likelihood.py:732-760)."""
        declared = [k for k in self.conv_config if not k.endswith(_RATE_SUFFIX)]
        used = []
        for v in self.conv_config.values():
            if isinstance(v, dict):
                used += [p for p in v["params"] if p not in used]
        assert set(declared) == set(used), "New parameters are not consistent, double check conv_config..."
        config = self._wrapped.base_model.config
        missing = ", ".join(p for p in declared if not config.get(p, False))
        assert missing == "", "%s are missing in the config" % missing

    @property
    def rate_parameters(self):
        """The wrapped rate parameters without the ones now computed from new parameters."""
        kept = deepcopy(self._wrapped.rate_parameters)
        for source_name in self._wrapped.rate_parameters:
            if source_name + _RATE_SUFFIX in self.conv_config:
                kept.pop(source_name)
        return kept

    @property
    def shape_parameters(self):
        """The wrapped shape parameters plus the new parameters of conv_config."""
        out = deepcopy(self._wrapped.shape_parameters)
        for name, spec in self.conv_config.items():
            if not name.endswith(_RATE_SUFFIX):
                out[name] = ({z: z for z in spec[0]}, spec[1], spec[2])
        return out

    @property
    def base_model(self):
        model = deepcopy(self._wrapped.base_model)
        model.simulate = self._simulate
        return model

    def set_data(self, d):
        self._wrapped.set_data(d)

    def get_bounds(self, parameter_name=None):
        if parameter_name is None:
            return [self.get_bounds(p) for p in self.shape_parameters.keys()]
        if parameter_name in list(self._wrapped.rate_parameters) + list(self._wrapped.shape_parameters):
            return self._wrapped.get_bounds(parameter_name)
        zs = list(self.shape_parameters[parameter_name][0].keys())
        return min(zs), max(zs)

    # -- conversion ---------------------------------------------------------------------------------
    def _parameter_converter(self, with_suffix=True, **kwargs):
        """New-parameter keywords -> keywords of the wrapped likelihood (likelihood.py:812-864).  with_suffix=False:
        rate parameters are named by their source (Model.simulate's rate_multipliers convention) on both sides."""
        rate_names = list(self._wrapped.rate_parameters.keys())
        if not with_suffix:
            kwargs = {(k + _RATE_SUFFIX if k in rate_names else k): v for k, v in kwargs.items()}
        converted = OrderedDict()
        consumed = set()
        for target, params, func in self._conversions():
            base = [self.pdf_base_config.get(p) for p in params]
            values = [kwargs.get(p, b) for p, b in zip(params, base)]
            converted[target] = func(*values) / func(*base)
            consumed.update(params)
        for k, v in kwargs.items():
            if k not in consumed:
                converted[k] = v
        if not with_suffix:
            converted = OrderedDict((k.split(_RATE_SUFFIX)[0], v) for k, v in converted.items())
        return deepcopy(converted)

    def _simulate(self, kwargs=None, livetime_days=None):
        """base_model.simulate in terms of the new parameters (likelihood.py:796-810)."""
        converted = self._parameter_converter(with_suffix=False, **(kwargs or {}))
        multipliers = {k: v for k, v in converted.items() if k in self._wrapped.rate_parameters}
        return self._wrapped.base_model.simulate(rate_multipliers=multipliers, livetime_days=livetime_days)

    # -- evaluation ---------------------------------------------------------------------------------
    def __call__(self, compute_pdf=False, livetime_days=None, **kwargs):
        return self._wrapped(compute_pdf=compute_pdf, livetime_days=livetime_days,
                             **self._parameter_converter(**kwargs))

    def parameter_names(self):
        return [s + _RATE_SUFFIX for s in self.rate_parameters.keys()] + list(self.shape_parameters.keys())

    def batch(self, params, names=None, livetime_days=None):
        """P parameter points in the NEW parameters with one device pass: row p equals
        self(**dict(zip(names, params[p]))).  The conversion functions are applied to whole columns when they accept
        arrays, row by row otherwise."""
        names = self.parameter_names() if names is None else list(names)
        params = np.asarray(params, dtype=np.float64).reshape(-1, max(len(names), 1))
        P = len(params)
        columns = OrderedDict()
        consumed = set()
        for target, new_names, func in self._conversions():
            base = [self.pdf_base_config.get(p) for p in new_names]
            cols = [params[:, names.index(p)] if p in names else np.full(P, b, dtype=np.float64)
                    for p, b in zip(new_names, base)]
            denominator = func(*base)
            values = None
            if P > 1:
                try:
                    values = np.asarray(func(*cols), dtype=np.float64) / denominator
                    if values.shape != (P,):
                        values = None
                except Exception:
                    values = None
            if values is None:
                values = np.array([func(*[c[i] for c in cols]) / denominator for i in range(P)], dtype=np.float64)
            columns[target] = values
            consumed.update(new_names)
        for j, name in enumerate(names):
            if name not in consumed:
                columns[name] = params[:, j]
        table = np.column_stack(list(columns.values())) if columns else np.zeros((P, 0))
        return self._wrapped.batch(table, list(columns.keys()), livetime_days=livetime_days)


# the inference helpers double as methods of every likelihood class (likelihood.py:1004-1007)
for _name in inference.__all__:
    for _cls in (LogLikelihoodBase, LogLikelihoodSum, LogAncillaryLikelihood, LogLikelihoodReParam):
        setattr(_cls, _name, getattr(inference, _name))
