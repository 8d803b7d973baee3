"""Device engine: owns HBM-resident anchor tensors and drives the C-ABI kernels.

PyTorch is used here only for device memory, pinned staging buffers and streams.  All numerics
run in blueice_b200/csrc (hand-written sm_100a CUDA) through blueice_b200._cabi; there is no CPU
fallback -- constructing an engine without a CUDA device raises.

Layout in HBM:
    mus_anchor  [G, S]            float64   expected events at every anchor, anchors in C order
    ps_anchor   [G, S, ld]        float64   per-event pdf values (unbinned) or pmf per bin (binned),
                                            ld = N rounded up to a multiple of 64, padding zeroed
Per batch of P points (workspace, reused between calls):
    zs [P, D], mult [P, S], scale [P], eff [P, S] -> cell [P, D], frac [P, D], corner [P, C],
    weight [P, C], mus [P, S], musum [P], status [P], partial [P, n_super], logl [P]
"""
import contextlib
import ctypes
import gc
import os
import sys
import time as _time

import numpy as np

from . import _cabi

_LD_ALIGN = 64


def _torch():
    import torch
    return torch


@contextlib.contextmanager
def capture_graph(torch, graph):
    """torch.cuda.graph(graph) with Python's cyclic collector switched off for the duration of the capture.

    A collection that happens to fire inside a capture destroys whatever garbage is around at that moment -- engines of
    earlier evaluations with their pinned buffers and CUDA graphs -- and those destructors call CUDA APIs a capturing
    thread must not call: the error is raised inside a destructor and the process aborts (seen once in eight full runs of
    the GPU tests, in the capture of BinnedEngine.evaluate's third call).  torch collects before the capture starts; this
    keeps it from collecting again until the capture has ended."""
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):     # other threads (NCCL watchdog) may call CUDA
            yield
    finally:
        if was_enabled:
            gc.enable()


def require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("blueice_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _cabi.load()
    return torch


def round_up(n, m):
    return ((int(n) + m - 1) // m) * m


class MorphGrid(object):
    """Host description of the regular anchor grid (pdf_morphers.py:45-50)."""

    def __init__(self, axes):
        self.axes = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)) for a in axes]
        self.n_dims = len(self.axes)
        if self.n_dims > _cabi.MAX_DIMS:
            raise ValueError("at most %d shape parameters per morph grid are supported" % _cabi.MAX_DIMS)
        self.shape = tuple(len(a) for a in self.axes)
        self.n_anchors = int(np.prod(self.shape)) if self.n_dims else 1
        self.n_corners = 1 << self.n_dims
        self.n_anchors_i32 = _cabi.as_i32(self.shape if self.n_dims else [0])
        self.axes_concat = _cabi.as_f64(np.concatenate(self.axes) if self.n_dims else [0.0])
        if self.axes_concat.size > _cabi.MAX_AXIS_POINTS:
            raise ValueError("too many anchor values in total (max %d)" % _cabi.MAX_AXIS_POINTS)
        # number of hypercube cells per dim (a one-point axis has a single degenerate cell)
        self.cells_per_dim = [max(n - 1, 1) for n in self.shape]

    def in_range(self, zs):
        """likelihood.py:345-347: min(anchor) <= z <= max(anchor) for every dim; NaN fails."""
        zs = np.asarray(zs, dtype=np.float64).reshape(-1, self.n_dims)
        ok = np.ones(len(zs), dtype=bool)
        with np.errstate(invalid='ignore'):
            for d, a in enumerate(self.axes):
                ok &= (a[0] <= zs[:, d]) & (zs[:, d] <= a[-1])
        return ok

    def cell_ids(self, zs):
        """Flat hypercube-cell id per point (host copy of the device rule, used only for bucketing)."""
        zs = np.asarray(zs, dtype=np.float64).reshape(-1, self.n_dims)
        flat = np.zeros(len(zs), dtype=np.int64)
        for d, a in enumerate(self.axes):
            n = len(a)
            if n == 1:
                c = np.zeros(len(zs), dtype=np.int64)
            else:
                c = np.clip(np.searchsorted(a, zs[:, d], side='right') - 1, 0, n - 2)
            flat = flat * self.cells_per_dim[d] + c
        return flat


class _Workspace(object):
    """Per-batch device buffers + one pinned staging buffer each way, grown on demand."""

    def __init__(self, torch, device):
        self.torch = torch
        self.device = device
        self.capacity = {}
        self.buf = {}
        self.version = 0              # bumped whenever a buffer is (re)allocated: captured CUDA graphs hold the old pointers

    def get(self, name, n, dtype, pinned=False):
        n = max(int(n), 1)
        key = (name, dtype, pinned)
        if self.capacity.get(key, 0) < n:
            cap = max(n, int(self.capacity.get(key, 0) * 1.5))
            self.version += 1
            if pinned:
                self.buf[key] = self.torch.empty(cap, dtype=dtype, pin_memory=True)
            else:
                self.buf[key] = self.torch.empty(cap, dtype=dtype, device=self.device)
            self.capacity[key] = cap
        return self.buf[key][:n]


class PointPlan(object):
    """Host-side schedule of a batch: which points go to which kernel (results do not depend on it)."""

    def __init__(self, n_points, in_range, stream_points, kernel='stream'):
        self.kernel = kernel                    # 'mma': work units of the DMMA kernel (scheduled on the device);
        self.n_points = n_points                # 'stream': the per-point streaming kernel (more contraction terms or
        self.in_range = in_range                # hypercube cells than the DMMA path takes; cross-check in the tests)
        self.stream_points = stream_points      # int32 [n_stream]


def plan_points(grid, zs):
    """Host-side schedule of the streaming path (pure function; results never depend on it): every in-range point is
    one streaming task list entry; out-of-range points get no work (their result is -inf)."""
    P = len(zs)
    in_range = grid.in_range(zs) if grid.n_dims else np.ones(P, dtype=bool)
    return PointPlan(P, in_range, np.nonzero(in_range)[0].astype(np.int32))


# up to ~32 (group, range) pairs per resident warp and full units + one remainder per cell (measured optimum);
# bit 30 selects the full-units split, 0 = sizes balanced inside a cell
_MMA_TARGET_UNITS = (int(os.environ.get('BI_MMA_TARGET_UNITS', 148 * 12 * 32))
                     | (int(os.environ.get('BI_MMA_FULL_UNITS', '1')) << 30))
_EMPTY_I32 = np.zeros(0, dtype=np.int32)
_X_SLOTS = int(os.environ.get('BI_X_SLOTS', '3'))                # rotating pinned landing buffers of a gathered result
_E2E_GRAPHS = os.environ.get('BI_E2E_GRAPHS', '1') != '0'     # replay the e2e sequence of a batch size as one CUDA graph
_DIRECT_IO = os.environ.get('BI_DIRECT_IO', '1') != '0'        # kernels read / write small batches in pinned host memory
_DIRECT_IO_MAX_BYTES = 1 << 20
_BM_MIN_EVENTS = 4000000                                      # toy sets from this size on are swept bin-major (K5c)


class _EngineBase(object):
    def __init__(self, grid, mus_anchor, allow_negative=None, device=None):
        torch = require_cuda()
        self.torch = torch
        self.lib = _cabi.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.grid = grid
        mus_anchor = np.ascontiguousarray(np.asarray(mus_anchor, dtype=np.float64)).reshape(grid.n_anchors, -1)
        self.n_sources = mus_anchor.shape[1]
        if self.n_sources > _cabi.MAX_SOURCES:
            raise ValueError("at most %d sources are supported" % _cabi.MAX_SOURCES)
        self.mus_anchor_host = mus_anchor
        self.mus_anchor = torch.from_numpy(mus_anchor).to(self.device)
        if allow_negative is not None and any(allow_negative):
            self.allow_negative = np.ascontiguousarray(np.asarray(allow_negative, dtype=np.uint8))
        else:
            self.allow_negative = None
        self.ws = _Workspace(torch, self.device)
        self.launches = 0          # number of kernel launches issued by this engine (for bench.py)

    # -- helpers --------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _upload_points(self, zs, mult, scale, eff):
        """One pinned staging buffer, one H2D copy.  Returns device views (zs, mult, scale, eff)."""
        torch = self.torch
        P = len(mult)
        D, S = self.grid.n_dims, self.n_sources
        n_scale = P if scale is not None else 0
        n_eff = P * S if eff is not None else 0
        total = P * D + P * S + n_scale + n_eff
        pin = self.ws.get("h2d", total, torch.float64, pinned=True)
        pin_np = pin.numpy()
        o = 0
        pin_np[o:o + P * D] = np.asarray(zs, dtype=np.float64).reshape(-1); o += P * D
        pin_np[o:o + P * S] = np.asarray(mult, dtype=np.float64).reshape(-1); o += P * S
        if n_scale:
            pin_np[o:o + P] = np.asarray(scale, dtype=np.float64).reshape(-1); o += P
        if n_eff:
            pin_np[o:o + P * S] = np.asarray(eff, dtype=np.float64).reshape(-1); o += P * S
        dev = self.ws.get("points_in", total, torch.float64)
        dev.copy_(pin, non_blocking=True)
        o = 0
        zs_d = dev[o:o + P * D]; o += P * D
        mult_d = dev[o:o + P * S]; o += P * S
        scale_d = dev[o:o + P] if n_scale else None; o += n_scale
        eff_d = dev[o:o + P * S] if n_eff else None
        return zs_d, mult_d, scale_d, eff_d, total * 8

    def _setup(self, P, zs_d, mult_d, scale_d, eff_d):
        """K1: bi_point_setup on device-resident inputs.  Returns the dict of device outputs."""
        torch = self.torch
        D, S, C = self.grid.n_dims, self.n_sources, self.grid.n_corners
        out = dict(
            cell=self.ws.get("cell", P * max(D, 1), torch.int32),
            frac=self.ws.get("frac", P * max(D, 1), torch.float64),
            corner=self.ws.get("corner", P * C, torch.int32),
            weight=self.ws.get("weight", P * C, torch.float64),
            mus=self.ws.get("mus", P * S, torch.float64),
            musum=self.ws.get("musum", P, torch.float64),
            status=self.ws.get("status", P, torch.int32),
        )
        rc = self.lib.bi_point_setup(
            D, _cabi.host_ptr(self.grid.n_anchors_i32), _cabi.host_ptr(self.grid.axes_concat), S, P,
            _cabi.dev_ptr(zs_d), _cabi.dev_ptr(mult_d), _cabi.dev_ptr(scale_d), _cabi.dev_ptr(eff_d),
            _cabi.dev_ptr(self.mus_anchor), _cabi.host_ptr(self.allow_negative),
            _cabi.dev_ptr(out["cell"]), _cabi.dev_ptr(out["frac"]), _cabi.dev_ptr(out["corner"]),
            _cabi.dev_ptr(out["weight"]), _cabi.dev_ptr(out["mus"]), _cabi.dev_ptr(out["musum"]),
            _cabi.dev_ptr(out["status"]), None, None, None, None, self._stream())
        _cabi.check(rc, "bi_point_setup")
        self.launches += 1
        return out

    def point_setup_host(self, zs, mult, scale=None, eff=None):
        """Run K1 and return its outputs as NumPy arrays (index-parity tests, 'error' mode messages)."""
        P = len(mult)
        D, S, C = self.grid.n_dims, self.n_sources, self.grid.n_corners
        zs_d, mult_d, scale_d, eff_d, _ = self._upload_points(zs, mult, scale, eff)
        o = self._setup(P, zs_d, mult_d, scale_d, eff_d)
        self.torch.cuda.current_stream(self.device).synchronize()
        return dict(cell=o["cell"].cpu().numpy().reshape(P, max(D, 1))[:, :D],
                    frac=o["frac"].cpu().numpy().reshape(P, max(D, 1))[:, :D],
                    corner=o["corner"].cpu().numpy().reshape(P, C),
                    weight=o["weight"].cpu().numpy().reshape(P, C),
                    mus=o["mus"].cpu().numpy().reshape(P, S),
                    musum=o["musum"].cpu().numpy(),
                    status=o["status"].cpu().numpy())


class UnbinnedEngine(_EngineBase):
    """Fused unbinned likelihood over a device-resident anchor tensor (K1 + K2 + finalize)."""

    def __init__(self, grid, mus_anchor, outlier_likelihood=1e-12, allow_negative=None, device=None):
        super().__init__(grid, mus_anchor, allow_negative, device)
        self.outlier_likelihood = float(outlier_likelihood)
        self.n_events = 0
        self.ld = 0
        self.ps_anchor = None
        self.force_kernel = None      # None (auto) | 'mma' | 'stream'  (tests / bench)
        self._fused_cache = {}        # batch size -> staging buffers, workspace, prebuilt C arguments
        self.peer_gather = None       # distributed.PeerGather: the exchange step of a sharded evaluation, issued on the
        self.peer_mode = 'gather'     # device right after finalize ('gather': logl rows of all ranks, point sharding;
        self.last_gathered = None     # 'sum': rank-ordered sum of the shards' log sums, event sharding)
        self.last_gathered_owned = True   # False: last_gathered is a view of a scratch buffer, copy before the next call
        self.last_total = None
        self.full_grid_layout = os.environ.get('BI_MMA_NO_TENSORMAP') is None   # rows = [G][S][ld] anchor tensor

    # -- set_data -------------------------------------------------------------------------------
    def allocate_ps_anchor(self, n_events):
        torch = self.torch
        self.n_events = int(n_events)
        self.ld = max(round_up(self.n_events, _LD_ALIGN), _LD_ALIGN)
        self.ps_anchor = torch.zeros((self.grid.n_anchors, self.n_sources, self.ld), dtype=torch.float64,
                                     device=self.device)
        self.n_super = int(self.lib.bi_num_superblocks(self.n_events))
        return self.ps_anchor

    def set_ps_anchor(self, ps_anchor_host):
        """ps_anchor_host: [G, S, N] (or [n1..nD, S, N]) float64 computed on the host by Source.pdf."""
        ps = np.asarray(ps_anchor_host, dtype=np.float64)
        n = ps.shape[-1]
        ps = ps.reshape(self.grid.n_anchors, self.n_sources, n)
        self.allocate_ps_anchor(n)
        if n:
            self.ps_anchor[:, :, :n].copy_(self.torch.from_numpy(np.ascontiguousarray(ps)))
        return self

    def set_rows(self, anchor_index, source_index, values_host):
        """Upload one [N] row (a Source.pdf result computed on the host)."""
        v = self.torch.from_numpy(np.ascontiguousarray(np.asarray(values_host, dtype=np.float64)))
        self.ps_anchor[anchor_index, source_index, :self.n_events].copy_(v)

    def _flat_row(self, anchor_index, source_index):
        return anchor_index * self.n_sources + source_index

    def lookup_rows(self, rows, templates_host, edges_list, coords_dev, method):
        """K3: fill ps_anchor rows [(anchor, source), ...] from histogram templates that share edges.

        templates_host [T, *bins]; coords_dev: torch [n_space, ld_coords] on device.
        Replaces the G*S calls of HistogramPdfSource.pdf in set_data (likelihood.py:557-560)."""
        torch = self.torch
        templates_host = np.ascontiguousarray(np.asarray(templates_host, dtype=np.float64))
        T = templates_host.shape[0]
        n_bins = _cabi.as_i32([len(e) - 1 for e in edges_list])
        edges = _cabi.as_f64(np.concatenate([np.asarray(e, dtype=np.float64) for e in edges_list]))
        tmpl = torch.from_numpy(templates_host.reshape(T, -1)).to(self.device)
        out = torch.empty((T, self.ld), dtype=torch.float64, device=self.device)
        rc = self.lib.bi_hist_lookup(_cabi.dev_ptr(tmpl), T, len(edges_list), _cabi.host_ptr(n_bins),
                                     _cabi.host_ptr(edges), _cabi.dev_ptr(coords_dev), coords_dev.shape[1],
                                     self.n_events, method, _cabi.dev_ptr(out), self.ld, None, self._stream())
        _cabi.check(rc, "bi_hist_lookup")
        self.launches += 1
        idx = torch.as_tensor([self._flat_row(a, s) for a, s in rows], device=self.device, dtype=torch.int64)
        flat = self.ps_anchor.view(-1, self.ld)
        if self.ld > self.n_events:
            out[:, self.n_events:] = 0.0
        flat.index_copy_(0, idx, out)

    # -- planning (host) ------------------------------------------------------------------------
    def plan(self, zs):
        """The batch's schedule: the fused DMMA path plans on the device; the streaming path lists its in-range points."""
        if self.uses_mma():
            return PointPlan(len(zs), None, _EMPTY_I32, kernel='mma')
        return plan_points(self.grid, zs)

    def uses_mma(self):
        """True when the batch runs through the fused device path (K1 -> device schedule -> DMMA K2 -> finalize)."""
        n_cells = int(np.prod([max(n - 1, 1) for n in self.grid.n_anchors_i32])) if self.grid.n_dims else 1
        return (self.force_kernel in (None, 'mma') and self.n_terms <= _cabi.MMA_MAX_TERMS
                and n_cells <= _cabi.PLAN_MAX_CELLS)

    @property
    def n_terms(self):
        """Contraction terms per point-event of K2: corners x sources of the full anchor grid."""
        return self.n_sources * self.grid.n_corners

    def mma_workspace(self, P):
        """(workspace tensor, dict of typed device views) for a P-point batch of the fused path."""
        torch = self.torch
        D, S = self.grid.n_dims, self.n_sources
        off = np.zeros(15, dtype=np.int64)
        _cabi.check(self.lib.bi_unbinned_workspace_layout(D, S, self.n_terms, P, self.n_events, _cabi.host_ptr(off)),
                    "bi_unbinned_workspace_layout")
        ws = self.ws.get("mma_ws", int(off[14]), torch.uint8)
        names = ["cell", "frac", "corner", "weight", "mus", "partial", "group_points", "groups", "header",
                 "row", "coef", "wterm", "term_source", "coef_chunks"]
        dtypes = [torch.int32, torch.float64, torch.int32, torch.float64, torch.float64, torch.float64,
                  torch.int32, torch.int32, torch.int32, torch.int32, torch.float64, torch.float64, torch.int32,
                  torch.float64]
        views = {"n_points": P}
        for i, (name, dt) in enumerate(zip(names, dtypes)):
            views[name] = ws[int(off[i]):int(off[i + 1])].view(dt)
        return ws, views

    def mma_plan(self, P, views, status_d):
        """Stage 2 of the fused path alone (bench / profiling): device-side schedule into the workspace views."""
        _cabi.check(self.lib.bi_unbinned_plan(
            self.grid.n_dims, _cabi.host_ptr(self.grid.n_anchors_i32), P, _cabi.dev_ptr(views["cell"]),
            _cabi.dev_ptr(status_d), int(self.lib.bi_mma_unit_points(self.n_terms, P)), self.n_events,
            _MMA_TARGET_UNITS & 0x3fffffff, _MMA_TARGET_UNITS >> 30,
            _cabi.dev_ptr(views["group_points"]), _cabi.dev_ptr(views["groups"]), _cabi.dev_ptr(views["header"]),
            self._stream()), "bi_unbinned_plan")

    def mma_k2(self, views):
        """Stage 3 of the fused path alone (bench / profiling): the persistent DMMA kernel on a planned workspace."""
        _cabi.check(self.lib.bi_unbinned_partials_mma(
            _cabi.dev_ptr(self.rows_tensor()), self.ld, self.n_events, self.n_terms, self.n_sources,
            _cabi.dev_ptr(views["group_points"]), _cabi.dev_ptr(views["groups"]), _cabi.dev_ptr(views["header"]),
            _cabi.dev_ptr(views["row"]), _cabi.dev_ptr(views["coef"]), _cabi.dev_ptr(views["wterm"]),
            _cabi.dev_ptr(views["term_source"]), _cabi.dev_ptr(views["mus"]), self.outlier_likelihood,
            _cabi.dev_ptr(views["partial"]), self.grid.n_dims if self.full_grid_layout else -1,
            _cabi.host_ptr(self.grid.n_anchors_i32), _cabi.dev_ptr(views["cell"]), views["n_points"],
            _cabi.dev_ptr(views["coef_chunks"]), self._stream()),
            "bi_unbinned_partials_mma")

    def _k2_sequence_launches(self, P):
        """Kernels of one fused evaluation: K1, schedule, K2, finalize -- plus the coefficient packing of the K-chunk kernel
        (contractions of more than 32 terms)."""
        return 5 if int(self.lib.bi_mma_coef_chunks_doubles(self.n_terms, P)) > 0 else 4

    def _small_ok(self, P):
        """True when bi_unbinned_ll_batch evaluates a P-point batch with its single fused launch (tiny batches)."""
        if os.environ.get('BI_SMALL') == '0' or self.n_events <= 0 or not self.uses_mma():
            return False
        return bool(self.lib.bi_unbinned_small_ok(self.grid.n_dims, self.n_sources, P, self.n_events))

    def rows_tensor(self):
        """The [n_rows, ld] per-event pdf matrix K2 contracts (the anchor tensor viewed as rows)."""
        return self.ps_anchor

    def _fused_args(self, P, zs_d, mult_d, scale_d, eff_d, ws, out):
        """(C function, argument tuple) of the fused call for these device buffers."""
        return self.lib.bi_unbinned_ll_batch, (
            self.grid.n_dims, _cabi.host_ptr(self.grid.n_anchors_i32), _cabi.host_ptr(self.grid.axes_concat),
            self.n_sources, P, _cabi.dev_ptr(zs_d), _cabi.dev_ptr(mult_d), _cabi.dev_ptr(scale_d), _cabi.dev_ptr(eff_d),
            _cabi.dev_ptr(self.mus_anchor), _cabi.host_ptr(self.allow_negative),
            _cabi.dev_ptr(self.ps_anchor), self.ld, self.n_events, self.outlier_likelihood, _MMA_TARGET_UNITS,
            _cabi.dev_ptr(ws), ws.numel(), _cabi.dev_ptr(out["logl"]), _cabi.dev_ptr(out["logsum"]),
            _cabi.dev_ptr(out["musum"]), _cabi.dev_ptr(out["status"]))

    def run_fused(self, P, zs_d, mult_d, scale_d, eff_d):
        """ONE C-ABI call: K1 -> device schedule -> K2 (DMMA) -> finalize.  Returns dict of device outputs."""
        torch = self.torch
        ws, _ = self.mma_workspace(P)
        out = dict(logl=self.ws.get("logl", P, torch.float64), logsum=self.ws.get("logsum", P, torch.float64),
                   musum=self.ws.get("musum", P, torch.float64), status=self.ws.get("status", P, torch.int32))
        fn, args = self._fused_args(P, zs_d, mult_d, scale_d, eff_d, ws, out)
        _cabi.check(fn(*args, self._stream()), fn.__name__)
        self.launches += 1 if self._small_ok(P) else (self._k2_sequence_launches(P) if self.n_super > 0 else 2)
        return out

    def scalar_runner(self, has_scale):
        """The leanest e2e path, for ONE point at a time (ll(**params) inside a minimiser or an interval search):
        returns (pinned input array [D + S (+ 1)], run) where run() evaluates the staged point and returns
        (logl, status).  With the single-launch kernel the call is one ctypes call + one stream synchronisation (inputs
        and results travel over PCIe inside the kernel); otherwise the cached sequence / CUDA graph of evaluate_fused."""
        st = self._fused_state(1, has_scale, False)
        stream = self.torch.cuda.current_stream(self.device)
        stream_ptr = ctypes.c_void_p(stream.cuda_stream)
        pin_f, pin_i = st["pin_f_np"], st["pin_i_np"]
        sync = stream.synchronize
        if st["zero_copy"]:
            fn, args = st["fn"], st["args"] + (stream_ptr,)

            def run():
                rc = fn(*args)
                if rc:
                    _cabi.check(rc, "bi_unbinned_ll_batch")
                sync()
                self.launches += 1
                return pin_f[0], pin_i[0]
        else:
            n_launch = 4 if self.n_super > 0 else 2

            def run():
                graph = self._fused_graph(st, 1)
                if graph is not None:
                    graph.replay()
                else:
                    self._fused_sequence(st, 1, stream)
                sync()
                self.launches += n_launch
                return pin_f[0], pin_i[0]
        return st["pin_in_np"], run

    def batch_runner(self, P, has_scale):
        """The lean e2e path for a P-point batch (ll.batch inside a minimiser or scan driver): returns
        (zs view [P, D], mult view [P, S], scale view [P] or None -- all in the pinned staging buffer -- and run), where
        run() evaluates the staged points and returns (logl [P], status [P]) as views of the pinned result buffers
        (valid until the next evaluation).  Same device sequence as evaluate_fused."""
        st = self._fused_state(P, has_scale, False)
        D, S = self.grid.n_dims, self.n_sources
        pin = st["pin_in_np"]
        zs_v = pin[:P * D].reshape(P, D)
        mult_v = pin[P * D:P * D + P * S].reshape(P, S)
        scale_v = pin[P * D + P * S:P * D + P * S + P] if has_scale else None
        stream = self.torch.cuda.current_stream(self.device)
        sync = stream.synchronize
        logl_v, status_v = st["pin_f_np"][:P], st["pin_i_np"]
        pg, mode = self.peer_gather, self.peer_mode            # a sharded evaluation: the exchange launch rides along
        n_launch = 1 if st["zero_copy"] else (4 if self.n_super > 0 else 2)
        n_x = 0
        x_np = None
        local_in_block = False
        if pg is not None:
            n_launch += 0 if pg.fallback is not None else 1
            n_x = P if mode == 'sum' else pg.world * pg.n
            # landing buffers of the exchange result in pinned memory.  Slot 0 is scratch (its contents are copied out by
            # the caller); a gather of all ranks' rows (point / toy sharding) rotates over slots 1.._X_SLOTS and hands the
            # buffer itself to the caller -- a slot is reused only when no array or view of the previous result is alive
            # (CPython reference count of the slot's ndarray: every NumPy view chains back to it), so a gathered block of
            # world * P values costs no host copy.  Every slot has its own CUDA graph (the address is part of the launch).
            n_slots = 1 + (_X_SLOTS if mode != 'sum' and pg.fallback is None else 0)
            x_np = [self._pin_x(st, n_x, k).numpy() for k in range(n_slots)]
            local_in_block = mode != 'sum' and pg.n == P and pg.fallback is None
        h2d, d2h = st["n_in"] * 8, (P * 4 if local_in_block else P * 12) + n_x * 8
        getrefcount = sys.getrefcount
        last_slot = [0]

        def run():
            if self.peer_gather is not pg or self.peer_mode != mode:
                raise RuntimeError("batch_runner: the engine's exchange set-up changed since the runner was built")
            slot = 0
            if pg is not None:
                self.last_gathered = self.last_total = None      # (the views of the previous call are released)
                for k in range(1, len(x_np)):
                    cand = 1 + (last_slot[0] + k - 1) % (len(x_np) - 1)
                    if getrefcount(x_np[cand]) == 2:             # the list and this call: nobody holds the last result
                        slot = last_slot[0] = cand
                        break
            graph = self._fused_graph(st, P, slot)
            if graph is not None:
                graph.replay()
            else:
                self._fused_sequence(st, P, stream, slot)
            sync()
            self.launches += n_launch
            self.last_h2d_bytes, self.last_d2h_bytes = h2d, d2h
            if pg is None:
                return logl_v, status_v
            self.last_gathered_owned = slot != 0                 # True: the caller may keep the block without copying it
            if mode == 'sum':
                self.last_total = x_np[slot][:P]
                return logl_v, status_v
            block = x_np[slot].reshape(pg.world, -1)
            self.last_gathered = block
            return (block[pg.rank] if local_in_block else logl_v), status_v
        return zs_v, mult_v, scale_v, run

    def _fused_state(self, P, has_scale, has_eff):
        """Everything the e2e fast path needs for a P-point batch, built once and reused while the dataset stays:
        pinned staging buffers both ways, device input / output buffers, the workspace and the prebuilt C arguments."""
        zero_copy = self.peer_gather is None and self._small_ok(P)
        D, S = self.grid.n_dims, self.n_sources
        n_in = P * D + P * S + (P if has_scale else 0) + (P * S if has_eff else 0)
        # mid-size batches: K1 reads the staged points from pinned host memory and the finalize kernel writes logL there
        # (both over PCIe, inside the kernels), which removes two of the three copy nodes around the four launches
        direct_in = not zero_copy and _DIRECT_IO and 0 < n_in * 8 <= _DIRECT_IO_MAX_BYTES
        direct_out = direct_in and self.peer_gather is None
        key = (P, has_scale, has_eff, self.n_events, self.ps_anchor.data_ptr(), zero_copy, direct_in, direct_out)
        st = self._fused_cache.get(key)
        if st is not None:
            return st
        torch = self.torch
        st = {"P": P, "zero_copy": zero_copy, "direct_in": direct_in, "direct_out": direct_out}
        st["pin_in"] = torch.empty(max(n_in, 1), dtype=torch.float64, pin_memory=True)
        st["pin_in_np"] = st["pin_in"].numpy()
        st["dev_in"] = torch.empty(max(n_in, 1), dtype=torch.float64, device=self.device)
        o = 0
        views = []
        for n in (P * D, P * S, P if has_scale else 0, P * S if has_eff else 0):
            views.append(st["dev_in"][o:o + n] if n else None)
            o += n
        st["n_in"] = n_in
        off = np.zeros(15, dtype=np.int64)
        _cabi.check(self.lib.bi_unbinned_workspace_layout(D, S, self.n_terms, P, self.n_events, _cabi.host_ptr(off)),
                    "bi_unbinned_workspace_layout")
        st["ws"] = torch.empty(int(off[14]), dtype=torch.uint8, device=self.device)
        st["out_f"] = torch.empty(3 * P, dtype=torch.float64, device=self.device)      # logl | logsum | musum
        st["out_i"] = torch.empty(P, dtype=torch.int32, device=self.device)
        st["pin_f"] = torch.empty(3 * P, dtype=torch.float64, pin_memory=True)
        st["pin_i"] = torch.empty(P, dtype=torch.int32, pin_memory=True)
        st["pin_f_np"], st["pin_i_np"] = st["pin_f"].numpy(), st["pin_i"].numpy()
        out = dict(logl=st["out_f"][:P], logsum=st["out_f"][P:2 * P], musum=st["out_f"][2 * P:], status=st["out_i"])
        if zero_copy or direct_in:
            # tiny batch: bi_unbinned_ll_batch runs ONE fused launch (bi_unbinned_ll_small) that reads the staged inputs
            # from pinned host memory and writes the results there -- no copy surrounds the launch
            o = 0
            views = []
            for n in (P * D, P * S, P if has_scale else 0, P * S if has_eff else 0):
                views.append(st["pin_in"][o:o + n] if n else None)
                o += n
        if zero_copy:
            out = dict(logl=st["pin_f"][:P], logsum=st["pin_f"][P:2 * P], musum=st["pin_f"][2 * P:], status=st["pin_i"])
        elif direct_out:
            out = dict(logl=st["pin_f"][:P], logsum=st["pin_f"][P:2 * P], musum=st["out_f"][2 * P:], status=st["out_i"])
        st["fn"], st["args"] = self._fused_args(P, views[0], views[1], views[2], views[3], st["ws"], out)
        if len(self._fused_cache) >= 8:
            self._fused_cache.clear()
        self._fused_cache[key] = st
        return st

    def _fused_sequence(self, st, n_f, stream, slot=0):
        """Everything the device does for one e2e evaluation of a cached state, issued on `stream`: H2D of the staged
        inputs, the fused call (four launches), the exchange step of a sharded evaluation (one launch), D2H."""
        P = st["P"]
        if st["zero_copy"]:
            _cabi.check(st["fn"](*st["args"], ctypes.c_void_p(stream.cuda_stream)), "bi_unbinned_ll_batch")
            return
        if st["n_in"] and not st["direct_in"]:
            st["dev_in"].copy_(st["pin_in"], non_blocking=True)
        _cabi.check(st["fn"](*st["args"], ctypes.c_void_p(stream.cuda_stream)), "bi_unbinned_ll_batch")
        if st["direct_out"]:
            if n_f > 2 * P:                                         # return_parts: the mu sums (K1 keeps them on the device)
                st["pin_f"][2 * P:3 * P].copy_(st["out_f"][2 * P:3 * P], non_blocking=True)
            st["pin_i"].copy_(st["out_i"], non_blocking=True)
            return
        pg = self.peer_gather
        if pg is not None:
            # the exchange kernel delivers its result to pinned host memory itself (no copy node)
            if self.peer_mode == 'sum':
                # event sharding: -musum + (rank-ordered sum of the shards' log sums), -inf where the point is unphysical
                pg.reduce(st["out_f"][P:2 * P], st["out_f"][2 * P:3 * P], st["out_i"], out=self._pin_x(st, P, slot))
            else:
                if n_f == P and pg.n == P and pg.fallback is None:
                    # point sharding: all ranks' logl rows; this rank's own rows are row `rank` of the gathered block and
                    # its status words travel in the same launch -- the sequence ends without a copy node
                    pg.gather(st["out_f"][:P], out=self._pin_x(st, pg.world * pg.n, slot), status=st["out_i"],
                              status_out=st["pin_i"])
                    return
                pg.gather(st["out_f"][:P], out=self._pin_x(st, pg.world * pg.n, slot))
        st["pin_f"][:n_f].copy_(st["out_f"][:n_f], non_blocking=True)
        st["pin_i"].copy_(st["out_i"], non_blocking=True)

    def _pin_x(self, st, n, slot=0):
        """Pinned landing buffer `slot` of the exchange results of a cached state (per size)."""
        pin_x = st.get(("pin_x", n, slot))
        if pin_x is None:
            pin_x = st[("pin_x", n, slot)] = self.torch.empty(n, dtype=self.torch.float64, pin_memory=True)
        return pin_x

    def _fused_graph(self, st, n_f, slot=0):
        """CUDA graph of the e2e sequence of this cached state (None: not built yet, disabled, or capture failed).
        Built on the SECOND call with a state, so that one-off evaluations and the lazily initialised kernel
        attributes of the first call stay outside the capture.  A sharded evaluation captures its exchange launch too
        (bi_peer_exchange keeps its epoch on the device); every rank issues one exchange per call either way."""
        if not _E2E_GRAPHS:
            return None
        pg = self.peer_gather
        if pg is not None and pg.fallback is not None:
            return None                                             # NCCL fallback: stay eager
        key = ("graph", n_f, None if pg is None else (id(pg), self.peer_mode), slot)
        if key in st:
            return st[key]
        ckey = ("calls",) + key[1:]
        st[ckey] = st.get(ckey, 0) + 1
        if st[ckey] < 2:
            return None
        torch = self.torch
        graph = None
        try:
            torch.cuda.current_stream(self.device).synchronize()
            g = torch.cuda.CUDAGraph()
            with capture_graph(torch, g):
                self._fused_sequence(st, n_f, torch.cuda.current_stream(self.device), slot)
            graph = g
        except Exception:                                           # capture not possible here: stay on the eager path
            graph = None
            try:
                torch.cuda.synchronize(self.device)
            except Exception:
                pass
        st[key] = graph
        return graph

    def evaluate_fused(self, zs, mult, scale, eff, return_status, return_parts):
        """The e2e path of the fused engine: stage -> H2D -> one C call (-> exchange) -> D2H -> sync."""
        P = len(mult)
        st = self._fused_state(P, scale is not None, eff is not None)
        pin = st["pin_in_np"]
        o = 0
        D, S = self.grid.n_dims, self.n_sources
        for arr, size in ((zs, P * D), (mult, P * S), (scale, P), (eff, P * S)):
            if arr is not None:
                a = np.asarray(arr, dtype=np.float64).reshape(-1)
                if a.size != size:
                    raise ValueError("batch input has %d values, expected %d" % (a.size, size))
                pin[o:o + size] = a
                o += size
        stream = self.torch.cuda.current_stream(self.device)
        n_f = 3 * P if return_parts else P
        pg = self.peer_gather
        graph = self._fused_graph(st, n_f)
        if graph is not None:
            # H2D, the four launches, the exchange and the D2H copies replayed as ONE CUDA graph (built on the second call
            # of a batch size): one launch instead of seven API calls on the host, tighter dependencies on the device
            graph.replay()
        else:
            self._fused_sequence(st, n_f, stream)
        self.launches += 1 if st["zero_copy"] else \
            (self._k2_sequence_launches(P) if self.n_super > 0 else 2) + (0 if pg is None or pg.fallback is not None else 1)
        stream.synchronize()
        self.last_gathered = self.last_total = None
        self.last_gathered_owned = True                 # (copies of the scratch slot)
        n_x = 0
        if pg is not None:
            n_x = P if self.peer_mode == 'sum' else pg.world * pg.n
            x = self._pin_x(st, n_x).numpy()
            if self.peer_mode == 'sum':
                self.last_total = x[:P].copy()
            else:
                self.last_gathered = x.reshape(pg.world, -1).copy()
                if n_f == P and pg.n == P and pg.fallback is None:
                    st["pin_f_np"][:P] = self.last_gathered[pg.rank]    # (the sequence skips the copy node of the local rows)
        self.last_h2d_bytes = st["n_in"] * 8            # zero-copy batches: read by the kernel over PCIe, same bytes
        self.last_d2h_bytes = (3 * P * 8 if st["zero_copy"] else n_f * 8) + P * 4 + n_x * 8
        if pg is not None and self.peer_mode != 'sum' and n_f == P and pg.n == P and pg.fallback is None:
            self.last_d2h_bytes -= P * 8
        res = st["pin_f_np"]
        if return_parts:
            return res[P:2 * P].copy(), res[2 * P:3 * P].copy(), st["pin_i_np"].copy()
        if return_status:
            return res[:P].copy(), st["pin_i_np"].copy()
        return res[:P].copy()

    def upload_plan(self, plan):
        """H2D of the schedule (one pinned buffer).  Returns device views + byte count."""
        torch = self.torch
        if plan.kernel == 'mma':
            return None, None, None, 0
        n_s = len(plan.stream_points)
        pin = self.ws.get("plan_h", n_s, torch.int32, pinned=True)
        pin.numpy()[:n_s] = plan.stream_points
        dev = self.ws.get("plan_d", n_s, torch.int32)
        dev.copy_(pin, non_blocking=True)
        return (dev[:n_s] if n_s else None), None, None, n_s * 4

    # -- evaluation -----------------------------------------------------------------------------
    def run_device(self, P, zs_d, mult_d, scale_d, eff_d, plan, plan_dev, want_setup=False):
        """Device-only part: K1 -> K2 (the fused DMMA call, or the streaming kernel) -> finalize.  Returns logl (device)."""
        torch = self.torch
        if self.ps_anchor is None:
            raise RuntimeError("set_ps_anchor / allocate_ps_anchor must be called first")
        if plan.kernel == 'mma':
            o = self.run_fused(P, zs_d, mult_d, scale_d, eff_d)
            return (o["logl"], o) if want_setup else o["logl"]
        S, C = self.n_sources, self.grid.n_corners
        o = self._setup(P, zs_d, mult_d, scale_d, eff_d)
        partial = self.ws.get("partial", P * max(self.n_super, 1), torch.float64)
        stream_d = plan_dev[0]
        st = self._stream()
        if self.n_super > 0 and len(plan.stream_points):
            rc = self.lib.bi_unbinned_partials_stream(
                _cabi.dev_ptr(self.ps_anchor), self.ld, self.n_events, S, C, _cabi.dev_ptr(stream_d),
                len(plan.stream_points), _cabi.dev_ptr(o["corner"]), _cabi.dev_ptr(o["weight"]),
                _cabi.dev_ptr(o["mus"]), _cabi.dev_ptr(o["status"]), self.outlier_likelihood,
                _cabi.dev_ptr(partial), st)
            _cabi.check(rc, "bi_unbinned_partials_stream")
            self.launches += 1
        logl = self.ws.get("logl", P, torch.float64)
        logsum = self.ws.get("logsum", P, torch.float64)
        o["logsum"] = logsum
        rc = self.lib.bi_unbinned_finalize(_cabi.dev_ptr(partial), self.n_super, _cabi.dev_ptr(o["musum"]),
                                           _cabi.dev_ptr(o["status"]), P, _cabi.dev_ptr(logl),
                                           _cabi.dev_ptr(logsum), st)
        _cabi.check(rc, "bi_unbinned_finalize")
        self.launches += 1
        if want_setup:
            return logl, o
        return logl

    def evaluate(self, zs, mult, scale=None, eff=None, return_status=False, return_parts=False):
        """Host buffers in, host buffers out (the e2e path): H2D, K1, K2, finalize, D2H.

        return_parts=True returns (sum_i log p_i [P], sum_s mu_s [P], status [P]) instead: the terms an
        event-sharded evaluation combines across ranks."""
        torch = self.torch
        P = len(mult)
        if P == 0:
            if return_parts:
                return np.zeros(0), np.zeros(0), np.zeros(0, dtype=np.int32)
            return (np.zeros(0), np.zeros(0, dtype=np.int32)) if return_status else np.zeros(0)
        if self.ps_anchor is None:
            raise RuntimeError("set_ps_anchor / allocate_ps_anchor must be called first")
        if self.uses_mma():
            return self.evaluate_fused(zs, mult, scale, eff, return_status, return_parts)
        zs = np.asarray(zs, dtype=np.float64).reshape(P, self.grid.n_dims)
        plan = self.plan(zs)
        zs_d, mult_d, scale_d, eff_d, nbytes = self._upload_points(zs, mult, scale, eff)
        plan_dev = self.upload_plan(plan)
        logl, o = self.run_device(P, zs_d, mult_d, scale_d, eff_d, plan, plan_dev[:3], want_setup=True)
        out_pin = self.ws.get("d2h", 3 * P, torch.float64, pinned=True)
        out_pin[:P].copy_(logl, non_blocking=True)
        if return_parts:
            out_pin[P:2 * P].copy_(o["logsum"], non_blocking=True)
            out_pin[2 * P:].copy_(o["musum"], non_blocking=True)
        st_pin = self.ws.get("d2h_status", P, torch.int32, pinned=True)
        st_pin.copy_(o["status"], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        self.last_h2d_bytes = nbytes + plan_dev[3]
        self.last_d2h_bytes = P * (28 if return_parts else 12)
        res = out_pin.numpy().copy()
        if return_parts:
            return res[P:2 * P], res[2 * P:], st_pin.numpy().copy()
        if return_status:
            return res[:P], st_pin.numpy().copy()
        return res[:P]

    def ps(self, z_row, mult_row, scale=None, eff=None):
        """(mus [S], ps [S, N]) for one point in the reference's operation order (full_output=True)."""
        torch = self.torch
        S, C = self.n_sources, self.grid.n_corners
        zs_d, mult_d, scale_d, eff_d, _ = self._upload_points(np.asarray(z_row, dtype=np.float64).reshape(1, -1),
                                                              np.asarray(mult_row, dtype=np.float64).reshape(1, -1),
                                                              None if scale is None else [scale],
                                                              None if eff is None else np.asarray(eff).reshape(1, -1))
        o = self._setup(1, zs_d, mult_d, scale_d, eff_d)
        out = torch.empty((S, max(self.n_events, 1)), dtype=torch.float64, device=self.device)
        rc = self.lib.bi_unbinned_ps(_cabi.dev_ptr(self.ps_anchor), self.ld, self.n_events, S, C,
                                     _cabi.dev_ptr(o["corner"]), _cabi.dev_ptr(o["weight"]), _cabi.dev_ptr(out),
                                     out.shape[1], self._stream())
        _cabi.check(rc, "bi_unbinned_ps")
        self.launches += 1
        return o["mus"].cpu().numpy().copy(), out[:, :self.n_events].cpu().numpy()


class SourcewiseUnbinnedEngine(UnbinnedEngine):
    """Unbinned likelihood with source-wise interpolation (likelihood.py:113-145,210-240,534-555): source s is
    morphed over its own sub-grid (the shape parameters in source_dims[s]), so the row matrix holds
    sum_s G_s rows instead of G * S and a point-event costs sum_s 2^D_s contraction terms instead of 2^D * S.

    grid: the FULL shape-parameter grid (bounds test, cell bucketing); source_dims[s]: sorted indices of the
    shape parameters source s depends on; mus_rows: expected events per (source, sub-anchor), sources
    concatenated, each source's sub-anchors in C order."""

    def __init__(self, grid, source_dims, mus_rows, outlier_likelihood=1e-12, allow_negative=None, device=None):
        self.source_dims = [tuple(int(d) for d in dims) for dims in source_dims]
        n_sources = len(self.source_dims)
        self.sub_shapes = [tuple(grid.shape[d] for d in dims) for dims in self.source_dims]
        n_rows_per_source = [int(np.prod(sh)) if len(sh) else 1 for sh in self.sub_shapes]
        self.row_base = _cabi.as_i32(np.concatenate([[0], np.cumsum(n_rows_per_source)[:-1]]))
        self.n_rows = int(np.sum(n_rows_per_source))
        self.dim_mask = np.ascontiguousarray(
            np.array([sum(1 << d for d in dims) for dims in self.source_dims], dtype=np.uint32))
        mus_rows = np.ascontiguousarray(np.asarray(mus_rows, dtype=np.float64).reshape(-1))
        if len(mus_rows) != self.n_rows:
            raise ValueError("mus_rows must hold %d values" % self.n_rows)
        # the base class wants a [G, S] table; it is not used by the source-wise K1
        super().__init__(grid, np.zeros((grid.n_anchors, n_sources)), outlier_likelihood, allow_negative, device)
        self.mus_rows_host = mus_rows
        self.mus_rows = self.torch.from_numpy(mus_rows).to(self.device)
        self.full_grid_layout = False
        self.copy_source = self.torch.from_numpy(
            np.array([len(d) == 0 for d in self.source_dims], dtype=np.uint8)).to(self.device)
        self._n_terms = int(self.lib.bi_sourcewise_terms(n_sources, _cabi.host_ptr(self.dim_mask)))

    @property
    def n_terms(self):
        return self._n_terms

    def allocate_ps_anchor(self, n_events):
        torch = self.torch
        self.n_events = int(n_events)
        self.ld = max(round_up(self.n_events, _LD_ALIGN), _LD_ALIGN)
        self.ps_anchor = torch.zeros((self.n_rows, self.ld), dtype=torch.float64, device=self.device)
        self.n_super = int(self.lib.bi_num_superblocks(self.n_events))
        return self.ps_anchor

    def set_ps_anchor(self, rows_host):
        """rows_host: [n_rows, N] per-event pdf values, sources concatenated (see class docstring)."""
        rows = np.ascontiguousarray(np.asarray(rows_host, dtype=np.float64)).reshape(self.n_rows, -1)
        self.allocate_ps_anchor(rows.shape[1])
        if rows.shape[1]:
            self.ps_anchor[:, :rows.shape[1]].copy_(self.torch.from_numpy(rows))
        return self

    def set_rows(self, row_index, source_index, values_host):
        """Upload one [N] row; row_index is the absolute row (row_base[source] + sub-anchor)."""
        v = self.torch.from_numpy(np.ascontiguousarray(np.asarray(values_host, dtype=np.float64)))
        self.ps_anchor[row_index, :self.n_events].copy_(v)

    def _flat_row(self, anchor_index, source_index):
        return anchor_index          # callers pass the absolute row (row_base[source] + sub-anchor) as "anchor"

    def _small_ok(self, P):
        return False                  # the single-launch path covers the full anchor grid only

    def uses_mma(self):
        if self.n_terms > _cabi.MMA_MAX_TERMS:
            raise NotImplementedError("source-wise interpolation supports at most %d contraction terms "
                                      "(sum over sources of 2^(shape parameters of the source)); got %d"
                                      % (_cabi.MMA_MAX_TERMS, self.n_terms))
        return True

    def _sw_setup(self, P, zs_d, mult_d, scale_d, eff_d, views, outs):
        rc = self.lib.bi_point_setup_sourcewise(
            self.grid.n_dims, _cabi.host_ptr(self.grid.n_anchors_i32), _cabi.host_ptr(self.grid.axes_concat),
            self.n_sources, _cabi.host_ptr(self.dim_mask), _cabi.host_ptr(self.row_base), P,
            _cabi.dev_ptr(zs_d), _cabi.dev_ptr(mult_d), _cabi.dev_ptr(scale_d), _cabi.dev_ptr(eff_d),
            _cabi.dev_ptr(self.mus_rows), _cabi.host_ptr(self.allow_negative),
            _cabi.dev_ptr(views["cell"]), _cabi.dev_ptr(views["frac"]), _cabi.dev_ptr(views["mus"]),
            _cabi.dev_ptr(outs["musum"]), _cabi.dev_ptr(outs["status"]), _cabi.dev_ptr(views["row"]),
            _cabi.dev_ptr(views["coef"]), _cabi.dev_ptr(views["wterm"]), _cabi.dev_ptr(views["term_source"]),
            self._stream())
        _cabi.check(rc, "bi_point_setup_sourcewise")
        self.launches += 1

    def _fused_args(self, P, zs_d, mult_d, scale_d, eff_d, ws, out):
        return self.lib.bi_unbinned_ll_batch_sourcewise, (
            self.grid.n_dims, _cabi.host_ptr(self.grid.n_anchors_i32), _cabi.host_ptr(self.grid.axes_concat),
            self.n_sources, _cabi.host_ptr(self.dim_mask), _cabi.host_ptr(self.row_base), P,
            _cabi.dev_ptr(zs_d), _cabi.dev_ptr(mult_d), _cabi.dev_ptr(scale_d), _cabi.dev_ptr(eff_d),
            _cabi.dev_ptr(self.mus_rows), _cabi.host_ptr(self.allow_negative),
            _cabi.dev_ptr(self.ps_anchor), self.ld, self.n_events, self.outlier_likelihood, _MMA_TARGET_UNITS,
            _cabi.dev_ptr(ws), ws.numel(), _cabi.dev_ptr(out["logl"]), _cabi.dev_ptr(out["logsum"]),
            _cabi.dev_ptr(out["musum"]), _cabi.dev_ptr(out["status"]))

    def _setup_views(self, zs, mult, scale, eff):
        torch = self.torch
        P = len(mult)
        zs_d, mult_d, scale_d, eff_d, _ = self._upload_points(zs, mult, scale, eff)
        _, views = self.mma_workspace(P)
        outs = dict(musum=self.ws.get("musum", P, torch.float64), status=self.ws.get("status", P, torch.int32))
        self._sw_setup(P, zs_d, mult_d, scale_d, eff_d, views, outs)
        return views, outs

    def point_setup_host(self, zs, mult, scale=None, eff=None):
        P = len(mult)
        D, S, K = self.grid.n_dims, self.n_sources, self.n_terms
        views, outs = self._setup_views(np.asarray(zs, dtype=np.float64).reshape(P, D), mult, scale, eff)
        self.torch.cuda.current_stream(self.device).synchronize()
        return dict(cell=views["cell"][:P * max(D, 1)].cpu().numpy().reshape(P, max(D, 1))[:, :D],
                    frac=views["frac"][:P * max(D, 1)].cpu().numpy().reshape(P, max(D, 1))[:, :D],
                    row=views["row"][:P * K].cpu().numpy().reshape(P, K),
                    wterm=views["wterm"][:P * K].cpu().numpy().reshape(P, K),
                    coef=views["coef"][:P * K].cpu().numpy().reshape(P, K),
                    term_source=views["term_source"][:K].cpu().numpy(),
                    mus=views["mus"][:P * S].cpu().numpy().reshape(P, S),
                    musum=outs["musum"].cpu().numpy(), status=outs["status"].cpu().numpy())

    def ps(self, z_row, mult_row, scale=None, eff=None):
        """(mus [S], ps [S, N]) for one point in the reference's operation order (full_output=True)."""
        torch = self.torch
        S = self.n_sources
        views, _ = self._setup_views(np.asarray(z_row, dtype=np.float64).reshape(1, -1),
                                     np.asarray(mult_row, dtype=np.float64).reshape(1, -1),
                                     None if scale is None else [scale],
                                     None if eff is None else np.asarray(eff).reshape(1, -1))
        out = torch.empty((S, max(self.n_events, 1)), dtype=torch.float64, device=self.device)
        rc = self.lib.bi_unbinned_ps_terms(_cabi.dev_ptr(self.ps_anchor), self.ld, self.n_events, self.n_terms, S,
                                           _cabi.dev_ptr(views["row"]), _cabi.dev_ptr(views["wterm"]),
                                           _cabi.dev_ptr(views["term_source"]), _cabi.dev_ptr(self.copy_source),
                                           _cabi.dev_ptr(out), out.shape[1], self._stream())
        _cabi.check(rc, "bi_unbinned_ps_terms")
        self.launches += 1
        return views["mus"][:S].cpu().numpy().copy(), out[:, :self.n_events].cpu().numpy()


class BinnedEngine(_EngineBase):
    """Binned Poisson likelihood with optional Beeston-Barlow over device-resident pmf tensors (K1 + K4)."""

    def __init__(self, grid, mus_anchor, pmf_anchor, n_model_anchor=None, bb_source=None, device=None,
                 bin_shape=None):
        """pmf_anchor / n_model_anchor: [n1..nD, S, *bins] (or [G, S, *bins] with bin_shape given)."""
        super().__init__(grid, mus_anchor, None, device)
        torch = self.torch
        pmf = np.asarray(pmf_anchor, dtype=np.float64)
        self.bin_shape = tuple(bin_shape) if bin_shape is not None else tuple(pmf.shape[grid.n_dims + 1:])
        self.n_bins = int(np.prod(self.bin_shape))
        pmf = pmf.reshape(grid.n_anchors, self.n_sources, self.n_bins)
        self.ld = round_up(self.n_bins, _LD_ALIGN)
        self.pmf_anchor = torch.zeros((grid.n_anchors, self.n_sources, self.ld), dtype=torch.float64, device=self.device)
        self.pmf_anchor[:, :, :self.n_bins].copy_(torch.from_numpy(np.ascontiguousarray(pmf)))
        self.bb_source = -1 if bb_source is None else int(bb_source)
        self.nm_anchor = None
        self.nm_sum_anchor = None
        if self.bb_source >= 0:
            nm = np.asarray(n_model_anchor, dtype=np.float64).reshape(grid.n_anchors, self.n_sources, self.n_bins)
            nm_i = np.ascontiguousarray(nm[:, self.bb_source, :])
            self.nm_anchor = torch.zeros((grid.n_anchors, self.ld), dtype=torch.float64, device=self.device)
            self.nm_anchor[:, :self.n_bins].copy_(torch.from_numpy(nm_i))
            # per-anchor sum over bins in NumPy's own (pairwise) order
            self.nm_sum_anchor = torch.from_numpy(np.ascontiguousarray(nm_i.sum(axis=1))).to(self.device)
        self.observed = None
        self.lgamma_obs = None
        self.n_chunks = int(self.lib.bi_num_superblocks(self.n_bins))
        self._graphs = {}             # CUDA graphs of repeated evaluations, per batch shape
        self._observed_version = 0    # bumped when the observed histograms are replaced (captured graphs hold their pointers)

    def set_observed(self, observed_host):
        from scipy.special import gammaln
        torch = self.torch
        obs = np.ascontiguousarray(np.asarray(observed_host, dtype=np.float64).reshape(-1))
        assert obs.size == self.n_bins
        self.observed = torch.from_numpy(obs).to(self.device)
        self.lgamma_obs = torch.from_numpy(np.ascontiguousarray(gammaln(obs + 1))).to(self.device)
        self._observed_version += 1
        self._graphs = {}
        return self

    def histogram_events(self, edges_list, coords_host):
        """np.histogramdd-compatible binning on device (likelihood.py:604-609); returns counts [*bins] on host."""
        torch = self.torch
        n_space = len(edges_list)
        n = len(coords_host[0]) if n_space else 0
        n_bins = _cabi.as_i32([len(e) - 1 for e in edges_list])
        edges = _cabi.as_f64(np.concatenate([np.asarray(e, dtype=np.float64) for e in edges_list]))
        counts = torch.zeros(int(np.prod(n_bins)), dtype=torch.int64, device=self.device)
        if n:
            coords = torch.from_numpy(np.ascontiguousarray(np.asarray(coords_host, dtype=np.float64))).to(self.device)
            rc = self.lib.bi_histogramdd(n_space, _cabi.host_ptr(n_bins), _cabi.host_ptr(edges), _cabi.dev_ptr(coords),
                                         coords.shape[1], n, _cabi.dev_ptr(counts), None, self._stream())
            _cabi.check(rc, "bi_histogramdd")
            self.launches += 1
        return counts.cpu().numpy().astype(np.float64).reshape([int(b) for b in n_bins])

    # -- many datasets, one parameter point each (binned toys) -------------------------------------------------
    def set_observed_rows(self, observed):
        """observed: [T, *bins] counts of T datasets (NumPy array or device tensor); point t of evaluate_toys is
        evaluated on row t.  gammaln(k + 1) comes from a host-computed table (scipy), as in set_observed, so a toy
        evaluates bit-identically to set_observed(row t) + evaluate."""
        from scipy.special import gammaln
        torch = self.torch
        if isinstance(observed, torch.Tensor):
            obs = observed.to(self.device, torch.float64).reshape(-1, self.n_bins)
        else:
            obs = torch.from_numpy(np.ascontiguousarray(np.asarray(observed, dtype=np.float64)).reshape(-1, self.n_bins)).to(self.device)
        T = obs.shape[0]
        k_max = int(obs.max().item()) if obs.numel() else 0
        integral = bool(torch.all((obs >= 0) & (obs == torch.floor(obs)))) if obs.numel() else True
        if not integral:
            raise ValueError("toy histograms must hold non-negative integer counts")
        table = torch.from_numpy(np.ascontiguousarray(gammaln(np.arange(k_max + 1, dtype=np.float64) + 1))).to(self.device)
        self.toy_observed = torch.zeros((T, self.ld), dtype=torch.float64, device=self.device)
        self.toy_observed[:, :self.n_bins] = obs
        self.toy_lgamma = torch.zeros((T, self.ld), dtype=torch.float64, device=self.device)
        self.toy_lgamma[:, :self.n_bins] = table[obs.to(torch.int64)]
        self.n_toys = T
        self._observed_version += 1
        self._graphs = {}
        return self

    def set_observed_toys(self, edges_list, coords, offsets):
        """Bin the events of T toys on the device (np.histogramdd semantics per toy: likelihood.py:604-609 applied to
        every toy) and keep the [T, bins] counts as the toys' observed histograms.  coords: [n_space, N] (device
        tensor or host array), offsets [T + 1]."""
        torch = self.torch
        n_space = len(edges_list)
        n_bins = _cabi.as_i32([len(e) - 1 for e in edges_list])
        if int(np.prod(n_bins)) != self.n_bins:
            raise ValueError("bin edges do not match the engine's %d bins" % self.n_bins)
        edges = _cabi.as_f64(np.concatenate([np.asarray(e, dtype=np.float64) for e in edges_list]))
        if not isinstance(coords, torch.Tensor):
            coords = torch.from_numpy(np.ascontiguousarray(np.asarray(coords, dtype=np.float64)))
        coords = coords.to(self.device, torch.float64).reshape(n_space, -1).contiguous()
        offsets = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
        T, n = len(offsets) - 1, int(coords.shape[1])
        if offsets[0] != 0 or offsets[-1] != n or np.any(np.diff(offsets) < 0):
            raise ValueError("dataset offsets must rise from 0 to the number of events")
        counts = torch.zeros((max(T, 1), self.n_bins), dtype=torch.int64, device=self.device)
        if n and T:
            off_d = torch.from_numpy(offsets).to(self.device)
            _cabi.check(self.lib.bi_histogramdd_toys(n_space, _cabi.host_ptr(n_bins), _cabi.host_ptr(edges),
                                                     _cabi.dev_ptr(coords), coords.shape[1], n, _cabi.dev_ptr(off_d), T,
                                                     _cabi.dev_ptr(counts), self.n_bins, self._stream()),
                        "bi_histogramdd_toys")
            self.launches += 1
        return self.set_observed_rows(counts[:T].to(torch.float64))

    def evaluate_toys(self, zs, mult, scale=None, eff=None, return_status=False):
        """Point t on toy t (set_observed_rows / set_observed_toys).  Returns logl [T] (+ status, bb flags)."""
        if getattr(self, 'toy_observed', None) is None:
            raise RuntimeError("set_observed_rows / set_observed_toys must be called first")
        if len(mult) != self.n_toys:
            raise ValueError("need one parameter point per toy: got %d points for %d toys" % (len(mult), self.n_toys))
        return self.evaluate(zs, mult, scale, eff, return_status, toys=True)

    def run_device(self, P, zs_d, mult_d, scale_d, eff_d, want_all=False, toys=False, need_mus_adj=False):
        torch = self.torch
        if toys:
            observed, lgamma_obs, stride = self.toy_observed, self.toy_lgamma, self.ld
        else:
            observed, lgamma_obs, stride = self.observed, self.lgamma_obs, 0
        if observed is None:
            raise RuntimeError("set_observed must be called first")
        S, C = self.n_sources, self.grid.n_corners
        o = self._setup(P, zs_d, mult_d, scale_d, eff_d)
        n_scratch = int(self.lib.bi_binned_scratch_doubles(P, self.n_bins))
        scratch = self.ws.get("scratch", n_scratch, torch.float64)
        logl = self.ws.get("logl", P, torch.float64)
        mus_adj = self.ws.get("mus_adj", P * S, torch.float64) if need_mus_adj else None    # full_output only
        flags = self.ws.get("flags", P, torch.int32)
        rc = self.lib.bi_binned_ll_batch_toys(
            _cabi.dev_ptr(self.pmf_anchor), _cabi.dev_ptr(self.nm_anchor), _cabi.dev_ptr(self.nm_sum_anchor),
            self.ld, self.n_bins, S, C, self.bb_source, _cabi.dev_ptr(observed), _cabi.dev_ptr(lgamma_obs), stride,
            _cabi.dev_ptr(o["corner"]), _cabi.dev_ptr(o["weight"]), _cabi.dev_ptr(o["mus"]), _cabi.dev_ptr(o["status"]),
            P, _cabi.dev_ptr(scratch), _cabi.dev_ptr(logl), _cabi.dev_ptr(mus_adj), _cabi.dev_ptr(flags), self._stream())
        _cabi.check(rc, "bi_binned_ll_batch_toys")
        self.launches += (5 if self.bb_source >= 0 else 3) + (1 if need_mus_adj else 0)
        if want_all:
            return logl, o, mus_adj, flags, scratch
        return logl

    def evaluate(self, zs, mult, scale=None, eff=None, return_status=False, toys=False):
        """Host in / host out.  Returns logl [P] (+ status [P], bb flags [P]).  Repeated evaluations of one batch shape
        (a minimiser's steps, a scan in chunks) replay the whole sequence -- H2D of the staged points, K1, the
        Beeston-Barlow passes and totals, D2H -- as ONE CUDA graph: at one point the launches take 0.13 ms on the device
        and the eager host path added 0.1 ms."""
        torch = self.torch
        P = len(mult)
        if P == 0:
            z = np.zeros(0)
            return (z, z.astype(np.int32), z.astype(np.int32)) if return_status else z
        zs = np.asarray(zs, dtype=np.float64).reshape(P, self.grid.n_dims)
        D, S = self.grid.n_dims, self.n_sources
        sizes = (P * D, P * S, P if scale is not None else 0, P * S if eff is not None else 0)
        total = sum(sizes)
        pin = self.ws.get("h2d", total, torch.float64, pinned=True)
        pin_np = pin.numpy()
        off = 0
        for arr, size in zip((zs, mult, scale, eff), sizes):
            if size:
                pin_np[off:off + size] = np.asarray(arr, dtype=np.float64).reshape(-1)
                off += size
        dev = self.ws.get("points_in", total, torch.float64)
        views, off = [], 0
        for size in sizes:
            views.append(dev[off:off + size] if size else None)
            off += size
        out_pin = self.ws.get("d2h", P, torch.float64, pinned=True)
        st_pin = self.ws.get("d2h_status", 2 * P, torch.int32, pinned=True)

        def device_sequence():
            dev.copy_(pin, non_blocking=True)
            logl, o, _, flags, _ = self.run_device(P, views[0], views[1], views[2], views[3], want_all=True, toys=toys)
            out_pin.copy_(logl, non_blocking=True)
            st_pin[:P].copy_(o["status"], non_blocking=True)
            st_pin[P:].copy_(flags, non_blocking=True)

        graph = None
        if _E2E_GRAPHS:
            gkey = (P, sizes, bool(toys), self._observed_version)
            entry = self._graphs.get(gkey)
            if entry is None:
                if len(self._graphs) >= 8:
                    self._graphs.clear()
                entry = self._graphs[gkey] = {"calls": 0, "graph": None, "version": None}
            entry["calls"] += 1
            if entry["graph"] is None and entry["calls"] >= 3 and entry.get("last_s", 1.0) < 5e-3:
                try:                                                # the first calls size the workspace buffers eagerly
                    torch.cuda.current_stream(self.device).synchronize()
                    launches = self.launches
                    g = torch.cuda.CUDAGraph()
                    with capture_graph(torch, g):
                        device_sequence()
                    entry["graph"], entry["version"], entry["n_launch"] = g, self.ws.version, self.launches - launches
                    self.launches = launches
                except Exception:
                    entry["graph"] = False
                    try:
                        torch.cuda.synchronize(self.device)
                    except Exception:
                        pass
            if entry["graph"] and entry["version"] == self.ws.version:
                graph = entry["graph"]
            elif entry["graph"]:
                entry["graph"], entry["calls"] = None, 1             # a workspace buffer moved: capture again later
        t_start = _time.perf_counter()
        if graph is not None:
            graph.replay()
            self.launches += entry["n_launch"]
        else:
            device_sequence()
        torch.cuda.current_stream(self.device).synchronize()
        if _E2E_GRAPHS:
            self._graphs[gkey]["last_s"] = _time.perf_counter() - t_start
        self.last_h2d_bytes = total * 8
        self.last_d2h_bytes = P * 16
        res = out_pin.numpy().copy()
        if return_status:
            st = st_pin.numpy().copy()
            return res, st[:P], st[P:]
        return res

    def pmfs(self, z_row, mult_row, scale=None, eff=None):
        """(logl, adjusted mus [S], adjusted pmfs [S, *bins], bb flags) for one point (full_output=True)."""
        torch = self.torch
        S, C = self.n_sources, self.grid.n_corners
        zs_d, mult_d, scale_d, eff_d, _ = self._upload_points(np.asarray(z_row, dtype=np.float64).reshape(1, -1),
                                                              np.asarray(mult_row, dtype=np.float64).reshape(1, -1),
                                                              None if scale is None else [scale],
                                                              None if eff is None else np.asarray(eff).reshape(1, -1))
        logl, o, mus_adj, flags, scratch = self.run_device(1, zs_d, mult_d, scale_d, eff_d, want_all=True, need_mus_adj=True)
        out = torch.empty((S, self.n_bins), dtype=torch.float64, device=self.device)
        o_t = int(self.lib.bi_binned_sum_t_offset(1, self.n_bins))
        sum_t = scratch[o_t:o_t + 1]
        rc = self.lib.bi_binned_pmfs(
            _cabi.dev_ptr(self.pmf_anchor), _cabi.dev_ptr(self.nm_anchor), _cabi.dev_ptr(self.nm_sum_anchor),
            self.ld, self.n_bins, S, C, self.bb_source, _cabi.dev_ptr(self.observed),
            _cabi.dev_ptr(o["corner"]), _cabi.dev_ptr(o["weight"]), _cabi.dev_ptr(o["mus"]),
            _cabi.dev_ptr(sum_t), _cabi.dev_ptr(out), self.n_bins, self._stream())
        _cabi.check(rc, "bi_binned_pmfs")
        self.launches += 1
        return (float(logl.cpu()[0]), mus_adj[:S].cpu().numpy().copy(),
                out.cpu().numpy().reshape((S,) + tuple(self.bin_shape)), int(flags.cpu()[0]))


_MIX_WIDE_MIN_SUPERBLOCKS = 16384       # 8.4e6 events (168 MB of prepared 2-D events): wide K5b groups from here on


def mixture_group_width(n_points, min_superblocks, wide_min_superblocks=None, tensor_pipe=True):
    """Points per group of a K5b (mixture-form) schedule, before the pairs are cut: 1 for a single point, 16 (two 8-point
    m-tiles per warp, one pass over the events per 16 points) when more than 8 points are evaluated on datasets of at
    least wide_min_superblocks superblocks -- large enough to outgrow the L2 and to fill the device with half as many
    units -- and the tensor-pipe kernel is in use (BI_MIX_MMA != '0'), else 8."""
    if wide_min_superblocks is None:
        wide_min_superblocks = _MIX_WIDE_MIN_SUPERBLOCKS
    if n_points <= 1:
        return 1
    if tensor_pipe and n_points > _cabi.MIX_GROUP_POINTS and min_superblocks >= wide_min_superblocks:
        return _cabi.MIX_GROUP_POINTS_WIDE
    return _cabi.MIX_GROUP_POINTS


def group_pairs(dataset_index, cells, n_cells, group_points):
    """Host-side grouping of (dataset, point) pairs for the template-space kernels (pure function).

    dataset_index [Q], cells [Q] (hypercube cell of every pair's point, -1 = out of range).  Pairs are sorted (stably) by
    (dataset, cell) and every run of equal keys is cut into groups of at most group_points pairs; out-of-range pairs are
    groups of their own.  Returns (order [Q], first [n_groups], count [n_groups]): group g holds the sorted positions
    first[g] .. first[g] + count[g] - 1, position j being pair order[j]."""
    dataset_index = np.asarray(dataset_index, dtype=np.int64)
    cells = np.asarray(cells, dtype=np.int64)
    Q = len(dataset_index)
    if Q == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z, z
    key = dataset_index * (int(n_cells) + 1) + (cells + 1)
    order = np.argsort(key, kind='stable')
    sk = key[order]
    starts = np.flatnonzero(np.r_[True, sk[1:] != sk[:-1]])
    ends = np.r_[starts[1:], Q]
    step = np.where(cells[order[starts]] < 0, 1, max(int(group_points), 1))
    n_chunks = -(-(ends - starts) // step)
    run = np.repeat(np.arange(len(starts)), n_chunks)
    idx_in_run = np.arange(int(n_chunks.sum())) - np.repeat(np.cumsum(n_chunks) - n_chunks, n_chunks)
    first = starts[run] + idx_in_run * step[run]
    count = np.minimum(step[run], ends[run] - first)
    return order, first.astype(np.int64), count.astype(np.int64)


class TemplateUnbinnedEngine(_EngineBase):
    """Template-space unbinned likelihood (K1 + K5 + ragged finalize): the per-event pdf values of the anchor
    tensor are looked up on the fly from HBM/L2-resident histogram templates instead of being stored.

    Serves (a) MANY datasets, each evaluated at its own parameter point(s) -- toy Monte Carlos (SURVEY.md section 8d
    config 4), and (b) one dataset whose dense anchor tensor [G, S, N] would not fit in HBM (config 5).  A
    (dataset, point) pair evaluates bit-identically to UnbinnedEngine on that dataset alone (same operations,
    same order: bi_template.cu).

    templates: [G * S, *bins] densities, row = anchor * S + source (anchors in C order); edges_list: bin edges
    per analysis dimension, shared by all templates; method: 'linear' | 'piecewise' (source.py:203,225-243)."""

    def __init__(self, grid, mus_anchor, templates, edges_list, method='linear', outlier_likelihood=1e-12,
                 allow_negative=None, device=None, mode='exact'):
        """mode 'exact': K5, per-event values formed like the reference (bit-identical to the anchor-tensor engine).
        mode 'mixture': K5b, the templates are morphed per point and the events looked up in the mixture template
        (one lookup per point-event instead of corners x sources; events are sorted by bin once per dataset so the
        kernel streams them at HBM speed; equal to 'exact' to ~1e-13 relative, not bit-identical; needs finite
        templates; one dataset at a time)."""
        super().__init__(grid, mus_anchor, allow_negative, device)
        torch = self.torch
        if mode not in ('exact', 'mixture'):
            raise ValueError("mode must be 'exact' or 'mixture'")
        self.mode = mode
        if mode == 'mixture' and not np.all(np.isfinite(np.asarray(templates, dtype=np.float64))):
            raise ValueError("mode='mixture' needs finite templates (the reference's per-source nansum cannot be "
                             "reproduced from a mixture template); use mode='exact'")
        self.outlier_likelihood = float(outlier_likelihood)
        self.edges_list = [np.ascontiguousarray(np.asarray(e, dtype=np.float64)) for e in edges_list]
        self.n_space = len(self.edges_list)
        if not 1 <= self.n_space <= _cabi.MAX_SPACE_DIMS:
            raise ValueError("templates must have 1..%d analysis dimensions" % _cabi.MAX_SPACE_DIMS)
        self.n_bins_i32 = _cabi.as_i32([len(e) - 1 for e in self.edges_list])
        self.edges_concat = _cabi.as_f64(np.concatenate(self.edges_list))
        self.method = {'linear': _cabi.LOOKUP_LINEAR, 'piecewise': _cabi.LOOKUP_PIECEWISE}[method]
        t = np.ascontiguousarray(np.asarray(templates, dtype=np.float64))
        self.n_rows = grid.n_anchors * self.n_sources
        self.n_template_bins = int(np.prod(self.n_bins_i32))
        t = t.reshape(self.n_rows, self.n_template_bins)
        self.n_terms = self.n_sources * grid.n_corners
        if self.n_terms > _cabi.TS_MAX_TERMS:
            raise NotImplementedError("the template-space kernel supports at most %d contraction terms "
                                      "(corners x sources); got %d" % (_cabi.TS_MAX_TERMS, self.n_terms))
        self.templates_rows = torch.from_numpy(t).to(self.device)             # [n_rows, B]: K3 / full_output / K5b
        linear = self.method == _cabi.LOOKUP_LINEAR
        if linear and mode == 'exact':
            # K5 reads a PACKED layout: every bin with its neighbours along the last (and second-last) dimension in 16 /
            # 32 aligned bytes, so the lookup corners come with one 128- / 256-bit gather; the neighbours of the last bin
            # of an axis are never used (cell <= n - 2)
            r = self.templates_rows
            s1 = 1 if self.n_bins_i32[-1] > 1 else 0
            parts = [r, torch.roll(r, -s1, dims=1)]
            if self.n_space >= 2:
                s2 = int(self.n_bins_i32[-1]) if self.n_bins_i32[-2] > 1 else 0
                parts += [torch.roll(r, -s2, dims=1), torch.roll(r, -(s2 + s1), dims=1)]
            self.templates = torch.stack(parts, dim=2).contiguous()
            self.row_stride, self.bin_stride = len(parts) * self.n_template_bins, len(parts)
        else:
            self.templates = self.templates_rows
            self.row_stride, self.bin_stride = self.n_template_bins, 1
        self.n_events = 0
        self.n_datasets = 0
        self.ev_bin = None
        self.peer_gather = None       # distributed.PeerGather (sharded evaluations), see UnbinnedEngine
        self.peer_mode = 'gather'
        self._x_slots = {}            # n -> [(pinned tensor, its ndarray)]: rotating landing buffers of gathered results
        self.last_gathered_owned = True
        self.last_gathered = None
        self.last_total = None
        self._graphs = {}             # CUDA graphs of repeated evaluations, per (schedule, batch shape)
        self._toy_schedule = None

    # -- datasets ---------------------------------------------------------------------------------
    def set_datasets(self, coords, offsets=None):
        """coords: [n_space, N_total] event coordinates (NumPy array or device tensor) of all datasets back to
        back; offsets: [T + 1] first event of each dataset (default: one dataset).  Runs the event preparation
        kernel once (low-corner bin + fractions per event)."""
        torch = self.torch
        if isinstance(coords, torch.Tensor):
            coords_dev = coords.to(self.device, torch.float64).contiguous()
        else:
            host = np.ascontiguousarray(np.asarray(coords, dtype=np.float64)).reshape(self.n_space, -1)
            if self.method == _cabi.LOOKUP_LINEAR and np.isnan(host).any():
                raise ValueError("One of the requested xi is out of bounds in dimension 0")
            coords_dev = torch.from_numpy(host).to(self.device)
        coords_dev = coords_dev.reshape(self.n_space, -1)
        n = int(coords_dev.shape[1])
        if offsets is None:
            offsets = np.array([0, n], dtype=np.int64)
        offsets = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
        if offsets[0] != 0 or offsets[-1] != n or np.any(np.diff(offsets) < 0):
            raise ValueError("dataset offsets must rise from 0 to the number of events")
        self.n_events = n
        self.n_datasets = len(offsets) - 1
        self.offsets_host = offsets
        self.offsets = torch.from_numpy(offsets).to(self.device)
        self.n_super_host = (np.diff(offsets) + _cabi.SUPERBLOCK - 1) // _cabi.SUPERBLOCK      # per dataset
        self.ld_frac = round_up(max(n, 1), 2)                                # even: 16-byte aligned fraction pairs
        self.ev_bin = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
        linear = self.method == _cabi.LOOKUP_LINEAR
        self.ev_frac = torch.empty((self.n_space, self.ld_frac), dtype=torch.float64, device=self.device) if linear else None
        if n:
            _cabi.check(self.lib.bi_template_prepare_events(
                self.n_space, _cabi.host_ptr(self.n_bins_i32), _cabi.host_ptr(self.edges_concat), self.method,
                _cabi.dev_ptr(coords_dev), coords_dev.shape[1], n, _cabi.dev_ptr(self.ev_bin),
                _cabi.dev_ptr(self.ev_frac), self.ld_frac, self._stream()), "bi_template_prepare_events")
            self.launches += 1
        self.coords = coords_dev                                            # kept for full_output (ps)
        self._graphs = {}
        if self.mode == 'mixture' and n > 1:
            # events of a dataset in bin order: the template loads of a warp become uniform (plumbing, once per dataset)
            key = self.ev_bin[:n].to(torch.int64)
            if self.n_datasets > 1:
                sizes = torch.from_numpy(np.diff(offsets)).to(self.device)
                key = key + self.n_template_bins * torch.repeat_interleave(
                    torch.arange(self.n_datasets, device=self.device, dtype=torch.int64), sizes)
            order = torch.argsort(key, stable=True)
            self.ev_bin = self.ev_bin[:n][order].contiguous()
            if self.ev_frac is not None:
                sorted_frac = torch.empty_like(self.ev_frac)
                sorted_frac[:, :n] = self.ev_frac[:, :n][:, order]
                self.ev_frac = sorted_frac
            del key, order
        self._toy_schedule = None
        self._single_cache = {}
        return self

    # -- schedules (host) -------------------------------------------------------------------------
    def _upload_schedule(self, pair_point, pair_dataset, group_first, group_count, group_points):
        """Device copies of a pair schedule.  pair_*: [Q]; groups: consecutive pairs [first, first + count)."""
        torch = self.torch
        Q, n_groups = len(pair_point), len(group_first)
        n_super_pair = self.n_super_host[pair_dataset] if Q else np.zeros(0, dtype=np.int64)
        partial_offset = np.zeros(Q + 1, dtype=np.int64)
        np.cumsum(n_super_pair, out=partial_offset[1:])
        groups = np.zeros((n_groups, 4), dtype=np.int32)
        groups[:, 0], groups[:, 1] = group_first, group_count
        groups[:, 2] = pair_dataset[group_first] if n_groups else 0
        n_super_group = self.n_super_host[groups[:, 2]] if n_groups else np.zeros(0, dtype=np.int64)
        if self.mode == 'mixture':                                           # K5b: a unit is a PAIR of superblocks
            n_super_group = (n_super_group + 1) // 2
        unit_offset = np.zeros(n_groups + 1, dtype=np.int64)
        np.cumsum(n_super_group, out=unit_offset[1:])
        n_units = int(unit_offset[-1])
        sched = dict(n_pairs=Q, n_groups=n_groups, n_units=n_units, group_points=group_points,
                     n_partials=int(partial_offset[-1]), max_partials=int(n_super_pair.max()) if Q else 0,
                     pair_point=torch.from_numpy(np.ascontiguousarray(pair_point, dtype=np.int32)).to(self.device),
                     partial_offset=torch.from_numpy(partial_offset).to(self.device),
                     groups=torch.from_numpy(groups).to(self.device),
                     unit_offset=torch.from_numpy(unit_offset).to(self.device), unit_group=None)
        if n_groups > 4096 and n_units < (1 << 31):
            # many small groups (toys): a direct unit -> group table instead of a binary search per unit
            sched["unit_group"] = torch.repeat_interleave(
                torch.arange(n_groups, dtype=torch.int32, device=self.device),
                torch.from_numpy(n_super_group).to(self.device))
        sched["h2d_bytes"] = 4 * Q + 8 * (Q + 1) + 16 * n_groups + 8 * (n_groups + 1)
        return sched

    def toy_schedule(self):
        """Pair t = (dataset t, point t), one pair per group; built once per set_datasets.  Large toy sets also get the
        bin-major event order of bi_template_ll_toys_bm (key "bm")."""
        if self._toy_schedule is None:
            T = self.n_datasets
            idx = np.arange(T, dtype=np.int64)
            self._toy_schedule = self._upload_schedule(idx, idx, idx, np.ones(T, dtype=np.int64), 1)
            self._toy_schedule["bm"] = self._bin_major_state()
        return self._toy_schedule

    def _bin_major_state(self):
        """Plumbing of the bin-major toy sweep (K5c, bi_template_bm.cu), once per toy set: the events of all toys sorted
        by their low-corner bin (with their toy, their position in toy order and their lookup fractions), the task list
        (bin, first event, <= 2048 events) and a bin-major copy of the packed templates.  None when the shape is not
        served or the toy set is too small for the per-bin set-up to pay (BI_TS_BM=1 / 0 forces / forbids it)."""
        torch = self.torch
        env = os.environ.get('BI_TS_BM')
        n, T = self.n_events, self.n_datasets
        if env == '0' or self.mode != 'exact' or self.method != _cabi.LOOKUP_LINEAR or n == 0 or T < 2:
            return None
        if env != '1' and (n < _BM_MIN_EVENTS or T < 1024):
            return None
        if not self.lib.bi_template_bm_supported(self.n_space, self.method, self.grid.n_dims,
                                                 _cabi.host_ptr(self.grid.n_anchors_i32), self.n_sources, self.n_rows):
            return None
        chunk = int(self.lib.bi_template_bm_chunk())
        key = self.ev_bin[:n]
        order = torch.argsort(key, stable=True)
        sizes = torch.from_numpy(np.diff(self.offsets_host)).to(self.device)
        toy_of_event = torch.repeat_interleave(torch.arange(T, device=self.device, dtype=torch.int32), sizes)
        counts = torch.bincount(key.to(torch.int64), minlength=self.n_template_bins).cpu().numpy().astype(np.int64)
        first = np.concatenate([[0], np.cumsum(counts)[:-1]])
        n_chunks = (counts + chunk - 1) // chunk
        task_bin = np.repeat(np.arange(self.n_template_bins, dtype=np.int64), n_chunks)
        within = np.arange(len(task_bin), dtype=np.int64) - np.repeat(np.cumsum(n_chunks) - n_chunks, n_chunks)
        task_start = first[task_bin] + within * chunk
        task_count = np.minimum(counts[task_bin] - within * chunk, chunk)
        big_first = np.argsort(-task_count, kind='stable')               # long tasks first: the static round-robin balances
        if getattr(self, "_templates_bm", None) is None:
            self._templates_bm = self.templates.reshape(self.n_rows, self.n_template_bins, self.bin_stride) \
                .permute(1, 0, 2).contiguous()
        return dict(n_tasks=len(task_bin),
                    task_bin=torch.from_numpy(task_bin[big_first].astype(np.int32)).to(self.device),
                    task_start=torch.from_numpy(task_start[big_first]).to(self.device),
                    task_count=torch.from_numpy(task_count[big_first].astype(np.int32)).to(self.device),
                    toy=toy_of_event[order].contiguous(), src=order.to(torch.int32).contiguous(),
                    frac=self.ev_frac[:, :n][:, order].contiguous(), ld=n,
                    record_doubles=int(self.lib.bi_template_bm_record_doubles(self.grid.n_dims, self.n_sources)))

    def _cells_for_grouping(self, zs):
        """Host copy of the hypercube cell of every point (-1: out of range, evaluated alone and skipped on the
        device).  The mixture form may group any points, so all in-range points share the key 0 there."""
        P = len(zs)
        if not self.grid.n_dims:
            return np.zeros(P, dtype=np.int64)
        ok = self.grid.in_range(zs)
        if self.mode == 'mixture':
            return np.where(ok, 0, -1).astype(np.int64)
        return np.where(ok, self.grid.cell_ids(np.where(ok[:, None], zs, self.grid.axes_concat[0])), -1)

    def pair_schedule(self, dataset_index, zs):
        """Pairs (dataset_index[q], point q): sorted by (dataset, hypercube cell) and cut into groups of at most
        TS_GROUP_POINTS pairs that share both (the template values of an event are gathered once per group).
        Returns (schedule, order): pair position j evaluates point order[j]."""
        P = len(zs)
        dataset_index = np.asarray(dataset_index, dtype=np.int64).reshape(P)
        if P and (dataset_index.min() < 0 or dataset_index.max() >= self.n_datasets):
            raise ValueError("dataset index outside [0, %d)" % self.n_datasets)
        cells = self._cells_for_grouping(zs)
        n_cells = int(np.prod(self.grid.cells_per_dim)) if self.grid.n_dims else 1
        if P <= 1:
            np_max = 1
        elif self.mode != 'mixture':
            np_max = _cabi.TS_GROUP_POINTS
        else:
            np_max = mixture_group_width(P, int(self.n_super_host[dataset_index].min()),
                                         getattr(self, 'mix_wide_min_superblocks', None),
                                         os.environ.get('BI_MIX_MMA') != '0')
        order, first, count = group_pairs(dataset_index, cells, n_cells, np_max)
        if P and count.max() == 1:
            np_max = 1
        elif P and self.mode == 'mixture' and count.max() <= _cabi.MIX_GROUP_POINTS:
            np_max = _cabi.MIX_GROUP_POINTS
        sched = self._upload_schedule(order, dataset_index[order], first, count, np_max)
        return sched, order

    def single_schedule(self, zs, dataset=0):
        """All P points on one dataset (cached while the points stay in their cells)."""
        P = len(zs)
        key = (P, dataset, self._cells_for_grouping(zs).tobytes())
        hit = self._single_cache.get(key)
        if hit is None:
            hit = self.pair_schedule(np.full(P, dataset, dtype=np.int64), zs)
            if len(self._single_cache) >= 8:
                self._single_cache.clear()
            self._single_cache[key] = hit
        return hit

    # -- evaluation -------------------------------------------------------------------------------
    def _setup_terms(self, P, zs_d, mult_d, scale_d, eff_d):
        """K1 with the contraction terms (rows, coefficients) K5 consumes."""
        torch = self.torch
        D, S, C, K = self.grid.n_dims, self.n_sources, self.grid.n_corners, self.n_terms
        o = dict(cell=self.ws.get("cell", P * max(D, 1), torch.int32), frac=self.ws.get("frac", P * max(D, 1), torch.float64),
                 corner=self.ws.get("corner", P * C, torch.int32), weight=self.ws.get("weight", P * C, torch.float64),
                 mus=self.ws.get("mus", P * S, torch.float64), musum=self.ws.get("musum", P, torch.float64),
                 status=self.ws.get("status", P, torch.int32), row=self.ws.get("row", P * K, torch.int32),
                 coef=self.ws.get("coef", P * K, torch.float64), wterm=self.ws.get("wterm", P * K, torch.float64),
                 term_source=self.ws.get("term_source", K, torch.int32))
        _cabi.check(self.lib.bi_point_setup(
            D, _cabi.host_ptr(self.grid.n_anchors_i32), _cabi.host_ptr(self.grid.axes_concat), S, P,
            _cabi.dev_ptr(zs_d), _cabi.dev_ptr(mult_d), _cabi.dev_ptr(scale_d), _cabi.dev_ptr(eff_d),
            _cabi.dev_ptr(self.mus_anchor), _cabi.host_ptr(self.allow_negative),
            _cabi.dev_ptr(o["cell"]), _cabi.dev_ptr(o["frac"]), _cabi.dev_ptr(o["corner"]), _cabi.dev_ptr(o["weight"]),
            _cabi.dev_ptr(o["mus"]), _cabi.dev_ptr(o["musum"]), _cabi.dev_ptr(o["status"]), _cabi.dev_ptr(o["row"]),
            _cabi.dev_ptr(o["coef"]), _cabi.dev_ptr(o["wterm"]), _cabi.dev_ptr(o["term_source"]), self._stream()),
            "bi_point_setup")
        self.launches += 1
        return o

    def run_schedule(self, sched, o):
        """K5 + ragged finalize on device-resident K1 outputs.  Returns (logl [Q], logsum [Q]) device tensors in
        PAIR order."""
        torch = self.torch
        Q = sched["n_pairs"]
        partial = self.ws.get("ts_partial", sched["n_partials"], torch.float64)
        logl = self.ws.get("ts_logl", Q, torch.float64)
        logsum = self.ws.get("ts_logsum", Q, torch.float64)
        if sched["n_units"] and self.mode == 'mixture':
            tmix = self.ws.get("ts_tmix", 4 * Q * self.n_template_bins, torch.float64)     # packed: <= 4 doubles per bin
            _cabi.check(self.lib.bi_template_mix(
                _cabi.dev_ptr(self.templates), self.row_stride, self.bin_stride, self.n_space,
                _cabi.host_ptr(self.n_bins_i32), self.method, self.n_terms,
                _cabi.dev_ptr(o["row"]), _cabi.dev_ptr(o["coef"]), _cabi.dev_ptr(o["status"]),
                _cabi.dev_ptr(sched["pair_point"]), Q, _cabi.dev_ptr(tmix), self._stream()), "bi_template_mix")
            _cabi.check(self.lib.bi_mixture_partials(
                _cabi.dev_ptr(tmix), self.n_space, _cabi.host_ptr(self.n_bins_i32), self.method,
                _cabi.dev_ptr(self.ev_bin), _cabi.dev_ptr(self.ev_frac), self.ld_frac, _cabi.dev_ptr(self.offsets),
                _cabi.dev_ptr(o["status"]), sched["n_groups"], sched["group_points"], _cabi.dev_ptr(sched["groups"]),
                _cabi.dev_ptr(sched["unit_offset"]), _cabi.dev_ptr(sched["unit_group"]), sched["n_units"],
                _cabi.dev_ptr(sched["pair_point"]), _cabi.dev_ptr(sched["partial_offset"]),
                self.outlier_likelihood, _cabi.dev_ptr(partial), self._stream()), "bi_mixture_partials")
            self.launches += 2
        elif sched["n_units"]:
            _cabi.check(self.lib.bi_template_partials(
                _cabi.dev_ptr(self.templates), self.row_stride, self.bin_stride, self.n_space,
                _cabi.host_ptr(self.n_bins_i32), self.method, _cabi.dev_ptr(self.ev_bin), _cabi.dev_ptr(self.ev_frac),
                self.ld_frac, _cabi.dev_ptr(self.offsets), self.n_terms, self.n_sources, _cabi.dev_ptr(o["row"]),
                _cabi.dev_ptr(o["coef"]), _cabi.dev_ptr(o["wterm"]), _cabi.dev_ptr(o["term_source"]),
                _cabi.dev_ptr(o["mus"]), _cabi.dev_ptr(o["status"]), sched["n_groups"], sched["group_points"],
                _cabi.dev_ptr(sched["groups"]), _cabi.dev_ptr(sched["unit_offset"]), _cabi.dev_ptr(sched["unit_group"]),
                sched["n_units"], _cabi.dev_ptr(sched["pair_point"]), _cabi.dev_ptr(sched["partial_offset"]),
                self.outlier_likelihood, _cabi.dev_ptr(partial), self._stream()), "bi_template_partials")
            self.launches += 1
        _cabi.check(self.lib.bi_template_finalize(
            _cabi.dev_ptr(partial), _cabi.dev_ptr(sched["partial_offset"]), _cabi.dev_ptr(sched["pair_point"]),
            _cabi.dev_ptr(o["musum"]), _cabi.dev_ptr(o["status"]), Q, sched["max_partials"], _cabi.dev_ptr(logl),
            _cabi.dev_ptr(logsum), self._stream()), "bi_template_finalize")
        self.launches += 1
        return logl, logsum

    def run_one_call(self, P, sched, zs_d, mult_d, scale_d, eff_d, logl_out=None, logsum_out=None):
        """ONE C-ABI call (bi_template_ll_batch): K1 -> (template morph) -> K5 / K5b -> ragged finalize, launched back to
        back.  Returns (dict with musum / status device tensors, logl [Q], logsum [Q]) in PAIR order.  The point inputs and
        logl_out / logsum_out may be pinned host tensors (the kernels then read / write them over PCIe directly)."""
        torch = self.torch
        Q = sched["n_pairs"]
        mixture = int(self.mode == 'mixture')
        nbytes = int(self.lib.bi_template_workspace_bytes(self.grid.n_dims, self.n_sources, P, sched["n_partials"], Q,
                                                          self.n_template_bins, mixture))
        ws = self.ws.get("ts_ws", nbytes, torch.uint8)
        o = dict(musum=self.ws.get("musum", P, torch.float64), status=self.ws.get("status", P, torch.int32))
        logl = self.ws.get("ts_logl", Q, torch.float64) if logl_out is None else logl_out
        logsum = self.ws.get("ts_logsum", Q, torch.float64) if logsum_out is None else logsum_out
        templates = self.templates_rows if mixture else self.templates
        row_stride, bin_stride = (self.n_template_bins, 1) if mixture else (self.row_stride, self.bin_stride)
        bm = sched.get("bm")
        if bm is not None and sched["n_units"]:
            # toy sweep with the densities formed bin-major (K5c): K1, records, densities, tree, finalize
            record = self.ws.get("bm_record", P * bm["record_doubles"] + 8, torch.float64)
            density = self.ws.get("bm_density", self.n_events, torch.float64)
            _cabi.check(self.lib.bi_template_ll_toys_bm(
                self.grid.n_dims, _cabi.host_ptr(self.grid.n_anchors_i32), _cabi.host_ptr(self.grid.axes_concat),
                self.n_sources, P, _cabi.dev_ptr(zs_d), _cabi.dev_ptr(mult_d), _cabi.dev_ptr(scale_d), _cabi.dev_ptr(eff_d),
                _cabi.dev_ptr(self.mus_anchor), _cabi.host_ptr(self.allow_negative),
                _cabi.dev_ptr(templates), row_stride, bin_stride, _cabi.dev_ptr(self._templates_bm), self.n_rows,
                self.n_space, _cabi.host_ptr(self.n_bins_i32),
                _cabi.dev_ptr(self.ev_bin), _cabi.dev_ptr(self.ev_frac), self.ld_frac, _cabi.dev_ptr(self.offsets),
                _cabi.dev_ptr(bm["task_bin"]), _cabi.dev_ptr(bm["task_start"]), _cabi.dev_ptr(bm["task_count"]),
                bm["n_tasks"], _cabi.dev_ptr(bm["toy"]), _cabi.dev_ptr(bm["src"]), _cabi.dev_ptr(bm["frac"]), bm["ld"],
                sched["n_groups"], _cabi.dev_ptr(sched["groups"]), _cabi.dev_ptr(sched["unit_offset"]),
                _cabi.dev_ptr(sched["unit_group"]), sched["n_units"], _cabi.dev_ptr(sched["pair_point"]),
                _cabi.dev_ptr(sched["partial_offset"]), sched["n_partials"], sched["max_partials"],
                self.outlier_likelihood, _cabi.dev_ptr(ws), ws.numel(), _cabi.dev_ptr(record), _cabi.dev_ptr(density),
                _cabi.dev_ptr(logl), _cabi.dev_ptr(logsum), _cabi.dev_ptr(o["musum"]), _cabi.dev_ptr(o["status"]),
                self._stream()), "bi_template_ll_toys_bm")
            self.launches += 5
            return o, logl, logsum
        _cabi.check(self.lib.bi_template_ll_batch(
            self.grid.n_dims, _cabi.host_ptr(self.grid.n_anchors_i32), _cabi.host_ptr(self.grid.axes_concat),
            self.n_sources, P, _cabi.dev_ptr(zs_d), _cabi.dev_ptr(mult_d), _cabi.dev_ptr(scale_d), _cabi.dev_ptr(eff_d),
            _cabi.dev_ptr(self.mus_anchor), _cabi.host_ptr(self.allow_negative),
            _cabi.dev_ptr(templates), row_stride, bin_stride, self.n_space, _cabi.host_ptr(self.n_bins_i32), self.method,
            mixture, _cabi.dev_ptr(self.ev_bin), _cabi.dev_ptr(self.ev_frac), self.ld_frac, _cabi.dev_ptr(self.offsets),
            sched["n_groups"], sched["group_points"], _cabi.dev_ptr(sched["groups"]), _cabi.dev_ptr(sched["unit_offset"]),
            _cabi.dev_ptr(sched["unit_group"]), sched["n_units"], _cabi.dev_ptr(sched["pair_point"]),
            _cabi.dev_ptr(sched["partial_offset"]), Q, sched["n_partials"], sched["max_partials"],
            self.outlier_likelihood, _cabi.dev_ptr(ws), ws.numel(), _cabi.dev_ptr(logl), _cabi.dev_ptr(logsum),
            _cabi.dev_ptr(o["musum"]), _cabi.dev_ptr(o["status"]), self._stream()), "bi_template_ll_batch")
        self.launches += (2 + (2 if mixture else 1)) if sched["n_units"] else 2
        return o, logl, logsum

    def mixture_kernel_only(self, sched, o):
        """bi_mixture_partials alone on the mixture templates the last run_schedule left in the workspace
        (bench / profiling: the HBM-bound kernel without K1, the template morph and the finalize)."""
        torch = self.torch
        Q = sched["n_pairs"]
        tmix = self.ws.get("ts_tmix", 4 * Q * self.n_template_bins, torch.float64)
        partial = self.ws.get("ts_partial", sched["n_partials"], torch.float64)
        _cabi.check(self.lib.bi_mixture_partials(
            _cabi.dev_ptr(tmix), self.n_space, _cabi.host_ptr(self.n_bins_i32), self.method,
            _cabi.dev_ptr(self.ev_bin), _cabi.dev_ptr(self.ev_frac), self.ld_frac, _cabi.dev_ptr(self.offsets),
            _cabi.dev_ptr(o["status"]), sched["n_groups"], sched["group_points"], _cabi.dev_ptr(sched["groups"]),
            _cabi.dev_ptr(sched["unit_offset"]), _cabi.dev_ptr(sched["unit_group"]), sched["n_units"],
            _cabi.dev_ptr(sched["pair_point"]), _cabi.dev_ptr(sched["partial_offset"]),
            self.outlier_likelihood, _cabi.dev_ptr(partial), self._stream()), "bi_mixture_partials")

    def _evaluate(self, sched, order, zs, mult, scale, eff, return_status, return_parts):
        torch = self.torch
        P = len(mult)
        if self.ev_bin is None:
            raise RuntimeError("set_datasets must be called first")
        zs = np.asarray(zs, dtype=np.float64).reshape(P, self.grid.n_dims)
        n_f = 3 * P if return_parts else P
        # ---- stage the inputs in pinned memory (host work only)
        D, S = self.grid.n_dims, self.n_sources
        sizes = (P * D, P * S, P if scale is not None else 0, P * S if eff is not None else 0)
        total = sum(sizes)
        pin = self.ws.get("h2d", total, torch.float64, pinned=True)
        pin_np = pin.numpy()
        off = 0
        for arr, size in zip((zs, mult, scale, eff), sizes):
            if size:
                pin_np[off:off + size] = np.asarray(arr, dtype=np.float64).reshape(-1)
                off += size
        dev = self.ws.get("points_in", total, torch.float64)
        views, off = [], 0
        for size in sizes:
            views.append(dev[off:off + size] if size else None)
            off += size
        nbytes = total * 8
        out_pin = self.ws.get("d2h", 3 * P, torch.float64, pinned=True)
        st_pin = self.ws.get("d2h_status", P, torch.int32, pinned=True)
        pg, mode = self.peer_gather, self.peer_mode
        n_x = 0 if pg is None else (P if mode == 'sum' else pg.world * pg.n)
        g_pin, x_slot = None, 0
        self.last_gathered = self.last_total = None                      # (releases the views of the previous call)
        if pg is not None:
            # landing buffer of the exchange result: a gather rotates over pinned buffers that are handed to the caller
            # without a host copy (reused only when no array or view of an earlier result is alive, see
            # UnbinnedEngine.batch_runner); sums, permuted results and the NCCL fallback copy out of a scratch buffer
            if mode != 'sum' and pg.fallback is None and order is None:
                pool = self._x_slots.setdefault(n_x, [])
                for k, (t, a) in enumerate(pool):
                    if sys.getrefcount(a) == 3:                          # the pool, the loop variable, this call
                        g_pin, x_slot = t, k + 1
                        break
                else:
                    if len(pool) < _X_SLOTS:
                        t = torch.empty(max(n_x, 1), dtype=torch.float64, pin_memory=True)
                        pool.append((t, t.numpy()))
                        g_pin, x_slot = t, len(pool)
            if g_pin is None:
                g_pin = self.ws.get("d2h_gather", n_x, torch.float64, pinned=True)
        state = {}

        # small batches: K1 reads the staged points from pinned host memory and (unsharded) the finalize kernel writes
        # logl / logsum there -- no copy node for them
        direct_in = _DIRECT_IO and 0 < nbytes <= _DIRECT_IO_MAX_BYTES and sched["n_pairs"] == P
        direct_out = direct_in and pg is None
        if direct_in:
            views, off = [], 0
            for size in sizes:
                views.append(pin[off:off + size] if size else None)
                off += size

        def device_sequence():
            """H2D of the staged inputs, bi_template_ll_batch, the exchange step of a sharded evaluation, D2H of the
            results: everything the device does."""
            if not direct_in:
                dev.copy_(pin, non_blocking=True)
            if direct_out:
                o, logl, logsum = self.run_one_call(P, sched, views[0], views[1], views[2], views[3],
                                                    out_pin[:P], out_pin[P:2 * P])
                if return_parts:
                    out_pin[2 * P:].copy_(o["musum"], non_blocking=True)
                st_pin.copy_(o["status"], non_blocking=True)
                state["o"], state["logl"], state["logsum"] = o, logl, logsum
                return
            o, logl, logsum = self.run_one_call(P, sched, views[0], views[1], views[2], views[3])
            if pg is not None:
                # sharded evaluation, exchanged over NVLink peer memory by one launch: the rank-ordered sum of the
                # shards' log sums (event sharding; pair order) or the logl rows of all ranks (point / toy sharding)
                if mode == 'sum':                                    # the kernel writes to pinned host memory itself
                    pg.reduce(logsum[:P], out=g_pin)
                else:
                    pg.gather(logl[:P], out=g_pin)
            out_pin[:P].copy_(logl, non_blocking=True)
            if return_parts:
                out_pin[P:2 * P].copy_(logsum, non_blocking=True)
                out_pin[2 * P:].copy_(o["musum"], non_blocking=True)
            st_pin.copy_(o["status"], non_blocking=True)
            state["o"], state["logl"], state["logsum"] = o, logl, logsum

        # ---- repeated evaluations of one schedule replay the sequence as ONE CUDA graph (sharded evaluations too: the
        # exchange launch keeps its epoch on the device; every rank issues one exchange per call, eager or replayed)
        graph = None
        use_graphs = _E2E_GRAPHS and (pg is None or pg.fallback is None)
        if use_graphs:
            gkey = (id(sched), P, sizes, return_parts, None if pg is None else (id(pg), mode), x_slot)
            entry = self._graphs.get(gkey)
            if entry is None:
                if len(self._graphs) >= 8:
                    self._graphs.clear()
                entry = self._graphs[gkey] = {"calls": 0, "graph": None, "ptrs": None, "sched": sched}
            entry["calls"] += 1
            # capturing costs tens of ms once: only evaluations short enough for the host overhead to matter are replayed
            if entry["graph"] is None and entry["calls"] >= 3 and entry.get("last_s", 1.0) < 2e-3:
                try:                                                # the first calls size the workspace buffers eagerly
                    torch.cuda.current_stream(self.device).synchronize()
                    g = torch.cuda.CUDAGraph()
                    with capture_graph(torch, g):
                        device_sequence()
                    entry["graph"] = g
                    entry["ptrs"] = self.ws.version
                except Exception:
                    entry["graph"] = False
                    try:
                        torch.cuda.synchronize(self.device)
                    except Exception:
                        pass
            if entry["graph"] and entry["ptrs"] == self.ws.version:
                graph = entry["graph"]
            elif entry["graph"]:
                entry["graph"], entry["calls"] = None, 1             # a workspace buffer moved: capture again later
        t_start = _time.perf_counter()
        n_launch = ((2 + (2 if self.mode == 'mixture' else 1)) if sched["n_units"] else 2) + \
            (0 if pg is None or pg.fallback is not None else 1) + (2 if sched.get("bm") is not None and sched["n_units"] else 0)
        if graph is not None:
            graph.replay()
            self.launches += n_launch
        else:
            device_sequence()
            if pg is not None and pg.fallback is None:
                self.launches += 1
        torch.cuda.current_stream(self.device).synchronize()
        if use_graphs:
            self._graphs[gkey]["last_s"] = _time.perf_counter() - t_start
        self.last_h2d_bytes = nbytes
        self.last_d2h_bytes = n_f * 8 + P * 4 + n_x * 8
        res = out_pin.numpy()[:n_f].copy()                                  # only what this call delivered
        status = st_pin.numpy()[:P].copy()
        ll, ls = res[:P], (res[P:2 * P] if return_parts else None)
        inv = None
        if order is not None:                                               # pair order -> point order
            inv = np.empty(P, dtype=np.int64)
            inv[order] = np.arange(P)
            ll, ls = ll[inv], ls[inv] if return_parts else ls
        self.last_gathered_owned = True
        if pg is not None and mode == 'sum':
            total = g_pin.numpy()[:P].copy()                                # rank-ordered sum of the log sums, pair order
            total = total if inv is None else total[inv]
            if return_parts:                                                # (the mu sums arrive with the parts only)
                self.last_total = np.where(status != 0, -np.inf, -res[2 * P:3 * P] + total)
        elif pg is not None:
            if x_slot:
                self.last_gathered = self._x_slots[n_x][x_slot - 1][1][:n_x].reshape(pg.world, -1)      # handed over, no copy
            else:
                gathered = g_pin.numpy().reshape(pg.world, -1).copy()
                self.last_gathered = gathered if inv is None else \
                    np.concatenate([gathered[:, :P][:, inv], gathered[:, P:]], axis=1)
        if return_parts:
            return ls, res[2 * P:], status
        return (ll, status) if return_status else ll

    def evaluate(self, zs, mult, scale=None, eff=None, return_status=False, return_parts=False, dataset=0):
        """P points on ONE dataset (same contract as UnbinnedEngine.evaluate)."""
        P = len(mult)
        if P == 0:
            if return_parts:
                return np.zeros(0), np.zeros(0), np.zeros(0, dtype=np.int32)
            return (np.zeros(0), np.zeros(0, dtype=np.int32)) if return_status else np.zeros(0)
        zs = np.asarray(zs, dtype=np.float64).reshape(P, self.grid.n_dims)
        sched, order = self.single_schedule(zs, dataset)
        return self._evaluate(sched, order, zs, mult, scale, eff, return_status, return_parts)

    def evaluate_toys(self, zs, mult, scale=None, eff=None, return_status=False, return_parts=False):
        """Point t on dataset t for all T datasets (one parameter point per toy)."""
        if self.mode != 'exact':
            raise NotImplementedError("toys are evaluated by the exact template kernel: build the engine with mode='exact'")
        if len(mult) != self.n_datasets:
            raise ValueError("need one parameter point per dataset: got %d points for %d datasets"
                             % (len(mult), self.n_datasets))
        return self._evaluate(self.toy_schedule(), None, zs, mult, scale, eff, return_status, return_parts)

    def evaluate_pairs(self, dataset_index, zs, mult, scale=None, eff=None, return_status=False):
        """Point q on dataset dataset_index[q]: several points per toy in one pass (finite-difference batches and
        line searches of many toy fits in lock step)."""
        if self.mode != 'exact':
            raise NotImplementedError("toys are evaluated by the exact template kernel: build the engine with mode='exact'")
        P = len(mult)
        if P == 0:
            return (np.zeros(0), np.zeros(0, dtype=np.int32)) if return_status else np.zeros(0)
        zs = np.asarray(zs, dtype=np.float64).reshape(P, self.grid.n_dims)
        sched, order = self.pair_schedule(dataset_index, zs)
        return self._evaluate(sched, order, zs, mult, scale, eff, return_status, False)

    def ps(self, z_row, mult_row, scale=None, eff=None, dataset=0):
        """(mus [S], ps [S, N]) of one point on one dataset, reference operation order (full_output=True): K3 on
        the point's K template rows, then the corner-weighted sum of bi_unbinned_ps."""
        torch = self.torch
        S, C, K = self.n_sources, self.grid.n_corners, self.n_terms
        zs_d, mult_d, scale_d, eff_d, _ = self._upload_points(np.asarray(z_row, dtype=np.float64).reshape(1, -1),
                                                              np.asarray(mult_row, dtype=np.float64).reshape(1, -1),
                                                              None if scale is None else [scale],
                                                              None if eff is None else np.asarray(eff).reshape(1, -1))
        o = self._setup_terms(1, zs_d, mult_d, scale_d, eff_d)
        lo, hi = int(self.offsets_host[dataset]), int(self.offsets_host[dataset + 1])
        n = hi - lo
        ld = max(round_up(n, _LD_ALIGN), _LD_ALIGN)
        rows = self.templates_rows.index_select(0, o["row"][:K].to(torch.int64))          # [K, B], k = c * S + s
        coords = self.coords[:, lo:hi].contiguous()
        a_sel = torch.zeros((K, ld), dtype=torch.float64, device=self.device)
        if n:
            _cabi.check(self.lib.bi_hist_lookup(_cabi.dev_ptr(rows), K, self.n_space, _cabi.host_ptr(self.n_bins_i32),
                                                _cabi.host_ptr(self.edges_concat), _cabi.dev_ptr(coords), n, n,
                                                self.method, _cabi.dev_ptr(a_sel), ld, None, self._stream()),
                        "bi_hist_lookup")
        out = torch.empty((S, max(n, 1)), dtype=torch.float64, device=self.device)
        corner = torch.arange(C, dtype=torch.int32, device=self.device)
        _cabi.check(self.lib.bi_unbinned_ps(_cabi.dev_ptr(a_sel), ld, n, S, C, _cabi.dev_ptr(corner),
                                            _cabi.dev_ptr(o["weight"]), _cabi.dev_ptr(out), out.shape[1], self._stream()),
                    "bi_unbinned_ps")
        self.launches += 2
        return o["mus"][:S].cpu().numpy().copy(), out[:, :n].cpu().numpy()
