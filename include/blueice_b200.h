/*
 * blueice_b200 -- C-ABI of the B200-native likelihood-evaluation hot path.
 *
 * The reference (JelleAalbers/blueice v1.2.1) is pure Python and has NO FFI / plugin interface
 * for this path; the drop-in boundary is its Python class API.  The entry
 * points below are what a ctypes binding inside the reference would call at the three seams where
 * the reference already isolates the numerics.  Each one cites the reference code it replaces
 * (file:line relative to the reference checkout).  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - `*_dev` pointers are CUDA device pointers owned by the caller; `*_host` are host pointers that
 *     are read during the call only.  The library never allocates persistent memory and never frees
 *     caller memory.  Every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     re-entrant per stream, and keeps no global state besides a thread-local error string.
 *   - return value: 0 = ok, negative = bi_status error code; text via bi_last_error().
 *   - all floating point is IEEE float64; indices are int32 unless noted; sizes are int64.
 *
 * Canonical summation order (what makes results independent of batch shape, kernel choice, grid
 * size and GPU count per shard) is documented in DESIGN.md section 4 and implemented in
 * blueice_b200/csrc/bi_common.cuh:
 *   block      = 32 consecutive events  -> L_b  (adjacent-pair binary tree)
 *   superblock = 16 consecutive blocks  -> S_j  (sequential)
 *   total      = 256 strided lanes (j mod 256, sequential) + xor-butterfly over the lanes
 */
#ifndef BLUEICE_B200_H
#define BLUEICE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BI_ABI_VERSION 1
#define BI_MAX_DIMS 6          /* max shape parameters per morph grid (2^6 = 64 corners)          */
#define BI_MAX_SOURCES 64      /* max sources per model                                           */
#define BI_MAX_SPACE_DIMS 4    /* max analysis-space dimensions of a histogram template           */
#define BI_MAX_AXIS_POINTS 256 /* max total anchor values (sum over dims) in one grid             */
#define BI_EVENT_BLOCK 32      /* canonical block: events per tree-reduced block                  */
#define BI_SUPERBLOCK 512      /* canonical superblock: events per sequentially summed partial    */

typedef enum bi_status {
    BI_OK = 0,
    BI_ERR_INVALID_ARGUMENT = -1,
    BI_ERR_CUDA = -2,
    BI_ERR_UNSUPPORTED = -3
} bi_status;

/* point status flags written by bi_point_setup (bitmask) */
#define BI_POINT_OK 0
#define BI_POINT_OUT_OF_RANGE 1   /* some z outside [min anchor, max anchor] or NaN: logL = -inf  */
#define BI_POINT_UNPHYSICAL 2     /* rates fail the check of likelihood.py:397-415: -inf / error  */

/* lookup methods for bi_hist_lookup */
#define BI_LOOKUP_LINEAR 0        /* source.py:225-240 */
#define BI_LOOKUP_PIECEWISE 1     /* source.py:242-243 */

/* binned flags returned per point by bi_binned_ll_batch (bitmask) */
#define BI_BB_ROOT1_POSITIVE 1    /* reference would fail `assert np.all(A_bins_1 <= 0)`  (likelihood.py:649) */
#define BI_BB_NEGATIVE_A 2        /* reference would fail `assert np.all(0 <= A_bins)`    (likelihood.py:655) */

const char* bi_last_error(void);
int bi_abi_version(void);

/* Number of superblock partials for n_events events: ceil(n_events / BI_SUPERBLOCK). */
int64_t bi_num_superblocks(int64_t n_events);

/*
 * K1 -- anchor-grid morphing set-up for a batch of parameter points.
 *
 * Replaces: GridInterpolator.make_interpolator / RegularGridInterpolator index search and corner
 * weights (pdf_morphers.py:57-70; scipy/interpolate/_rgi.py:520-549 + find_indices), the `mus`
 * interpolation call (likelihood.py:355), rate / livetime / efficiency scaling (likelihood.py:366-393)
 * and the unphysical-rate test (likelihood.py:397-415).
 *
 *   n_dims            D >= 0 shape parameters (0: no morphing, one anchor)
 *   n_anchors_host    [D] anchors per dimension
 *   axes_host         concatenated sorted anchor z values, sum(n_anchors) doubles
 *   zs_dev            [P, D] shape-parameter values (row major)
 *   rate_mult_dev     [P, S] rate multipliers
 *   scale_dev         [P] livetime scale (livetime_days / base) or NULL (= not applied)
 *   eff_dev           [P, S] efficiency multipliers (1 where a source has none) or NULL
 *   mus_anchor_dev    [G, S] expected events at each anchor, anchors in C order (first dim slowest)
 *   allow_negative_host [S] bytes (likelihood.py:82) or NULL (= all false)
 * outputs
 *   cell_dev          [P, D] int32 lower cell index per dim (bit-exact vs scipy find_indices)
 *   frac_dev          [P, D] normalised distance per dim
 *   corner_dev        [P, C] int32 flat anchor index of each hypercube corner, C = 2^D, first dim slowest
 *   weight_dev        [P, C] corner weights, product taken in scipy's order
 *   mus_dev           [P, S] scaled expected events
 *   musum_dev         [P]    sum_s mus (numpy pairwise order)
 *   status_dev        [P]    BI_POINT_* flags
 * optional outputs (all four or none; NULL = skip) -- the contraction terms bi_unbinned_partials_mma consumes,
 * K = C * S terms per point, term k = c * S + s:
 *   row_dev           [P, K] int32 row of the [G * S, ld] anchor tensor (= corner_c * S + s)
 *   coef_dev          [P, K] fl(weight_c * mus_s)
 *   wterm_dev         [P, K] weight_c
 *   term_source_dev   [K]    int32 source index of term k (= k % S)
 */
int bi_point_setup(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                   int32_t n_sources, int64_t n_points,
                   const double* zs_dev, const double* rate_mult_dev, const double* scale_dev,
                   const double* eff_dev, const double* mus_anchor_dev,
                   const uint8_t* allow_negative_host,
                   int32_t* cell_dev, double* frac_dev, int32_t* corner_dev, double* weight_dev,
                   double* mus_dev, double* musum_dev, int32_t* status_dev,
                   int32_t* row_dev, double* coef_dev, double* wterm_dev, int32_t* term_source_dev, void* stream);

/*
 * K1, source-wise form (likelihood.py:113-145,210-240: `source_wise_interpolation`): source s is morphed
 * over its own sub-grid spanned by the shape parameters in dim_mask_host[s] (bit d = parameter d), C order,
 * by its own RegularGridInterpolator.  Rows of the row matrix / mus_rows_dev: row_base_host[s] + flat
 * sub-anchor index.  Contraction terms: (source s, corner c_s), s-major; K = bi_sourcewise_terms(...).
 * Outputs as bi_point_setup (no corner / weight: the per-term weights are wterm_dev).
 */
int32_t bi_sourcewise_terms(int32_t n_sources, const uint32_t* dim_mask_host);
int bi_point_setup_sourcewise(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                              int32_t n_sources, const uint32_t* dim_mask_host, const int32_t* row_base_host,
                              int64_t n_points,
                              const double* zs_dev, const double* rate_mult_dev, const double* scale_dev,
                              const double* eff_dev, const double* mus_rows_dev, const uint8_t* allow_negative_host,
                              int32_t* cell_dev, double* frac_dev, double* mus_dev, double* musum_dev,
                              int32_t* status_dev, int32_t* row_dev, double* coef_dev, double* wterm_dev,
                              int32_t* term_source_dev, void* stream);

/*
 * K2 -- fused morph + mixture density + log + reduce, unbinned (per-superblock partial sums).
 *
 * Replaces: the `ps` interpolation call (likelihood.py:356 -> pdf_morphers.py:70 ->
 * _rgi.py:520-549 on the [n1..nD, S, N] tensor) fused with extended_loglikelihood
 * (likelihood.py:678-690): p_i = nansum_s(mu_s * ps[s,i]); non-positive/NaN p_i -> outlier_likelihood
 * (if non-zero); sum_i log p_i.
 *
 *   ps_anchor_dev   [G, S, ld_events] per-event pdf values at each anchor; ld_events even, >= n_events
 *   corner/weight/mus/status: outputs of bi_point_setup
 *   partial_dev     [P, n_super] out: S_j partial log sums (n_super = bi_num_superblocks(n_events))
 *
 * bi_unbinned_partials_stream: lanes = events; any P, any n_sources, n_corners <= 32; HBM-bound when P is small.
 *   point_index_dev [n_points] the points to evaluate (NULL: 0..n_points-1)
 * It is the general path (more than BI_MMA_MAX_TERMS contraction terms or more than BI_PLAN_MAX_CELLS hypercube
 * cells) and produces partials bit-identical to bi_unbinned_partials_mma for the same point.
 */
int bi_unbinned_partials_stream(const double* ps_anchor_dev, int64_t ld_events, int64_t n_events,
                                int32_t n_sources, int32_t n_corners,
                                const int32_t* point_index_dev, int64_t n_points,
                                const int32_t* corner_dev, const double* weight_dev,
                                const double* mus_dev, const int32_t* status_dev,
                                double outlier_likelihood, double* partial_dev, void* stream);

/*
 * Device-side schedule for bi_unbinned_partials_mma (no host round trip per batch).  Buckets the
 * evaluable points (status == 0) by hypercube cell and cuts them into point groups of at most
 * `unit_points` = bi_mma_unit_points(n_terms, n_points) points.
 *   cell_dev / status_dev   outputs of bi_point_setup ([P, max(D, 1)] and [P])
 *   group_points_dev [P]    out: point indices, cell-major
 *   groups_dev [(P + 1) * 4] out: (first, count, range counter = 0, 0) per point group, 16-byte aligned
 *   header_dev [8]          out: n_groups, n_ranges, superblocks_per_range, n_groups * n_ranges, 0 (group ticket
 *                           counter), n_evaluable, 0, 0.
 *   target_units            aimed-at number of work units (a few per resident warp)
 *   full_units              0: sizes balanced inside a cell (e.g. 7+7+7+6+6 m-tiles); 1: full units + one remainder
 * Grids with more than BI_PLAN_MAX_CELLS hypercube cells are rejected (BI_ERR_INVALID_ARGUMENT).
 */
#define BI_PLAN_MAX_CELLS 16384
int64_t bi_plan_max_cells(void);
int bi_unbinned_plan(int32_t n_dims, const int32_t* n_anchors_host, int64_t n_points,
                     const int32_t* cell_dev, const int32_t* status_dev, int32_t unit_points,
                     int64_t n_events, int32_t target_units, int32_t full_units,
                     int32_t* group_points_dev, int32_t* groups_dev, int32_t* header_dev, void* stream);

/*
 * bi_unbinned_partials_mma: K2 on the FP64 tensor pipe (DMMA.8x8x4), persistent: every warp picks a point
 * group (ticket header_dev[4]) and fetches superblock ranges from that group's counter (groups_dev[4 g + 2])
 * until all header_dev[0] x header_dev[1] (group, range) pairs are done.
 * The density is the contraction f_i = sum_k coef[p, k] * rows[row[p, k], i] over n_terms terms
 * (bi_point_setup / bi_point_setup_sourcewise write row / coef / wterm / term_source).
 *   rows_dev   [n_rows, ld_events] per-event pdf values, one row per (anchor, source); ld_events even
 *   group_points_dev / groups_dev / header_dev as written by bi_unbinned_plan (or by the caller: all
 *   points of a group MUST share their row list, have status 0, and count <= bi_mma_unit_points(n_terms, n_points);
 *   header_dev[4] and every group's range counter must be 0 on entry).  Requires n_terms <= BI_MMA_MAX_TERMS.
 * grid_dims >= 0 declares the full-grid layout (rows_dev = [G][S][ld] anchor tensor over grid_dims shape
 * parameters with n_anchors_host anchors each, term k = corner * S + source, cell_dev = bi_point_setup's cells):
 * event tiles are then fetched with ONE tiled TMA instruction (tensor map over [ld][S][n_D]..[n_1], grid_dims <= 3)
 * instead of one bulk copy per row.  grid_dims = -1: arbitrary row lists (source-wise interpolation).
 * Densities that leave [2^-126, 2^127) (zero, negative, NaN, inf ...) are re-evaluated with the
 * reference's nansum / outlier semantics from wterm / term_source / mus (likelihood.py:686-689).
 * Contractions of more than 32 terms run the K-chunk form of the kernel (k_unbinned_mma_wide: a CTA's four warps share
 * event tiles whose rows arrive in chunks of 32 through a CTA-wide TMA ring, accumulators carried across the chunks:
 * the same sequential fma chain over k); units are then (group, range) pairs taken in launch order, no counters used.
 * That kernel reads the coefficients from a chunk-major copy in schedule order which the call writes first
 * (k_wide_pack_coef): coef_chunks_dev must hold bi_mma_coef_chunks_doubles(n_terms, n_points) doubles (0 for short
 * contractions: the pointer may then be NULL), 16-byte aligned; n_points = points of the batch (>= every index in
 * group_points_dev + 1 is not required: it sizes the copy, header_dev[5] <= n_points evaluable points are packed).
 */
#define BI_MMA_MAX_TERMS 4096
int bi_unbinned_partials_mma(const double* rows_dev, int64_t ld_events, int64_t n_events,
                             int32_t n_terms, int32_t n_sources,
                             const int32_t* group_points_dev, int32_t* groups_dev, int32_t* header_dev,
                             const int32_t* row_dev, const double* coef_dev, const double* wterm_dev,
                             const int32_t* term_source_dev, const double* mus_dev,
                             double outlier_likelihood, double* partial_dev,
                             int32_t grid_dims, const int32_t* n_anchors_host, const int32_t* cell_dev,
                             int64_t n_points, double* coef_chunks_dev, void* stream);
int32_t bi_mma_unit_points(int32_t n_terms, int64_t n_points);
int64_t bi_mma_coef_chunks_doubles(int32_t n_terms, int64_t n_points);

/*
 * The whole unbinned hot path in ONE call: K1 point set-up -> device-side schedule -> K2 (DMMA) ->
 * finalize.  Replaces LogLikelihoodBase.__call__'s numerics for a batch of P points
 * (likelihood.py:318-427 with UnbinnedLogLikelihood._compute_likelihood, :571-573).
 *   workspace_dev   bi_unbinned_workspace_bytes(n_dims, n_sources, n_terms, n_points, n_events) bytes,
 *                   256-byte aligned; n_terms = 2^n_dims * n_sources for bi_unbinned_ll_batch
 *   logl_dev [P], logsum_dev [P] (may be NULL), musum_dev [P], status_dev [P]: outputs
 */
int64_t bi_unbinned_workspace_bytes(int32_t n_dims, int32_t n_sources, int32_t n_terms, int64_t n_points,
                                    int64_t n_events);
/* byte offsets of the workspace regions: cell, frac, corner, weight, mus, partial, group_points, groups,
 * header, row, coef, wterm, term_source, coef_chunks, total (15 int64) -- lets a caller run / inspect the stages separately */
int bi_unbinned_workspace_layout(int32_t n_dims, int32_t n_sources, int32_t n_terms, int64_t n_points,
                                 int64_t n_events, int64_t* offsets_host);
int bi_unbinned_ll_batch(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                         int32_t n_sources, int64_t n_points,
                         const double* zs_dev, const double* rate_mult_dev, const double* scale_dev,
                         const double* eff_dev, const double* mus_anchor_dev, const uint8_t* allow_negative_host,
                         const double* ps_anchor_dev, int64_t ld_events, int64_t n_events,
                         double outlier_likelihood, int32_t target_units,
                         void* workspace_dev, int64_t workspace_bytes,
                         double* logl_dev, double* logsum_dev, double* musum_dev, int32_t* status_dev, void* stream);

/*
 * The same evaluation for a TINY batch in ONE launch (K1 + K2 + finalize fused: one CTA per point; the latency regime of
 * a minimiser or interval search calling ll(**params) thousands of times, inference.py:131-178,332-389).  Results are
 * bit-identical to bi_unbinned_ll_batch's four launches, which calls it by itself whenever bi_unbinned_small_ok(...)
 * says the batch qualifies (terms <= 128, events <= 8192, points x superblocks <= 2048; BI_SMALL=0 in the environment
 * disables that).  The point inputs (zs, rate_mult, scale, eff) and the outputs (logl, logsum, musum, status) only need
 * to be DEVICE-ACCESSIBLE: pinned host memory works, the kernel then reads / writes it directly and no copy surrounds
 * the launch.
 */
int32_t bi_unbinned_small_ok(int32_t n_dims, int32_t n_sources, int64_t n_points, int64_t n_events);
int bi_unbinned_ll_small(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                         int32_t n_sources, int64_t n_points,
                         const double* zs, const double* rate_mult, const double* scale, const double* eff,
                         const double* mus_anchor_dev, const uint8_t* allow_negative_host,
                         const double* ps_anchor_dev, int64_t ld_events, int64_t n_events,
                         double outlier_likelihood,
                         double* logl, double* logsum, double* musum, int32_t* status, void* stream);

/* bi_unbinned_ll_batch for source-wise interpolation (K1 = bi_point_setup_sourcewise). */
int bi_unbinned_ll_batch_sourcewise(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                                    int32_t n_sources, const uint32_t* dim_mask_host, const int32_t* row_base_host,
                                    int64_t n_points,
                                    const double* zs_dev, const double* rate_mult_dev, const double* scale_dev,
                                    const double* eff_dev, const double* mus_rows_dev,
                                    const uint8_t* allow_negative_host,
                                    const double* rows_dev, int64_t ld_events, int64_t n_events,
                                    double outlier_likelihood, int32_t target_units,
                                    void* workspace_dev, int64_t workspace_bytes,
                                    double* logl_dev, double* logsum_dev, double* musum_dev, int32_t* status_dev,
                                    void* stream);

/* Morphed per-event pdf values ps[S, N] of ONE point from its contraction terms (row_dev / wterm_dev point at
 * that point's K entries), reference operation order; copy_source_dev[s] != 0: the source has no shape
 * parameter and its single row is copied (full_output=True with source-wise interpolation). */
int bi_unbinned_ps_terms(const double* rows_dev, int64_t ld_events, int64_t n_events, int32_t n_terms,
                         int32_t n_sources, const int32_t* row_dev, const double* wterm_dev,
                         const int32_t* term_source_dev, const uint8_t* copy_source_dev,
                         double* ps_out_dev, int64_t ld_out, void* stream);

/* logL[p] = -musum[p] + total(partial[p, :]) in canonical order; status != 0 -> -inf.
 * (likelihood.py:690 `-mu.sum() + np.sum(np.log(p_events))`, :347/:402 soft failures.)
 * logsum_dev (may be NULL) receives total(partial[p, :]) alone: the per-shard term that is summed over
 * ranks when the events are sharded across GPUs. */
int bi_unbinned_finalize(const double* partial_dev, int64_t n_super, const double* musum_dev,
                         const int32_t* status_dev, int64_t n_points, double* logl_dev,
                         double* logsum_dev, void* stream);

/* Morphed per-event pdf values ps[S, N] for ONE point (full_output=True, likelihood.py:424-425). */
int bi_unbinned_ps(const double* ps_anchor_dev, int64_t ld_events, int64_t n_events,
                   int32_t n_sources, int32_t n_corners, const int32_t* corner_dev,
                   const double* weight_dev, double* ps_out_dev, int64_t ld_out, void* stream);

/*
 * K3 -- histogram-template lookup of N events in T templates that share their bin edges.
 *
 * Replaces: HistogramPdfSource.pdf (source.py:219-246) called G*S times by
 * UnbinnedLogLikelihood.set_data (likelihood.py:557-560 -> model.py:97-99).
 *
 *   templates_dev [T, prod(n_bins)] densities, C order
 *   n_space       analysis-space dimensions (1..BI_MAX_SPACE_DIMS)
 *   n_bins_host   [n_space]
 *   edges_host    concatenated bin edges, sum(n_bins + 1) doubles
 *   coords_dev    [n_space, ld_coords] event coordinates
 *   out_dev       [T, ld_out]
 * LINEAR: clip to the bin-centre range, multilinear over bin centres, scipy's operation order
 *         (2-D: evaluate_linear_2d fast path; otherwise the generic corner loop).
 * PIECEWISE: idx_d = clip(searchsorted(edges_d, x_d, 'left') - 1, 0, n_bins_d - 1).
 */
int bi_hist_lookup(const double* templates_dev, int64_t n_templates, int32_t n_space,
                   const int32_t* n_bins_host, const double* edges_host,
                   const double* coords_dev, int64_t ld_coords, int64_t n_events, int32_t method,
                   double* out_dev, int64_t ld_out, int32_t* bin_index_dev, void* stream);

/*
 * K5 -- template-space unbinned likelihood: fused template lookup + morph + mixture + log + reduce over many
 * datasets (toy Monte Carlos: SURVEY.md section 8d config 4) or one dataset too large for the dense anchor tensor
 * (config 5).  Replaces the same reference code as K3 + K2 (source.py:219-246 called by likelihood.py:557-562, then
 * likelihood.py:355-356,678-690) without materialising A[G, S, N]; every (dataset, point) pair evaluates
 * BIT-IDENTICALLY to bi_hist_lookup + bi_unbinned_ll_batch on that dataset alone.
 *
 * bi_template_prepare_events (once per dataset set): event coordinates -> low-corner bin + per-dimension fractions
 *   coords_dev [n_space, ld_coords]; ev_bin_dev [n_events] int32; ev_frac_dev [n_space, ld_frac] (linear only)
 *
 * bi_template_partials:
 *   templates_dev        element (row, bin) at templates_dev + row * row_stride + bin * bin_stride (in doubles).
 *                        PIECEWISE: the bin's value (plain [n_rows, n_bins]: row_stride = n_bins, bin_stride = 1).
 *                        LINEAR: a PACKED layout holding every bin together with its lookup neighbours, so that the
 *                        corners come with one wide gather (the kernel is bound by the number of L2 requests):
 *                        1-D  [n_rows, n_bins, 2] = (T[b], T[b + 1])                          (strides 2 n_bins, 2; 16-byte aligned)
 *                        >=2-D [n_rows, n_bins, 4] = (T[b], T[b + 1], T[b + s], T[b + s + 1]) (strides 4 n_bins, 4; 32-byte aligned),
 *                        s = stride of the second-last dimension (0 for a one-bin dimension, as the + 1)
 *   dataset_offset_dev   [n_datasets + 1] int64 first event of each dataset in ev_bin_dev / ev_frac_dev
 *   row/coef/wterm/term_source/mus/status: outputs of bi_point_setup[_sourcewise] for the points ([P, K], ..., [P])
 *   pair list            pair q evaluates point pair_point_dev[q] on one dataset; its superblock partials go to
 *                        partial_dev[pair_partial_offset_dev[q] + j], j < bi_num_superblocks(events of the dataset)
 *   groups_dev           [n_groups, 4] int32 (first pair, pair count <= group_points, dataset, 0): pairs of a group
 *                        are consecutive, share the dataset and MUST share their row list (same hypercube cell)
 *   group_points         1 or BI_TS_GROUP_POINTS (selects the kernel instantiation)
 *   unit_offset_dev      [n_groups + 1] int64 prefix sum of the groups' superblock counts; n_units = its last entry
 *   unit_group_dev       [n_units] int32 group of each unit, or NULL (binary search in unit_offset_dev)
 * Points with status != 0 are skipped (bi_template_finalize reports -inf).
 *
 * bi_template_finalize: logl[q] = -musum[point of q] + total(partials of pair q) in the canonical order of
 * bi_unbinned_finalize; pair_point_dev NULL = identity; pair_partial_offset_dev has n_pairs + 1 entries;
 * max_partials = the largest number of partials of one pair (a hint: picks one CTA or one warp per pair).
 */
#define BI_TS_MAX_TERMS 256
#define BI_TS_GROUP_POINTS 8
#define BI_MIX_GROUP_POINTS 8      /* points per group of bi_mixture_partials (K5b) */
#define BI_MIX_GROUP_POINTS_WIDE 16 /* two 8-point m-tiles per warp (tensor-pipe kernel only) */
int bi_template_prepare_events(int32_t n_space, const int32_t* n_bins_host, const double* edges_host,
                               int32_t method, const double* coords_dev, int64_t ld_coords, int64_t n_events,
                               int32_t* ev_bin_dev, double* ev_frac_dev, int64_t ld_frac, void* stream);
int bi_template_partials(const double* templates_dev, int64_t row_stride, int64_t bin_stride,
                         int32_t n_space, const int32_t* n_bins_host, int32_t method,
                         const int32_t* ev_bin_dev, const double* ev_frac_dev, int64_t ld_frac,
                         const int64_t* dataset_offset_dev, int32_t n_terms, int32_t n_sources,
                         const int32_t* row_dev, const double* coef_dev, const double* wterm_dev,
                         const int32_t* term_source_dev, const double* mus_dev, const int32_t* status_dev,
                         int64_t n_groups, int32_t group_points, const int32_t* groups_dev,
                         const int64_t* unit_offset_dev, const int32_t* unit_group_dev, int64_t n_units,
                         const int32_t* pair_point_dev, const int64_t* pair_partial_offset_dev,
                         double outlier_likelihood, double* partial_dev, void* stream);
int bi_template_finalize(const double* partial_dev, const int64_t* pair_partial_offset_dev,
                         const int32_t* pair_point_dev, const double* musum_dev, const int32_t* status_dev,
                         int64_t n_pairs, int64_t max_partials, double* logl_dev, double* logsum_dev, void* stream);

/*
 * K5b -- mixture form of the template-space likelihood: morph the TEMPLATES per point (bi_template_mix:
 * tmix[q, b] = fma chain over k of templates[row[q, k], b] * coef[q, k]), then look the events up in the mixture
 * template (bi_mixture_partials: f_i = lookup(tmix[q], x_i), scipy's operation order, then the canonical log-sum).
 * Equal to K5 / K2 in exact arithmetic (the lookup is linear in the template values); in float64 the operation order
 * differs by ~1e-16 relative per event -- inside the 1e-9 * N contract, NOT bit-identical to the other kernels.
 * One lookup per point-event instead of n_terms; with the events of a dataset sorted by bin the template loads are
 * warp-uniform and the kernel streams the prepared events (4 + 8 * n_space bytes each) at HBM speed.
 * Requires finite templates.  templates_dev of bi_template_mix: the plain [n_rows, prod(n_bins)] layout.
 * tmix_dev: [n_pairs, prod(n_bins), pack] (32-byte aligned), row q belongs to pair q, PACKED like K5's templates:
 * pack = 1 (piecewise), 2 (1-D linear: m[b], m[b + 1]) or 4 (m[b], m[b + 1], m[b + s], m[b + s + 1]); size it as
 * n_pairs * prod(n_bins) * 4 doubles.  Groups as in
 * bi_template_partials except that the points of a group need NOT share a hypercube cell and that a unit is a PAIR
 * of consecutive superblocks (one per half-warp): unit_offset_dev = prefix sum of ceil(superblocks / 2) per group;
 * group_points is 1, BI_MIX_GROUP_POINTS or BI_MIX_GROUP_POINTS_WIDE.  Groups of more than one point run on the FP64
 * tensor pipe (k_mixture_partials_mma: the lookup of 8 points x 8 events that share a bin is one DMMA.8x8x4 over the
 * lookup corners, bit-identical to the single-point kernel); BI_MIX_MMA=0 in the environment keeps the 8-point gather
 * kernel (A/B measurements; it does not take the wide groups).
 * The lookup uses pre-multiplied corner weights, r = fma chain over corners of V[c] * w_c (this form's own order).
 */
int bi_template_mix(const double* templates_dev, int64_t row_stride, int64_t bin_stride,
                    int32_t n_space, const int32_t* n_bins_host, int32_t method,
                    int32_t n_terms, const int32_t* row_dev, const double* coef_dev, const int32_t* status_dev,
                    const int32_t* pair_point_dev, int64_t n_pairs, double* tmix_dev, void* stream);
int bi_mixture_partials(const double* tmix_dev, int32_t n_space, const int32_t* n_bins_host, int32_t method,
                        const int32_t* ev_bin_dev, const double* ev_frac_dev, int64_t ld_frac,
                        const int64_t* dataset_offset_dev, const int32_t* status_dev,
                        int64_t n_groups, int32_t group_points, const int32_t* groups_dev,
                        const int64_t* unit_offset_dev, const int32_t* unit_group_dev, int64_t n_units,
                        const int32_t* pair_point_dev, const int64_t* pair_partial_offset_dev,
                        double outlier_likelihood, double* partial_dev, void* stream);

/*
 * The whole template-space evaluation in ONE call (the counterpart of bi_unbinned_ll_batch): K1 bi_point_setup ->
 * (mixture != 0: bi_template_mix) -> bi_template_partials / bi_mixture_partials -> bi_template_finalize, launched back
 * to back.  Arguments as in those entry points; n_partials = pair_partial_offset[n_pairs]; mixture != 0 selects K5b
 * (templates_dev then is the plain [n_rows, n_bins] layout).  workspace_dev: bi_template_workspace_bytes(...) bytes,
 * 256-byte aligned (K1 outputs, partials, mixture templates).  Outputs: logl_dev / logsum_dev [n_pairs] in PAIR order,
 * musum_dev / status_dev [n_points].
 */
int64_t bi_template_workspace_bytes(int32_t n_dims, int32_t n_sources, int64_t n_points, int64_t n_partials,
                                    int64_t n_pairs, int64_t n_template_bins, int32_t mixture);
int bi_template_ll_batch(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                         int32_t n_sources, int64_t n_points,
                         const double* zs_dev, const double* rate_mult_dev, const double* scale_dev,
                         const double* eff_dev, const double* mus_anchor_dev, const uint8_t* allow_negative_host,
                         const double* templates_dev, int64_t row_stride, int64_t bin_stride,
                         int32_t n_space, const int32_t* n_bins_host, int32_t method, int32_t mixture,
                         const int32_t* ev_bin_dev, const double* ev_frac_dev, int64_t ld_frac,
                         const int64_t* dataset_offset_dev,
                         int64_t n_groups, int32_t group_points, const int32_t* groups_dev,
                         const int64_t* unit_offset_dev, const int32_t* unit_group_dev, int64_t n_units,
                         const int32_t* pair_point_dev, const int64_t* pair_partial_offset_dev,
                         int64_t n_pairs, int64_t n_partials, int64_t max_partials,
                         double outlier_likelihood, void* workspace_dev, int64_t workspace_bytes,
                         double* logl_dev, double* logsum_dev, double* musum_dev, int32_t* status_dev, void* stream);

/*
 * K5c -- a toy sweep (ONE parameter point per dataset) with the densities formed BIN-MAJOR (bi_template_bm.cu).
 * Replaces, per toy, set_data -> Model.score_events -> HistogramPdfSource.pdf and the density of
 * extended_loglikelihood (likelihood.py:531-562,678-690, model.py:97-99, source.py:225-240), bit-identical to
 * bi_template_ll_batch on the same toys.
 *
 * Plumbing the caller prepares once per toy set (engine.TemplateUnbinnedEngine.set_datasets): the events of all toys sorted
 * by their low-corner bin -- bm_toy_dev / bm_src_dev [N] (dataset index = point index, and position in toy order, of the
 * bin-sorted event), bm_frac_dev [n_space, ld_bm] (lookup fractions, bin-sorted) -- and the task list: task t evaluates
 * task_count[t] <= 2048 events of bin task_bin[t] starting at bin-sorted position task_start[t].
 * templates_bm_dev: [prod(n_bins)][n_rows][pack] the packed templates of bi_template_partials, bin-major.
 *
 * bi_template_bm_supported     : 1 when the shape is served (linear lookup, 1-2 analysis dimensions, 1-4 shape
 *                                parameters, <= 8 sources, <= 4096 hypercube cells, rows of one bin <= 160 kB)
 * bi_template_bm_record_doubles: doubles per point record; record_dev holds n_points of them, 32-byte aligned
 * bi_template_bm_density       : density_dev[i] = f(theta_{toy(i)}, x_i) for every event i (toy order) of a toy whose
 *                                point has status 0; cell_dev / frac_dev / mus_dev / status_dev are K1's outputs
 * bi_template_ll_toys_bm       : the whole sweep in one call: K1 -> records -> densities -> range test + canonical tree
 *                                (bi_template_partials' kernel on the densities) -> bi_template_finalize.  Schedule
 *                                arguments: the one-pair-per-dataset schedule of bi_template_ll_batch; workspace_dev:
 *                                bi_template_workspace_bytes(..., mixture = 0).
 */
int bi_template_bm_supported(int32_t n_space, int32_t method, int32_t n_dims, const int32_t* n_anchors_host,
                             int32_t n_sources, int64_t n_rows);
int64_t bi_template_bm_record_doubles(int32_t n_dims, int32_t n_sources);
int32_t bi_template_bm_chunk(void);        /* events per task (the caller cuts the bins' event lists into tasks of this size) */
int bi_template_bm_density(const double* templates_bm_dev, int64_t n_rows, int32_t n_space,
                           int32_t n_dims, const int32_t* n_anchors_host, int32_t n_sources, int64_t n_points,
                           const int32_t* cell_dev, const double* frac_dev, const double* mus_dev,
                           const int32_t* status_dev,
                           const int32_t* task_bin_dev, const int64_t* task_start_dev,
                           const int32_t* task_count_dev, int64_t n_tasks,
                           const int32_t* bm_toy_dev, const int32_t* bm_src_dev, const double* bm_frac_dev,
                           int64_t ld_bm, double* record_dev, double* density_dev, void* stream);
int bi_template_ll_toys_bm(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                           int32_t n_sources, int64_t n_points,
                           const double* zs_dev, const double* rate_mult_dev, const double* scale_dev,
                           const double* eff_dev, const double* mus_anchor_dev, const uint8_t* allow_negative_host,
                           const double* templates_dev, int64_t row_stride, int64_t bin_stride,
                           const double* templates_bm_dev, int64_t n_rows,
                           int32_t n_space, const int32_t* n_bins_host,
                           const int32_t* ev_bin_dev, const double* ev_frac_dev, int64_t ld_frac,
                           const int64_t* dataset_offset_dev,
                           const int32_t* task_bin_dev, const int64_t* task_start_dev,
                           const int32_t* task_count_dev, int64_t n_tasks,
                           const int32_t* bm_toy_dev, const int32_t* bm_src_dev, const double* bm_frac_dev,
                           int64_t ld_bm,
                           int64_t n_groups, const int32_t* groups_dev, const int64_t* unit_offset_dev,
                           const int32_t* unit_group_dev, int64_t n_units, const int32_t* pair_point_dev,
                           const int64_t* pair_partial_offset_dev, int64_t n_partials, int64_t max_partials,
                           double outlier_likelihood, void* workspace_dev, int64_t workspace_bytes,
                           double* record_dev, double* density_dev,
                           double* logl_dev, double* logsum_dev, double* musum_dev, int32_t* status_dev, void* stream);

/*
 * On-device toy Monte Carlo generation.
 *
 * Replaces Model.simulate (model.py:69-91) for models whose sources are histogram templates:
 * HistogramPdfSource.simulate (source.py:248-264) -> multihist Histdd.get_random.  Philox4x32-10 keyed by `seed`
 * with the counter (index, toy id, domain): toy `toy_id0 + t` is the same events whatever the batch or the number
 * of GPUs.  Parity with the reference is distributional (its Mersenne-Twister stream is not reproduced).
 *
 * bi_toy_counts : counts_dev[t, s] ~ Poisson(mus_dev[t, s]) (mus_per_toy != 0) or Poisson(mus_dev[s])
 * bi_toy_events : the events of all toys; offsets_dev [n_toys + 1] = prefix sum of the toys' total counts (caller),
 *                 n_events = its last entry; events of a toy are grouped by source in source order (model.py:88).
 *   cdf_dev     [S, prod(n_bins)] per source: cumsum(pmf.ravel()) / sum  (Histdd.get_random's table)
 *   coords_dev  [n_space, ld_coords] out; source_dev [n_events] out (may be NULL): the record array's 'source' field
 */
int bi_toy_counts(int32_t n_sources, int64_t n_toys, int64_t toy_id0, const double* mus_dev, int32_t mus_per_toy,
                  uint64_t seed, int32_t* counts_dev, void* stream);
int bi_toy_events(int32_t n_space, const int32_t* n_bins_host, const double* edges_host, int32_t n_sources,
                  const double* cdf_dev, int64_t n_toys, int64_t toy_id0, const int32_t* counts_dev,
                  const int64_t* offsets_dev, int64_t n_events, uint64_t seed, double* coords_dev, int64_t ld_coords,
                  int32_t* source_dev, void* stream);

/*
 * Event binning with np.histogramdd semantics (multihist.Histdd.add; likelihood.py:604-609).
 *   counts_dev [prod(n_bins)] uint64, ZEROED BY THE CALLER, incremented with integer atomics
 *   bin_index_dev [N] optional out: flat bin index or -1 for dropped events (may be NULL)
 */
int bi_histogramdd(int32_t n_space, const int32_t* n_bins_host, const double* edges_host,
                   const double* coords_dev, int64_t ld_coords, int64_t n_events,
                   unsigned long long* counts_dev, int32_t* bin_index_dev, void* stream);

/* bi_histogramdd for MANY datasets back to back (binned toys): dataset_offset_dev [n_datasets + 1] first event of each
 * dataset; counts_dev [n_datasets, ld_counts] uint64, zeroed by the caller. */
int bi_histogramdd_toys(int32_t n_space, const int32_t* n_bins_host, const double* edges_host,
                        const double* coords_dev, int64_t ld_coords, int64_t n_events,
                        const int64_t* dataset_offset_dev, int64_t n_datasets,
                        unsigned long long* counts_dev, int64_t ld_counts, void* stream);

/*
 * K4 -- binned Poisson log-likelihood with optional Beeston-Barlow adjustment, batch of points.
 *
 * Replaces: BinnedLogLikelihood.adjust_expectations (likelihood.py:618-660),
 * beeston_barlow_root1/2 (likelihood.py:693-712), _compute_likelihood (likelihood.py:662-675) and the
 * pmf / n_model_events interpolation calls (likelihood.py:356-357).
 *
 *   pmf_anchor_dev      [G, S, ld_bins] pmf per bin at each anchor
 *   n_model_anchor_dev  [G, ld_bins] calibration events per bin of source `bb_source` at each anchor
 *                       (NULL when bb_source < 0)
 *   n_model_sum_anchor_dev [G] sum over bins of n_model_anchor_dev per anchor (NULL when bb_source < 0);
 *                       n_model_events[source_i].sum() (likelihood.py:645) is taken as the morph of these
 *   observed_dev        [n_bins] observed counts (float64)
 *   lgamma_obs_dev      [n_bins] gammaln(observed + 1), computed once per dataset by the caller
 *   scratch_dev         bi_binned_scratch_doubles(P, n_bins) doubles, 16-byte aligned
 * outputs
 *   logl_dev [P] (status != 0 -> -inf), mus_adj_dev [P, S] adjusted mus (may be NULL),
 *   flags_dev [P] BI_BB_* bits; the sum over bins of A*w of each point is left in
 *   scratch_dev[bi_binned_sum_t_offset(P, n_bins) + p] (input of bi_binned_pmfs).
 * Kernel: the anchor rows of a 256-bin tile are staged in shared memory by 1-D TMA bulk copies once per (tile, group of
 * <= 32 points that share their hypercube cell) and every point of the group is evaluated on them; with Beeston-Barlow
 * the first pass stores t_b = A_b * w_b per (point, bin) so that the second pass neither repeats the root solve nor
 * re-reads the calibration counts.  Points are processed in passes of <= 1024 (per-bin scratch <= 2 GiB).
 * ld_bins must be even and the tensors 16-byte aligned for the tiled kernel (else a gather kernel is used; BI_BINNED_LEGACY=1
 * in the environment forces it, for A/B runs); both give bit-identical results.
 */
int64_t bi_binned_scratch_doubles(int64_t n_points, int64_t n_bins);
int64_t bi_binned_sum_t_offset(int64_t n_points, int64_t n_bins);
int bi_binned_ll_batch(const double* pmf_anchor_dev, const double* n_model_anchor_dev,
                       const double* n_model_sum_anchor_dev,
                       int64_t ld_bins, int64_t n_bins, int32_t n_sources, int32_t n_corners,
                       int32_t bb_source, const double* observed_dev, const double* lgamma_obs_dev,
                       const int32_t* corner_dev, const double* weight_dev, const double* mus_dev,
                       const int32_t* status_dev, int64_t n_points, double* scratch_dev,
                       double* logl_dev, double* mus_adj_dev, int32_t* flags_dev, void* stream);

/* bi_binned_ll_batch with ONE DATASET PER POINT (binned toys): point p is evaluated on the observed counts
 * observed_dev[p * observed_stride + b] (lgamma_obs_dev likewise); observed_stride = 0 is bi_binned_ll_batch. */
int bi_binned_ll_batch_toys(const double* pmf_anchor_dev, const double* n_model_anchor_dev,
                            const double* n_model_sum_anchor_dev,
                            int64_t ld_bins, int64_t n_bins, int32_t n_sources, int32_t n_corners,
                            int32_t bb_source, const double* observed_dev, const double* lgamma_obs_dev,
                            int64_t observed_stride,
                            const int32_t* corner_dev, const double* weight_dev, const double* mus_dev,
                            const int32_t* status_dev, int64_t n_points, double* scratch_dev,
                            double* logl_dev, double* mus_adj_dev, int32_t* flags_dev, void* stream);

/* Morphed (and BB-adjusted) pmf grid [S, n_bins] for ONE point (full_output=True);
 * corner/weight/mus point at that point's rows, sum_t_dev at its sum over bins of A*w. */
int bi_binned_pmfs(const double* pmf_anchor_dev, const double* n_model_anchor_dev,
                   const double* n_model_sum_anchor_dev,
                   int64_t ld_bins, int64_t n_bins, int32_t n_sources, int32_t n_corners,
                   int32_t bb_source, const double* observed_dev,
                   const int32_t* corner_dev, const double* weight_dev, const double* mus_dev,
                   const double* sum_t_dev, double* pmf_out_dev, int64_t ld_out, void* stream);

/*
 * The exchange steps of the sharded evaluations (the reference has no multi-device path) over
 * NVLink / NVSwitch PEER MEMORY: every rank maps every rank's exchange buffer into its own address space (e.g. torch
 * symmetric memory's buffer_ptrs) and passes the mapped pointers as peer_ptrs_host[world]; world <= 16.
 *
 * bi_peer_exchange: ONE launch = the whole exchange.  (1) this rank's n_src values are stored into every rank's buffer
 * (P2P stores), (2) this rank's flag is released on every rank, (3) the kernel waits (acquire) until every rank's flag has
 * arrived here, (4) epilogue:
 *   mode 0 (gather, point / toy sharding): out_dev[r * n + i] = value i of rank r          (out_dev: world * n doubles)
 *   mode 1 (sum, event sharding):          total_i = ((v_0i + v_1i) + v_2i) + ... in RANK ORDER (bit-identical on every
 *                                          rank, independent of timing); out_dev[i] = -musum_dev[i] + total_i
 *                                          (likelihood.py:690), -inf where status_dev[i] != 0; musum_dev / status_dev may
 *                                          be NULL (then out_dev[i] = total_i)
 * The epoch counter lives in the buffer itself, so the launch can be captured in a CUDA graph and replayed; every rank
 * must issue the same sequence of exchanges on a buffer.  The buffer holds bi_peer_exchange_words(world, n) 8-byte
 * words: data [2][world][n] (two slots used alternately), flag [world], local [4] (epoch, two CTA counters, error: set
 * to the epoch if a peer did not arrive within 20 s).  It must be ZEROED, followed by one cross-rank barrier, before
 * the first exchange.  Results a caller reads out of out_dev are stream-ordered after the launch.
 *
 * bi_peer_gather_status: bi_peer_exchange in mode 0 that also copies this rank's n_src point status words
 * (status_dev -> status_out, e.g. pinned host memory) in the same launch, so that a sharded evaluation ends without a copy.
 *
 * bi_peer_broadcast: step (1) alone, peer[r][dst_offset + i] = src_dev[i]; the caller provides the barrier.
 */
int64_t bi_peer_exchange_words(int32_t world, int64_t n);
int bi_peer_exchange(const double* src_dev, int64_t n_src, int64_t n, const uint64_t* peer_ptrs_host,
                     int32_t world, int32_t rank, int32_t mode, const double* musum_dev,
                     const int32_t* status_dev, double* out_dev, void* stream);
int bi_peer_gather_status(const double* src_dev, int64_t n_src, int64_t n, const uint64_t* peer_ptrs_host,
                          int32_t world, int32_t rank, const int32_t* status_dev, int32_t* status_out,
                          double* out_dev, void* stream);
int bi_peer_broadcast(const double* src_dev, int64_t n, const uint64_t* peer_ptrs_host, int32_t world,
                      int64_t dst_offset, void* stream);

/*
 * Micro-benchmarks used by bench.py to measure the roofline denominators that
 * MEASURED_PEAKS.json does not hold: dependent-free FP64 FMA throughput
 * and a plain streaming read.  Each returns elapsed milliseconds (CUDA events on `stream`) in *ms_host.
 */
int bi_bench_fp64_fma(int64_t fma_per_thread, int32_t n_blocks, double* sink_dev, float* ms_host,
                      double* flops_host, void* stream);
int bi_bench_fp64_mma(int64_t mma_per_warp, int32_t n_blocks, double* sink_dev, float* ms_host,
                      double* flops_host, void* stream);
int bi_bench_stream_read(const double* src_dev, int64_t n_doubles, double* sink_dev, float* ms_host,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BLUEICE_B200_H */
