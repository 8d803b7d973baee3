// blueice_b200 -- C-ABI entry of the DMMA K2 kernels: bi_unbinned_mma.cuh (K = C*S <= 128 contraction terms, per-warp tile
// rings) and bi_unbinned_wide.cu (longer contractions: K-chunk loop over CTA-shared tiles).
#include <stdlib.h>

#include "bi_unbinned_mma.cuh"

#define BI_MMA_NARROW_MAX_TERMS 128

// bi_unbinned_wide.cu
int bi_launch_mma_wide(const double* A, int64_t ld, int64_t N, int K, int S, const int32_t* group_points,
                       int32_t* groups, int32_t* header, int64_t n_super, const int32_t* row, const double* coef,
                       const double* wterm, const int32_t* term_source, const double* mus, double outlier,
                       double* partial, int64_t n_points, double* coef_chunks, cudaStream_t st);
int bi_mma_wide_unit_points(void);
int64_t bi_mma_wide_scratch_doubles(int K, int64_t n_points);

#define BI_MMA_WIDE_MIN_TERMS_DEFAULT 33
#define BI_MMA_WIDE_MIN_POINTS_DEFAULT 1

// Which of the two kernels (they give identical bits) -- K2 alone, measured on B200 (profiles/r2_wide_kernel.md):
//  * 4096-point scans over 1e5 events: the per-warp rings win at K = 16 (22.2 against 17.0 TFLOP/s), the K-chunk kernel
//    from K = 24 on (19.8 / 22.8 / 23.5 / 23.2 against 18.9 / 21.2 / 12.8 / 14.2 at K = 24 / 32 / 40 / 48; from K = 80 on the
//    per-warp rings hold one m-tile per warp and fall to 5-6 TFLOP/s against 25-27);
//  * P = 1 / 11 / 64 / 256 / 1024 points over datasets larger than L2: at K = 32 the per-warp rings are faster up to 1024
//    points (0.065 / 0.19 / 0.33 / 0.79 / 2.80 ms against 0.097 / 0.40 / 0.78 / 1.10 / 2.98), at K = 64 and 128 the K-chunk
//    kernel is faster at every batch size (K = 128: 0.11 / 0.38 / 0.76 / 1.49 / 2.90 ms against 0.21 / 0.79 / 1.75 / 4.25 / 14.1).
// So the K-chunk kernel takes every contraction of more than 32 terms (the k-step count at which the per-warp rings leave
// their register-resident B fragments).  Both thresholds -- terms, and points of the batch -- can be set in the
// environment (tests, A/B runs); BI_MMA_WIDE_MIN_TERMS set in the environment applies to every batch size.
static bool bi_mma_use_wide(int n_terms, int64_t n_points) {
    static int min_terms = -1, min_points = -1;
    static bool terms_forced = false;
    if (min_terms < 0) {
        const char* env = getenv("BI_MMA_WIDE_MIN_TERMS");
        const int v = env ? atoi(env) : 0;
        terms_forced = v >= 1;
        min_terms = terms_forced ? v : BI_MMA_WIDE_MIN_TERMS_DEFAULT;
        const char* envp = getenv("BI_MMA_WIDE_MIN_POINTS");
        const int vp = envp ? atoi(envp) : 0;
        min_points = vp >= 1 ? vp : BI_MMA_WIDE_MIN_POINTS_DEFAULT;
    }
    if (n_terms > BI_MMA_NARROW_MAX_TERMS) return true;
    if (n_terms < min_terms) return false;
    return terms_forced || n_points >= min_points;
}

static int bi_mma_k4(int K) {
    const int k4 = (K + 3) / 4;
    return k4 <= 8 ? k4 : (k4 <= 12 ? 12 : (k4 <= 16 ? 16 : (k4 <= 24 ? 24 : 32)));
}

// Tensor map of the anchor tensor [n_1]..[n_D][S][ld] (innermost first: events, sources, last shape parameter ...)
// with a box of one event tile x all sources x the 2 anchors per shape parameter of a hypercube cell.
// Returns the rank (0: not applicable / encoder unavailable -> per-row bulk copies).
static int bi_make_tensor_map(CUtensorMap* tmap, const double* rows, int64_t ld, int32_t S, int32_t n_dims,
                              const int32_t* n_anchors_host, int row_stride) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    static bool looked_up = false;
    if (!looked_up) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &status) == cudaSuccess &&
            status == cudaDriverEntryPointSuccess)
            encode = (encode_fn)fn;
        looked_up = true;
    }
    if (!encode || n_dims < 0 || n_dims > 3 || S > 256) return 0;
    const int rank = n_dims + 2;
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
    dims[0] = (cuuint64_t)ld;  box[0] = (cuuint32_t)row_stride;
    dims[1] = (cuuint64_t)S;   box[1] = (cuuint32_t)S;
    strides[0] = (cuuint64_t)ld * sizeof(double);
    cuuint64_t stride = strides[0] * (cuuint64_t)S;
    for (int d = 0; d < n_dims; ++d) {                   // tensor dim 2 + d = shape parameter n_dims - 1 - d
        dims[2 + d] = (cuuint64_t)n_anchors_host[n_dims - 1 - d];
        box[2 + d] = 2;
        strides[1 + d] = stride;
        stride *= dims[2 + d];
    }
    for (int i = 0; i < rank - 1; ++i)
        if (strides[i] % 16 || strides[i] >= ((cuuint64_t)1 << 40)) return 0;
    const CUresult res = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, (void*)rows, dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return res == CUDA_SUCCESS ? rank : 0;
}

extern "C" int bi_unbinned_partials_mma(const double* rows_dev, int64_t ld_events, int64_t n_events,
                                        int32_t n_terms, int32_t n_sources,
                                        const int32_t* group_points_dev, int32_t* groups_dev, int32_t* header_dev,
                                        const int32_t* row_dev, const double* coef_dev, const double* wterm_dev,
                                        const int32_t* term_source_dev, const double* mus_dev,
                                        double outlier_likelihood, double* partial_dev,
                                        int32_t grid_dims, const int32_t* n_anchors_host, const int32_t* cell_dev,
                                        int64_t n_points, double* coef_chunks_dev, void* stream) {
    BI_REQUIRE(n_events >= 0, "n_events < 0");
    const int64_t n_super = bi_num_superblocks(n_events);
    if (n_super == 0) return BI_OK;
    BI_REQUIRE(rows_dev && group_points_dev && groups_dev && header_dev && row_dev && coef_dev && wterm_dev &&
                   term_source_dev && mus_dev && partial_dev, "bi_unbinned_partials_mma: NULL pointer");
    BI_REQUIRE(ld_events >= n_events && (ld_events % 2) == 0, "ld_events=%lld must be even and >= n_events=%lld",
               (long long)ld_events, (long long)n_events);
    BI_REQUIRE(((uintptr_t)rows_dev & 15) == 0, "rows_dev must be 16-byte aligned");
    BI_REQUIRE(((uintptr_t)groups_dev & 15) == 0, "groups_dev must be 16-byte aligned");
    BI_REQUIRE(n_sources >= 1 && n_sources <= BI_MAX_SOURCES, "n_sources=%d outside [1,%d]", n_sources, BI_MAX_SOURCES);
    BI_REQUIRE(n_terms >= 1 && n_terms <= BI_MMA_MAX_TERMS, "bi_unbinned_partials_mma supports 1..%d contraction terms (got %d)",
               BI_MMA_MAX_TERMS, n_terms);
    BI_REQUIRE(grid_dims < 0 || (grid_dims <= BI_MAX_DIMS && (grid_dims == 0 || (n_anchors_host && cell_dev))),
               "bi_unbinned_partials_mma: grid_dims needs n_anchors_host and cell_dev");
    cudaStream_t st = (cudaStream_t)stream;
    if (bi_mma_use_wide(n_terms, n_points))
        return bi_launch_mma_wide(rows_dev, ld_events, n_events, n_terms, n_sources, group_points_dev, groups_dev, header_dev,
                                  n_super, row_dev, coef_dev, wterm_dev, term_source_dev, mus_dev, outlier_likelihood,
                                  partial_dev, n_points, coef_chunks_dev, st);
    const int k4 = bi_mma_k4(n_terms);
    // full-grid layout ([G][S][ld], term k = corner * S + source): one tiled TMA instruction per event tile
    alignas(64) CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int tmap_rank = 0;
    if (grid_dims >= 0 && k4 <= 8 && n_terms == (n_sources << grid_dims) && ld_events < ((int64_t)1 << 31))
        tmap_rank = bi_make_tensor_map(&tmap, rows_dev, ld_events, n_sources, grid_dims, n_anchors_host, bi_mma_row_stride(k4));
#define BI_MMA_CASE(KK)                                                                                          \
    case KK:                                                                                                     \
        return bi_launch_mma<KK>(rows_dev, ld_events, n_events, n_terms, n_sources, group_points_dev, groups_dev, \
                                 header_dev, n_super, row_dev, coef_dev, wterm_dev, term_source_dev, mus_dev,    \
                                 outlier_likelihood, partial_dev, tmap, tmap_rank, cell_dev,                     \
                                 grid_dims > 0 ? grid_dims : 0, st);
    switch (k4) {
        BI_MMA_CASE(1) BI_MMA_CASE(2) BI_MMA_CASE(3) BI_MMA_CASE(4)
        BI_MMA_CASE(5) BI_MMA_CASE(6) BI_MMA_CASE(7) BI_MMA_CASE(8)
        BI_MMA_CASE(12) BI_MMA_CASE(16) BI_MMA_CASE(24) BI_MMA_CASE(32)
    }
#undef BI_MMA_CASE
    bi_set_error("unsupported contraction length %d", n_terms);
    return BI_ERR_UNSUPPORTED;
}

extern "C" int64_t bi_mma_coef_chunks_doubles(int32_t n_terms, int64_t n_points) {
    if (n_terms < 1 || n_points < 0 || !bi_mma_use_wide(n_terms, n_points)) return 0;
    return bi_mma_wide_scratch_doubles(n_terms, n_points);
}

extern "C" int32_t bi_mma_unit_points(int32_t n_terms, int64_t n_points) {
    if (bi_mma_use_wide(n_terms, n_points)) return bi_mma_wide_unit_points();
    const int k4 = bi_mma_k4(n_terms);
    return 8 * (k4 <= 2 ? BI_MMA_MT_SMALL : (k4 <= 4 ? 4 : (k4 <= 16 ? 2 : 1)));
}

// ---------------------------------------------------------------------------------------------
// the whole unbinned hot path in one call: K1 -> schedule -> K2 -> finalize
// ---------------------------------------------------------------------------------------------
static inline int64_t bi_align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

struct BiUnbinnedWorkspace {
    int64_t cell, frac, corner, weight, mus, partial, group_points, groups, header, row, coef, wterm, term_source,
        coef_chunks, total;
};
#define BI_WORKSPACE_REGIONS 15

static BiUnbinnedWorkspace bi_unbinned_layout(int32_t D, int32_t S, int64_t K, int64_t P, int64_t n_events) {
    const int64_t C = (int64_t)1 << D, Dd = D > 0 ? D : 1, n_super = bi_num_superblocks(n_events);
    BiUnbinnedWorkspace w;
    int64_t o = 0;
    w.cell = o;         o += bi_align256(P * Dd * 4);
    w.frac = o;         o += bi_align256(P * Dd * 8);
    w.corner = o;       o += bi_align256(P * C * 4);
    w.weight = o;       o += bi_align256(P * C * 8);
    w.mus = o;          o += bi_align256(P * S * 8);
    w.partial = o;      o += bi_align256(P * (n_super > 0 ? n_super : 1) * 8);
    w.group_points = o; o += bi_align256(P * 4);
    w.groups = o;       o += bi_align256((P + 1) * 16);
    w.header = o;       o += 256;
    w.row = o;          o += bi_align256(P * K * 4);
    w.coef = o;         o += bi_align256(P * K * 8);
    w.wterm = o;        o += bi_align256(P * K * 8);
    w.term_source = o;  o += bi_align256(K * 4);
    w.coef_chunks = o;  o += bi_align256(bi_mma_coef_chunks_doubles((int32_t)K, P) * 8);   // K-chunk kernel only
    w.total = o;
    return w;
}

extern "C" int64_t bi_unbinned_workspace_bytes(int32_t n_dims, int32_t n_sources, int32_t n_terms, int64_t n_points,
                                               int64_t n_events) {
    if (n_dims < 0 || n_dims > BI_MAX_DIMS || n_sources < 1 || n_terms < 1 || n_points < 0 || n_events < 0) return -1;
    return bi_unbinned_layout(n_dims, n_sources, n_terms, n_points, n_events).total;
}

// offsets (bytes) of the workspace regions, in the order of BiUnbinnedWorkspace (15 entries incl. the total)
extern "C" int bi_unbinned_workspace_layout(int32_t n_dims, int32_t n_sources, int32_t n_terms, int64_t n_points,
                                            int64_t n_events, int64_t* offsets_host) {
    BI_REQUIRE(offsets_host && n_dims >= 0 && n_dims <= BI_MAX_DIMS && n_sources >= 1 && n_terms >= 1 && n_points >= 0 &&
                   n_events >= 0, "bi_unbinned_workspace_layout: bad arguments");
    const BiUnbinnedWorkspace w = bi_unbinned_layout(n_dims, n_sources, n_terms, n_points, n_events);
    const int64_t v[BI_WORKSPACE_REGIONS] = {w.cell, w.frac, w.corner, w.weight, w.mus, w.partial, w.group_points,
                                             w.groups, w.header, w.row, w.coef, w.wterm, w.term_source, w.coef_chunks,
                                             w.total};
    for (int i = 0; i < BI_WORKSPACE_REGIONS; ++i) offsets_host[i] = v[i];
    return BI_OK;
}

// shared tail of the fused calls: schedule -> K2 -> finalize on a workspace whose K1 regions are filled
static int bi_unbinned_after_setup(int32_t n_dims, const int32_t* n_anchors_host, bool full_grid, int32_t n_sources, int32_t n_terms,
                                   int64_t n_points, const double* rows_dev, int64_t ld_events, int64_t n_events,
                                   double outlier_likelihood, int32_t target_units, char* base,
                                   const BiUnbinnedWorkspace& w, double* logl_dev, double* logsum_dev,
                                   double* musum_dev, int32_t* status_dev, void* stream) {
    const int64_t n_super = bi_num_superblocks(n_events);
    double* partial = (double*)(base + w.partial);
    if (n_super > 0) {
        int rc = bi_unbinned_plan(n_dims, n_anchors_host, n_points, (int32_t*)(base + w.cell), status_dev,
                                  bi_mma_unit_points(n_terms, n_points), n_events, target_units & 0x3fffffff, target_units >> 30,
                                  (int32_t*)(base + w.group_points), (int32_t*)(base + w.groups),
                                  (int32_t*)(base + w.header), stream);
        if (rc != BI_OK) return rc;
        rc = bi_unbinned_partials_mma(rows_dev, ld_events, n_events, n_terms, n_sources,
                                      (int32_t*)(base + w.group_points), (int32_t*)(base + w.groups),
                                      (int32_t*)(base + w.header), (int32_t*)(base + w.row), (double*)(base + w.coef),
                                      (double*)(base + w.wterm), (int32_t*)(base + w.term_source),
                                      (double*)(base + w.mus), outlier_likelihood, partial,
                                      full_grid ? n_dims : -1, n_anchors_host, (int32_t*)(base + w.cell), n_points,
                                      (double*)(base + w.coef_chunks), stream);
        if (rc != BI_OK) return rc;
    }
    return bi_unbinned_finalize(partial, n_super, musum_dev, status_dev, n_points, logl_dev, logsum_dev, stream);
}

extern "C" int bi_unbinned_ll_batch(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                                    int32_t n_sources, int64_t n_points,
                                    const double* zs_dev, const double* rate_mult_dev, const double* scale_dev,
                                    const double* eff_dev, const double* mus_anchor_dev,
                                    const uint8_t* allow_negative_host,
                                    const double* ps_anchor_dev, int64_t ld_events, int64_t n_events,
                                    double outlier_likelihood, int32_t target_units,
                                    void* workspace_dev, int64_t workspace_bytes,
                                    double* logl_dev, double* logsum_dev, double* musum_dev, int32_t* status_dev,
                                    void* stream) {
    BI_REQUIRE(n_points >= 0 && n_events >= 0, "negative size");
    if (n_points == 0) return BI_OK;
    BI_REQUIRE(n_dims >= 0 && n_dims <= BI_MAX_DIMS, "n_dims=%d outside [0,%d]", n_dims, BI_MAX_DIMS);
    BI_REQUIRE(n_sources >= 1 && n_sources <= BI_MAX_SOURCES, "n_sources=%d outside [1,%d]", n_sources, BI_MAX_SOURCES);
    const int32_t K = (1 << n_dims) * n_sources;
    const BiUnbinnedWorkspace w = bi_unbinned_layout(n_dims, n_sources, K, n_points, n_events);
    BI_REQUIRE(workspace_dev && workspace_bytes >= w.total, "workspace too small: %lld < %lld bytes",
               (long long)workspace_bytes, (long long)w.total);
    BI_REQUIRE(((uintptr_t)workspace_dev & 255) == 0, "workspace_dev must be 256-byte aligned");
    BI_REQUIRE(logl_dev && musum_dev && status_dev, "bi_unbinned_ll_batch: NULL output pointer");
    // tiny batches: the whole evaluation in ONE launch (bit-identical; BI_SMALL=0 keeps the four launches, for A/B runs)
    const char* small_env = getenv("BI_SMALL");
    if ((small_env == nullptr || small_env[0] != '0') && bi_unbinned_small_ok(n_dims, n_sources, n_points, n_events))
        return bi_unbinned_ll_small(n_dims, n_anchors_host, axes_host, n_sources, n_points, zs_dev, rate_mult_dev, scale_dev,
                                    eff_dev, mus_anchor_dev, allow_negative_host, ps_anchor_dev, ld_events, n_events,
                                    outlier_likelihood, logl_dev, logsum_dev, musum_dev, status_dev, stream);
    char* base = (char*)workspace_dev;
    int rc = bi_point_setup(n_dims, n_anchors_host, axes_host, n_sources, n_points, zs_dev, rate_mult_dev, scale_dev,
                            eff_dev, mus_anchor_dev, allow_negative_host, (int32_t*)(base + w.cell),
                            (double*)(base + w.frac), (int32_t*)(base + w.corner), (double*)(base + w.weight),
                            (double*)(base + w.mus), musum_dev, status_dev, (int32_t*)(base + w.row),
                            (double*)(base + w.coef), (double*)(base + w.wterm), (int32_t*)(base + w.term_source), stream);
    if (rc != BI_OK) return rc;
    return bi_unbinned_after_setup(n_dims, n_anchors_host, true, n_sources, K, n_points, ps_anchor_dev, ld_events, n_events,
                                   outlier_likelihood, target_units, base, w, logl_dev, logsum_dev, musum_dev,
                                   status_dev, stream);
}

extern "C" int bi_unbinned_ll_batch_sourcewise(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                                               int32_t n_sources, const uint32_t* dim_mask_host,
                                               const int32_t* row_base_host, int64_t n_points,
                                               const double* zs_dev, const double* rate_mult_dev,
                                               const double* scale_dev, const double* eff_dev,
                                               const double* mus_rows_dev, const uint8_t* allow_negative_host,
                                               const double* rows_dev, int64_t ld_events, int64_t n_events,
                                               double outlier_likelihood, int32_t target_units,
                                               void* workspace_dev, int64_t workspace_bytes,
                                               double* logl_dev, double* logsum_dev, double* musum_dev,
                                               int32_t* status_dev, void* stream) {
    BI_REQUIRE(n_points >= 0 && n_events >= 0, "negative size");
    if (n_points == 0) return BI_OK;
    BI_REQUIRE(n_dims >= 0 && n_dims <= BI_MAX_DIMS, "n_dims=%d outside [0,%d]", n_dims, BI_MAX_DIMS);
    const int32_t K = bi_sourcewise_terms(n_sources, dim_mask_host);
    BI_REQUIRE(K >= 1, "bi_unbinned_ll_batch_sourcewise: bad source descriptors");
    const BiUnbinnedWorkspace w = bi_unbinned_layout(n_dims, n_sources, K, n_points, n_events);
    BI_REQUIRE(workspace_dev && workspace_bytes >= w.total, "workspace too small: %lld < %lld bytes",
               (long long)workspace_bytes, (long long)w.total);
    BI_REQUIRE(((uintptr_t)workspace_dev & 255) == 0, "workspace_dev must be 256-byte aligned");
    BI_REQUIRE(logl_dev && musum_dev && status_dev, "bi_unbinned_ll_batch_sourcewise: NULL output pointer");
    char* base = (char*)workspace_dev;
    int rc = bi_point_setup_sourcewise(n_dims, n_anchors_host, axes_host, n_sources, dim_mask_host, row_base_host,
                                       n_points, zs_dev, rate_mult_dev, scale_dev, eff_dev, mus_rows_dev,
                                       allow_negative_host, (int32_t*)(base + w.cell), (double*)(base + w.frac),
                                       (double*)(base + w.mus), musum_dev, status_dev, (int32_t*)(base + w.row),
                                       (double*)(base + w.coef), (double*)(base + w.wterm),
                                       (int32_t*)(base + w.term_source), stream);
    if (rc != BI_OK) return rc;
    return bi_unbinned_after_setup(n_dims, n_anchors_host, false, n_sources, K, n_points, rows_dev, ld_events, n_events,
                                   outlier_likelihood, target_units, base, w, logl_dev, logsum_dev, musum_dev,
                                   status_dev, stream);
}

// ps[S, N] of ONE point from its contraction terms, reference operation order per source:
// value = 0; value = value + V * w over the source's corners; a source with a single term of weight
// exactly 1 from a 0-dimensional sub-grid is copied (likelihood.py:543-544).
__global__ void __launch_bounds__(256)
k_unbinned_ps_terms(const double* __restrict__ rows, int64_t ld, int64_t N, int K, int S,
                    const int32_t* __restrict__ row, const double* __restrict__ wterm,
                    const int32_t* __restrict__ term_source, const uint8_t* __restrict__ copy_source,
                    double* __restrict__ out, int64_t ld_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int s = 0; s < S; ++s) {
        double acc = 0.0;
        for (int k = 0; k < K; ++k) {
            if (term_source[k] != s) continue;
            const double v = rows[(int64_t)row[k] * ld + i];
            acc = copy_source[s] ? v : __dadd_rn(acc, __dmul_rn(v, wterm[k]));
        }
        out[(int64_t)s * ld_out + i] = acc;
    }
}

extern "C" int bi_unbinned_ps_terms(const double* rows_dev, int64_t ld_events, int64_t n_events, int32_t n_terms,
                                    int32_t n_sources, const int32_t* row_dev, const double* wterm_dev,
                                    const int32_t* term_source_dev, const uint8_t* copy_source_dev,
                                    double* ps_out_dev, int64_t ld_out, void* stream) {
    BI_REQUIRE(n_events >= 0 && n_terms >= 1 && n_sources >= 1, "bi_unbinned_ps_terms: bad sizes");
    if (n_events == 0) return BI_OK;
    BI_REQUIRE(rows_dev && row_dev && wterm_dev && term_source_dev && copy_source_dev && ps_out_dev && ld_out >= n_events,
               "bi_unbinned_ps_terms: bad arguments");
    const int64_t blocks = (n_events + 255) / 256;
    k_unbinned_ps_terms<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rows_dev, ld_events, n_events, n_terms,
                                                                            n_sources, row_dev, wterm_dev,
                                                                            term_source_dev, copy_source_dev,
                                                                            ps_out_dev, ld_out);
    BI_LAUNCH_CHECK();
    return BI_OK;
}
