// blueice_b200 -- C-ABI entry of the DMMA K2 kernel (bi_unbinned_mma.cuh), K = C*S <= 32 contraction terms.
#include "bi_unbinned_mma.cuh"

extern "C" int bi_unbinned_partials_mma(const double* ps_anchor_dev, int64_t ld_events, int64_t n_events,
                                        int32_t n_sources, int32_t n_corners,
                                        const int32_t* group_points_dev, const int32_t* groups_dev, int32_t* header_dev,
                                        const int32_t* corner_dev, const double* weight_dev, const double* mus_dev,
                                        double outlier_likelihood, double* partial_dev, void* stream) {
    BI_REQUIRE(n_events >= 0, "n_events < 0");
    const int64_t n_super = bi_num_superblocks(n_events);
    if (n_super == 0) return BI_OK;
    BI_REQUIRE(ps_anchor_dev && group_points_dev && groups_dev && header_dev && corner_dev && weight_dev && mus_dev &&
                   partial_dev, "bi_unbinned_partials_mma: NULL pointer");
    BI_REQUIRE(ld_events >= n_events && (ld_events % 2) == 0, "ld_events=%lld must be even and >= n_events=%lld",
               (long long)ld_events, (long long)n_events);
    BI_REQUIRE(((uintptr_t)ps_anchor_dev & 15) == 0, "ps_anchor_dev must be 16-byte aligned");
    BI_REQUIRE(((uintptr_t)groups_dev & 7) == 0, "groups_dev must be 8-byte aligned");
    BI_REQUIRE(n_sources >= 1 && n_corners >= 1 && (n_corners & (n_corners - 1)) == 0, "bad n_sources / n_corners");
    const int K = n_sources * n_corners;
    BI_REQUIRE(K <= BI_MMA_MAX_TERMS, "bi_unbinned_partials_mma supports n_corners * n_sources <= %d (got %d)",
               BI_MMA_MAX_TERMS, K);
    cudaStream_t st = (cudaStream_t)stream;
#define BI_MMA_CASE(KK)                                                                                          \
    case KK:                                                                                                     \
        return bi_launch_mma<KK>(ps_anchor_dev, ld_events, n_events, n_sources, n_corners, group_points_dev,     \
                                 groups_dev, header_dev, n_super, corner_dev, weight_dev, mus_dev,               \
                                 outlier_likelihood, partial_dev, st);
    switch ((K + 3) / 4) {
        BI_MMA_CASE(1) BI_MMA_CASE(2) BI_MMA_CASE(3) BI_MMA_CASE(4)
        BI_MMA_CASE(5) BI_MMA_CASE(6) BI_MMA_CASE(7) BI_MMA_CASE(8)
    }
#undef BI_MMA_CASE
    bi_set_error("unsupported contraction length %d", K);
    return BI_ERR_UNSUPPORTED;
}

extern "C" int32_t bi_mma_unit_points(int32_t n_sources, int32_t n_corners) {
    const int k4 = (n_sources * n_corners + 3) / 4;
    return 8 * (k4 <= 2 ? BI_MT_SMALL : (k4 <= 4 ? 4 : 2));
}

// ---------------------------------------------------------------------------------------------
// the whole unbinned hot path in one call: K1 -> schedule -> K2 -> finalize
// ---------------------------------------------------------------------------------------------
static inline int64_t bi_align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

struct BiUnbinnedWorkspace {
    int64_t cell, frac, corner, weight, mus, partial, group_points, groups, header, total;
};

static BiUnbinnedWorkspace bi_unbinned_layout(int32_t D, int32_t S, int64_t P, int64_t n_events) {
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
    w.groups = o;       o += bi_align256((P + 1) * 8);
    w.header = o;       o += 256;
    w.total = o;
    return w;
}

extern "C" int64_t bi_unbinned_workspace_bytes(int32_t n_dims, int32_t n_sources, int64_t n_points, int64_t n_events) {
    if (n_dims < 0 || n_dims > BI_MAX_DIMS || n_sources < 1 || n_points < 0 || n_events < 0) return -1;
    return bi_unbinned_layout(n_dims, n_sources, n_points, n_events).total;
}

// offsets (bytes) of the workspace regions, in the order of BiUnbinnedWorkspace (10 entries incl. the total)
extern "C" int bi_unbinned_workspace_layout(int32_t n_dims, int32_t n_sources, int64_t n_points, int64_t n_events,
                                            int64_t* offsets_host) {
    BI_REQUIRE(offsets_host && n_dims >= 0 && n_dims <= BI_MAX_DIMS && n_sources >= 1 && n_points >= 0 && n_events >= 0,
               "bi_unbinned_workspace_layout: bad arguments");
    const BiUnbinnedWorkspace w = bi_unbinned_layout(n_dims, n_sources, n_points, n_events);
    const int64_t v[10] = {w.cell, w.frac, w.corner, w.weight, w.mus, w.partial, w.group_points, w.groups, w.header, w.total};
    for (int i = 0; i < 10; ++i) offsets_host[i] = v[i];
    return BI_OK;
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
    const BiUnbinnedWorkspace w = bi_unbinned_layout(n_dims, n_sources, n_points, n_events);
    BI_REQUIRE(workspace_dev && workspace_bytes >= w.total, "workspace too small: %lld < %lld bytes",
               (long long)workspace_bytes, (long long)w.total);
    BI_REQUIRE(((uintptr_t)workspace_dev & 255) == 0, "workspace_dev must be 256-byte aligned");
    BI_REQUIRE(logl_dev && musum_dev && status_dev, "bi_unbinned_ll_batch: NULL output pointer");
    char* base = (char*)workspace_dev;
    int32_t* cell = (int32_t*)(base + w.cell);
    double* frac = (double*)(base + w.frac);
    int32_t* corner = (int32_t*)(base + w.corner);
    double* weight = (double*)(base + w.weight);
    double* mus = (double*)(base + w.mus);
    double* partial = (double*)(base + w.partial);
    int32_t* group_points = (int32_t*)(base + w.group_points);
    int32_t* groups = (int32_t*)(base + w.groups);
    int32_t* header = (int32_t*)(base + w.header);
    const int32_t C = 1 << n_dims;
    int rc = bi_point_setup(n_dims, n_anchors_host, axes_host, n_sources, n_points, zs_dev, rate_mult_dev, scale_dev,
                            eff_dev, mus_anchor_dev, allow_negative_host, cell, frac, corner, weight, mus, musum_dev,
                            status_dev, stream);
    if (rc != BI_OK) return rc;
    const int64_t n_super = bi_num_superblocks(n_events);
    if (n_super > 0) {
        rc = bi_unbinned_plan(n_dims, n_anchors_host, n_points, cell, status_dev, bi_mma_unit_points(n_sources, C),
                              n_events, target_units, group_points, groups, header, stream);
        if (rc != BI_OK) return rc;
        rc = bi_unbinned_partials_mma(ps_anchor_dev, ld_events, n_events, n_sources, C, group_points, groups, header,
                                      corner, weight, mus, outlier_likelihood, partial, stream);
        if (rc != BI_OK) return rc;
    }
    return bi_unbinned_finalize(partial, n_super, musum_dev, status_dev, n_points, logl_dev, logsum_dev, stream);
}
