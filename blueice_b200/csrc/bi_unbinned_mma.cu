// blueice_b200 -- C-ABI entry of the DMMA K2 kernel (bi_unbinned_mma.cuh), K = C*S <= 32 contraction terms.
#include "bi_unbinned_mma.cuh"

extern "C" int bi_unbinned_partials_mma(const double* ps_anchor_dev, int64_t ld_events, int64_t n_events,
                                        int32_t n_sources, int32_t n_corners,
                                        const int32_t* group_points_dev, const int32_t* work_dev, int64_t n_work,
                                        const int32_t* corner_dev, const double* weight_dev,
                                        const double* mus_dev, const int32_t* status_dev,
                                        double outlier_likelihood, double* partial_dev, void* stream) {
    BI_REQUIRE(n_events >= 0 && n_work >= 0, "negative size");
    const int64_t n_super = bi_num_superblocks(n_events);
    if (n_work == 0 || n_super == 0) return BI_OK;
    BI_REQUIRE(ps_anchor_dev && group_points_dev && work_dev && corner_dev && weight_dev && mus_dev && status_dev &&
                   partial_dev, "bi_unbinned_partials_mma: NULL pointer");
    BI_REQUIRE(ld_events >= n_events && (ld_events % 2) == 0, "ld_events=%lld must be even and >= n_events=%lld",
               (long long)ld_events, (long long)n_events);
    BI_REQUIRE(((uintptr_t)ps_anchor_dev & 15) == 0, "ps_anchor_dev must be 16-byte aligned");
    BI_REQUIRE(((uintptr_t)work_dev & 15) == 0, "work_dev must be 16-byte aligned");
    BI_REQUIRE(n_sources >= 1 && n_corners >= 1 && (n_corners & (n_corners - 1)) == 0, "bad n_sources / n_corners");
    const int K = n_sources * n_corners;
    BI_REQUIRE(K <= BI_MMA_MAX_TERMS, "bi_unbinned_partials_mma supports n_corners * n_sources <= %d (got %d)",
               BI_MMA_MAX_TERMS, K);
    cudaStream_t st = (cudaStream_t)stream;
#define BI_MMA_CASE(KK)                                                                                          \
    case KK:                                                                                                     \
        return bi_launch_mma<KK>(ps_anchor_dev, ld_events, n_events, n_sources, n_corners, group_points_dev,     \
                                 work_dev, n_work, n_super, corner_dev, weight_dev, mus_dev, status_dev,         \
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
