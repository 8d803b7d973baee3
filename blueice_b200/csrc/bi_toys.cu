// blueice_b200 -- on-device toy Monte Carlo generation: the step BEFORE the hot path
// in a Neyman construction.
//
// Replaces, for models whose sources are histogram templates:
//   Model.simulate                      blueice/model.py:69-91   n_s ~ Poisson(mu_s / fraction_in_range) per source,
//                                                                 events of the sources concatenated in source order
//   HistogramPdfSource.simulate         blueice/source.py:248-264 -> (hist * bin_volumes).get_random(n)
//   multihist Histdd.get_random         bin = searchsorted(cdf / cdf[-1], u) over the flattened histogram, then a
//                                       uniform position inside the bin, lo + u * (hi - lo), per dimension
// The range cut of model.py:90 is the identity here (events are drawn inside the bin edges).
//
// Random numbers: counter-based Philox4x32-10 keyed by the seed with the counter (index, toy id, domain), so toy t is
// the same events whatever the batch, the launch geometry or the number of GPUs the toys are sharded over.
// Parity with the reference is distributional (NumPy's Mersenne-Twister stream is not reproduced); the event stage is
// restated bit for bit in oracle/toys.py.
#include "bi_space.cuh"

struct BiPhilox { uint32_t v[4]; };

__host__ __device__ inline BiPhilox bi_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    BiPhilox out;
    out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}

// 53-bit uniform in [0, 1) from two 32-bit words (the construction of NumPy's random_double)
__host__ __device__ inline double bi_uniform53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// two uniforms of (toy, index, domain)
__device__ __forceinline__ void bi_toy_uniforms(uint64_t seed, int64_t toy, uint32_t index, uint32_t domain, double* u0, double* u1) {
    const BiPhilox r = bi_philox4x32_10(index, (uint32_t)toy, (uint32_t)((uint64_t)toy >> 32), domain, (uint32_t)seed,
                                        (uint32_t)(seed >> 32));
    *u0 = bi_uniform53(r.v[0], r.v[1]);
    *u1 = bi_uniform53(r.v[2], r.v[3]);
}

// ---------------------------------------------------------------------------------------------
// counts[t, s] ~ Poisson(mu[t, s]): multiplication method below 10, Hoermann's PTRS above (the two algorithms of
// NumPy's legacy generator, numpy/random/src/legacy); one thread per (toy, source)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_toy_counts(int n_sources, int64_t n_toys, int64_t toy_id0, const double* __restrict__ mus, int mus_per_toy,
             uint64_t seed, int32_t* __restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_toys * n_sources) return;
    const int64_t t = i / n_sources;
    const int s = (int)(i - t * n_sources);
    const double lam = mus[(mus_per_toy ? t * n_sources : 0) + s];
    const int64_t toy = toy_id0 + t;
    const uint32_t domain = ((uint32_t)s << 8);                      // low byte 0: the count stream of source s
    int32_t k = 0;
    if (!(lam > 0.0)) {
        k = 0;                                                       // mu <= 0 or NaN: no events
    } else if (lam < 10.0) {
        const double enlam = exp(-lam);
        double prod = 1.0;
        for (uint32_t it = 0;; ++it) {
            double u0, u1;
            bi_toy_uniforms(seed, toy, it, domain, &u0, &u1);
            prod *= u0;
            if (!(prod > enlam)) break;
            ++k;
            prod *= u1;
            if (!(prod > enlam)) break;
            ++k;
        }
    } else {
        const double slam = sqrt(lam), loglam = log(lam);
        const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
        const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
        for (uint32_t it = 0;; ++it) {
            double U, V;
            bi_toy_uniforms(seed, toy, it, domain, &U, &V);
            U -= 0.5;
            const double us = 0.5 - fabs(U);
            const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
            if (us >= 0.07 && V <= vr) { k = (int32_t)kf; break; }
            if (kf < 0.0 || (us < 0.013 && V > us)) continue;
            if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + kf * loglam - lgamma(kf + 1.0)) { k = (int32_t)kf; break; }
        }
    }
    counts[i] = k;
}

// ---------------------------------------------------------------------------------------------
// events: one thread per event of the batch; (toy, index in toy) by binary search in the toy offsets
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_toy_events(const __grid_constant__ BiSpace sp, const __grid_constant__ BiPoints pts, int n_points_total,
             int n_sources, const double* __restrict__ cdf, int64_t n_toys, int64_t toy_id0,
             const int32_t* __restrict__ counts, const int64_t* __restrict__ offsets, uint64_t seed,
             double* __restrict__ coords, int64_t ld_coords, int32_t* __restrict__ source_out) {
    __shared__ double s_pts[BI_MAX_EDGE_POINTS];
    bi_stage_points(pts, n_points_total, s_pts);
    const int64_t n_events = offsets[n_toys];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_events) return;
    int64_t lo = 0, hi = n_toys;                                     // largest t with offsets[t] <= e
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= e) lo = mid; else hi = mid;
    }
    const int64_t t = lo;
    const uint32_t j = (uint32_t)(e - offsets[t]);
    int s = 0;                                                       // events of a toy are grouped by source, in order
    uint32_t before = 0;
    while (s < n_sources - 1 && j >= before + (uint32_t)counts[t * n_sources + s]) before += (uint32_t)counts[t * n_sources + s++];
    const int64_t toy = toy_id0 + t;
    double u_bin, u_pos[BI_MAX_SPACE_DIMS];
    bi_toy_uniforms(seed, toy, j, 1, &u_bin, &u_pos[0]);
    if (sp.n_space > 1) bi_toy_uniforms(seed, toy, j, 2, &u_pos[1], &u_pos[2]);
    if (sp.n_space > 3) { double unused; bi_toy_uniforms(seed, toy, j, 3, &u_pos[3], &unused); }
    // bin = min(searchsorted(cdf, u, 'left'), n_cells - 1)
    const double* c = cdf + (int64_t)s * sp.n_cells;
    int flat = bi_lower_bound(c, (int)sp.n_cells, u_bin);
    if (flat > (int)sp.n_cells - 1) flat = (int)sp.n_cells - 1;
    for (int d = 0; d < sp.n_space; ++d) {
        const int idx = (flat / sp.stride[d]) % sp.n_bins[d];
        const double* edges = s_pts + sp.offset[d];
        const double a = edges[idx], b = edges[idx + 1];
        coords[(int64_t)d * ld_coords + e] = __dadd_rn(a, __dmul_rn(u_pos[d], __dsub_rn(b, a)));
    }
    if (source_out) source_out[e] = s;
}

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" int bi_toy_counts(int32_t n_sources, int64_t n_toys, int64_t toy_id0, const double* mus_dev,
                             int32_t mus_per_toy, uint64_t seed, int32_t* counts_dev, void* stream) {
    BI_REQUIRE(n_sources >= 1 && n_sources <= BI_MAX_SOURCES, "n_sources=%d outside [1,%d]", n_sources, BI_MAX_SOURCES);
    BI_REQUIRE(n_toys >= 0 && toy_id0 >= 0, "negative toy count or id");
    if (n_toys == 0) return BI_OK;
    BI_REQUIRE(mus_dev && counts_dev, "bi_toy_counts: NULL device pointer");
    const int64_t n = n_toys * n_sources;
    k_toy_counts<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_sources, n_toys, toy_id0, mus_dev,
                                                                               mus_per_toy, seed, counts_dev);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

extern "C" int bi_toy_events(int32_t n_space, const int32_t* n_bins_host, const double* edges_host, int32_t n_sources,
                             const double* cdf_dev, int64_t n_toys, int64_t toy_id0, const int32_t* counts_dev,
                             const int64_t* offsets_dev, int64_t n_events, uint64_t seed, double* coords_dev,
                             int64_t ld_coords, int32_t* source_dev, void* stream) {
    BiSpace sp;
    int rc = bi_fill_space(&sp, n_space, n_bins_host);
    if (rc != BI_OK) return rc;
    BiPoints pts;
    int total = 0;
    rc = bi_fill_points(&sp, &pts, edges_host, false, &total);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(n_sources >= 1 && n_sources <= BI_MAX_SOURCES, "n_sources=%d outside [1,%d]", n_sources, BI_MAX_SOURCES);
    BI_REQUIRE(n_toys >= 0 && toy_id0 >= 0 && n_events >= 0, "negative size");
    if (n_toys == 0 || n_events == 0) return BI_OK;
    BI_REQUIRE(cdf_dev && counts_dev && offsets_dev && coords_dev && ld_coords >= n_events, "bi_toy_events: bad arguments");
    k_toy_events<<<(unsigned)((n_events + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        sp, pts, total, n_sources, cdf_dev, n_toys, toy_id0, counts_dev, offsets_dev, seed, coords_dev, ld_coords, source_dev);
    BI_LAUNCH_CHECK();
    return BI_OK;
}
