// blueice_b200 -- the exchange steps of the sharded evaluations over NVLink / NVSwitch peer memory.
//
// The reference has no multi-device path; these kernels serve blueice_b200.distributed:
//   point / toy sharding   all-gather of the ranks' logL rows                        -> bi_peer_exchange, mode GATHER
//   event sharding         rank-ordered sum of the shards' per-point log sums, then
//                          logL = -sum(mu) + total (likelihood.py:690)              -> bi_peer_exchange, mode SUM
// ONE launch does the whole exchange: P2P stores of this rank's rows into every rank's buffer, a release of this rank's
// flag on every rank, an acquire-wait for every rank's flag here, and the epilogue (copy-out or rank-ordered sum).  It
// keeps its epoch in device memory, so the launch can be captured in a CUDA graph and replayed.
#include <string.h>

#include "bi_common.cuh"

#define BI_MAX_PEERS 16
struct BiPeers { double* p[BI_MAX_PEERS]; };

// ---------------------------------------------------------------------------------------------
// stores only (the caller provides the barrier): kept for callers that confirm delivery once per many steps
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_peer_broadcast(const double* __restrict__ src, int64_t n, const __grid_constant__ BiPeers peers, int world,
                 int64_t dst_offset) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = src[i];
        for (int r = 0; r < world; ++r) peers.p[r][dst_offset + i] = v;
    }
    __threadfence_system();
}

extern "C" int bi_peer_broadcast(const double* src_dev, int64_t n, const uint64_t* peer_ptrs_host, int32_t world,
                                 int64_t dst_offset, void* stream) {
    BI_REQUIRE(world >= 1 && world <= BI_MAX_PEERS, "world=%d outside [1,%d]", world, BI_MAX_PEERS);
    BI_REQUIRE(n >= 0 && dst_offset >= 0, "negative size");
    if (n == 0) return BI_OK;
    BI_REQUIRE(src_dev && peer_ptrs_host, "bi_peer_broadcast: NULL pointer");
    BiPeers peers;
    memset(&peers, 0, sizeof(peers));
    for (int r = 0; r < world; ++r) {
        BI_REQUIRE(peer_ptrs_host[r] != 0, "bi_peer_broadcast: peer %d has no buffer", r);
        peers.p[r] = reinterpret_cast<double*>(peer_ptrs_host[r]);
    }
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    k_peer_broadcast<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src_dev, n, peers, world, dst_offset);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

// ---------------------------------------------------------------------------------------------
// bi_peer_exchange
//
// Exchange buffer of every rank (same layout everywhere, 8-byte words, ZEROED before the first exchange and followed by
// one cross-rank barrier):
//   data  [2][world][n]   double   slot (epoch parity, source rank): two slots used alternately -- a rank can only start
//                                  exchange e + 2 (which overwrites the slot of e) after every rank has signalled e + 1,
//                                  and a rank signals e + 1 only after its own stream-ordered reads of e are done
//   flag  [world]         uint64   flag[r] = last epoch whose rows rank r has delivered HERE (written by rank r, release)
//   local [4]             uint64   epoch, arrive counter, done counter, error (this rank only)
// ---------------------------------------------------------------------------------------------
#define BI_EXCHANGE_TIMEOUT_NS 20000000000ull      /* a peer that never arrives: give up, set the error word */

__device__ __forceinline__ void bi_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long bi_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long bi_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int MODE>      // 0: gather -> out[world, n];  1: rank-ordered sum -> out[n] = -musum + total (status != 0: -inf)
__global__ void __launch_bounds__(256)
k_peer_exchange(const double* __restrict__ src, int64_t n_src, int64_t n, const __grid_constant__ BiPeers peers,
                int world, int rank, const double* __restrict__ musum, const int32_t* __restrict__ status,
                double* __restrict__ out, int32_t* __restrict__ status_out) {
    double* mine = peers.p[rank];
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(mine + 2 * (int64_t)world * n);
    unsigned long long* local = flags + world;
    __shared__ unsigned long long s_epoch;
    __shared__ int s_last;
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned long long*>(local) + 1;
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    const int64_t slot = (int64_t)(epoch & 1) * world * n;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;

    // 1. this rank's rows into every rank's slot (rows beyond n_src repeat nothing: they are left as they are)
    for (int64_t i = tid; i < n_src; i += nthr) {
        const double v = src[i];
        for (int r = 0; r < world; ++r) peers.p[(rank + r) % world][slot + (int64_t)rank * n + i] = v;
    }
    // (gather mode: this rank's point status travels to the caller's buffer -- pinned host memory -- in the same launch,
    // while the peers' rows are still on their way)
    if (MODE == 0 && status_out != nullptr)
        for (int64_t i = tid; i < n_src; i += nthr) status_out[i] = status[i];
    __threadfence_system();
    __syncthreads();
    // 2. the last CTA to arrive publishes: flag[rank] = epoch on every rank
    if (threadIdx.x == 0) {
        const unsigned long long arrived = atomicAdd(local + 1, 1ull);
        s_last = (arrived == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < world) {
            unsigned long long* peer_flags =
                reinterpret_cast<unsigned long long*>(peers.p[threadIdx.x] + 2 * (int64_t)world * n);
            bi_st_release_sys(peer_flags + rank, epoch);
        }
    }
    // 3. every CTA waits until all ranks have delivered this epoch here
    if (threadIdx.x < world) {
        const unsigned long long t0 = bi_globaltimer();
        while (bi_ld_acquire_sys(flags + threadIdx.x) < epoch) {
            if (bi_globaltimer() - t0 > BI_EXCHANGE_TIMEOUT_NS) { atomicExch(local + 3, epoch); break; }
            __nanosleep(20);
        }
    }
    __syncthreads();
    // 4. epilogue on the delivered rows (L1 bypassed: they were written by other devices)
    if (MODE == 0) {
        const int64_t total = (int64_t)world * n;
        const double* rows = mine + slot;
        if ((((uintptr_t)rows | (uintptr_t)out) & 15) == 0) {         // 16-byte stores towards (pinned host) memory
            const int64_t pairs = total >> 1;
            for (int64_t i = tid; i < pairs; i += nthr)
                reinterpret_cast<double2*>(out)[i] = __ldcg(reinterpret_cast<const double2*>(rows) + i);
            if ((total & 1) && tid == 0) out[total - 1] = __ldcg(rows + total - 1);
        } else {
            for (int64_t i = tid; i < total; i += nthr) out[i] = __ldcg(rows + i);
        }
    } else {
        for (int64_t i = tid; i < n; i += nthr) {
            double acc = __ldcg(mine + slot + i);                        // ((r0 + r1) + r2) + ...: fixed rank order, so
            for (int r = 1; r < world; ++r)                               // every rank holds the same bits
                acc = __dadd_rn(acc, __ldcg(mine + slot + (int64_t)r * n + i));
            const bool bad = status != nullptr && status[i] != 0;
            out[i] = bad ? -INFINITY : (musum != nullptr ? __dadd_rn(-musum[i], acc) : acc);
        }
    }
    // 5. the last CTA to finish closes the epoch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long done = atomicAdd(local + 2, 1ull);
        if (done == gridDim.x - 1) {
            local[1] = 0;
            local[2] = 0;
            __threadfence();
            *reinterpret_cast<volatile unsigned long long*>(local) = epoch;
        }
    }
}

extern "C" int64_t bi_peer_exchange_words(int32_t world, int64_t n) {
    if (world < 1 || n < 0) return 0;
    return 2 * (int64_t)world * (n > 0 ? n : 1) + world + 4;
}

static int bi_peer_exchange_impl(const double* src_dev, int64_t n_src, int64_t n, const uint64_t* peer_ptrs_host,
                                 int32_t world, int32_t rank, int32_t mode, const double* musum_dev,
                                 const int32_t* status_dev, double* out_dev, int32_t* status_out, void* stream) {
    BI_REQUIRE(world >= 1 && world <= BI_MAX_PEERS, "world=%d outside [1,%d]", world, BI_MAX_PEERS);
    BI_REQUIRE(rank >= 0 && rank < world, "rank=%d outside [0,%d)", rank, world);
    BI_REQUIRE(n >= 1 && n_src >= 0 && n_src <= n, "bi_peer_exchange: need 0 <= n_src <= n, n >= 1");
    BI_REQUIRE(mode == 0 || mode == 1, "bi_peer_exchange: mode must be 0 (gather) or 1 (sum)");
    BI_REQUIRE(peer_ptrs_host && out_dev && (src_dev || n_src == 0), "bi_peer_exchange: NULL pointer");
    BiPeers peers;
    memset(&peers, 0, sizeof(peers));
    for (int r = 0; r < world; ++r) {
        BI_REQUIRE(peer_ptrs_host[r] != 0, "bi_peer_exchange: peer %d has no buffer", r);
        peers.p[r] = reinterpret_cast<double*>(peer_ptrs_host[r]);
    }
    // enough CTAs to keep a few thousand stores in flight; all of them must be able to be resident while they wait
    int64_t blocks = ((int64_t)world * n + 1023) / 1024;
    if (blocks < 1) blocks = 1;
    if (blocks > 128) blocks = 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0)
        k_peer_exchange<0><<<(unsigned)blocks, 256, 0, st>>>(src_dev, n_src, n, peers, world, rank, musum_dev, status_dev, out_dev,
                                                             status_out);
    else
        k_peer_exchange<1><<<(unsigned)blocks, 256, 0, st>>>(src_dev, n_src, n, peers, world, rank, musum_dev, status_dev, out_dev,
                                                             nullptr);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

extern "C" int bi_peer_exchange(const double* src_dev, int64_t n_src, int64_t n, const uint64_t* peer_ptrs_host,
                                int32_t world, int32_t rank, int32_t mode, const double* musum_dev,
                                const int32_t* status_dev, double* out_dev, void* stream) {
    return bi_peer_exchange_impl(src_dev, n_src, n, peer_ptrs_host, world, rank, mode, musum_dev, status_dev, out_dev, nullptr,
                                 stream);
}

// gather (mode 0) that also delivers this rank's n_src point status words to status_out (e.g. pinned host memory)
extern "C" int bi_peer_gather_status(const double* src_dev, int64_t n_src, int64_t n, const uint64_t* peer_ptrs_host,
                                     int32_t world, int32_t rank, const int32_t* status_dev, int32_t* status_out,
                                     double* out_dev, void* stream) {
    BI_REQUIRE(status_dev && status_out, "bi_peer_gather_status: NULL status pointer");
    return bi_peer_exchange_impl(src_dev, n_src, n, peer_ptrs_host, world, rank, 0, nullptr, status_dev, out_dev, status_out,
                                 stream);
}
