// Sequential Gaussian-bandwidth clustering as ONE persistent cooperative kernel (sm_100a).
//
// Replaces SequentialClustering._process (stemseg/inference/clusterers.py:60-166) -- the reference runs ~15-20
// torch launches and >= 4 host syncs per iteration (boolean-mask gathers, argmax, .tolist()); here every iteration
// is one pass over the points plus one grid barrier, and the secondary assignment recomputes the K distances from
// the centres kept in shared memory instead of materialising the [N,K] matrix (clusterers.py:150).
//
// HBM-bound integer/float scan: per iteration a point costs 4E (embedding) + 4 (seediness) + 4 (state read)
// [+ 4 state write when claimed]; no tensor cores, coalesced/vectorised loads along the point axis, grid sized to
// the SM count (cooperative launch, all CTAs resident).
//
// Bit-exactness contract (oracle/cluster_oracle.py): IEEE fp32 ops without FMA contraction (__f*_rn intrinsics),
// ATen's CPU summation order over the embedding dimension, IEEE sqrt, thresholds tested in the distance domain.
#include "common.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace stemseg {
namespace {

constexpr int kThreads = 512;
constexpr int kMaxE = STEMSEG_MAX_EMBEDDING_DIMS;
constexpr int kMaxI = STEMSEG_MAX_INSTANCES;

struct ClusterArgs {
    const float* emb;
    const float* bw;
    const float* seed;
    const int* n_dev;            // optional device-side point count (<= n); n is then the capacity
    long long n;
    int v;                       // learned-bandwidth columns (E - n_free)
    float free_bw[kMaxE];
    float d1, d2, min_seed;
    int max_inst;
    long long label_start;
    long long* labels;
    int* primary;
    unsigned int* meta;
    unsigned long long* best;    // [max_inst + 1] winner key per iteration (zeroed before launch)
    unsigned int* barrier;       // grid barrier counter (zeroed before launch)
    unsigned int* avail;         // streaming variant: 1 bit per point, set while the point is unassigned
};

// Sum of E terms in the order ATen's CPU sum kernel uses for a contiguous inner reduction (SumKernel.cpp:
// row_sum with 4 partial sums for E < 8; one 8-lane vector accumulator + scalar tail for E >= 8).  Mirrors
// oracle/cluster_oracle.py:aten_inner_sum_f32, which is pinned against torch for E = 1..24.
template <int E>
__device__ __forceinline__ float aten_inner_sum(const float (&t)[E]) {
    if constexpr (E < 8) {
        float p[4] = {0.f, 0.f, 0.f, 0.f};
        constexpr int size_ilp = E / 4;
#pragma unroll
        for (int i = 0; i < size_ilp; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k] = __fadd_rn(p[k], t[4 * i + k]);
#pragma unroll
        for (int i = size_ilp * 4; i < E; ++i) p[0] = __fadd_rn(p[0], t[i]);
#pragma unroll
        for (int k = 1; k < 4; ++k) p[0] = __fadd_rn(p[0], p[k]);
        return p[0];
    } else {
        constexpr int nvec = E / 8;            // 1 or 2 for E <= 16 -> size_ilp == 0: all vectors go into partial 0
        float lane[8];
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            float p0 = 0.f;
#pragma unroll
            for (int vv = 0; vv < nvec; ++vv) p0 = __fadd_rn(p0, t[vv * 8 + l]);
            // (partial0 += partial1..3, all zero, is exact and omitted)
            lane[l] = p0;
        }
        float acc = 0.f;
#pragma unroll
        for (int k = nvec * 8; k < E; ++k) acc = __fadd_rn(acc, t[k]);
#pragma unroll
        for (int l = 0; l < 8; ++l) acc = __fadd_rn(acc, lane[l]);
        return acc;
    }
}

// clusterers.py:57-58: sqrt(sum((x - c)^2 * bw))
template <int E>
__device__ __forceinline__ float mahalanobis(const float (&x)[E], const float* c, const float* b) {
    float t[E];
#pragma unroll
    for (int k = 0; k < E; ++k) {
        const float d = __fsub_rn(x[k], c[k]);
        t[k] = __fmul_rn(__fmul_rn(d, d), b[k]);
    }
    return __fsqrt_rn(aten_inner_sum<E>(t));
}

template <int E, bool VEC>
__device__ __forceinline__ void load_point(const float* __restrict__ emb, long long idx, float (&x)[E]) {
    if constexpr (VEC && E % 4 == 0) {
        const float4* p = reinterpret_cast<const float4*>(emb) + idx * (E / 4);
#pragma unroll
        for (int q = 0; q < E / 4; ++q) {
            const float4 v = __ldg(p + q);
            x[4 * q + 0] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else if constexpr (VEC && E % 2 == 0) {
        const float2* p = reinterpret_cast<const float2*>(emb) + idx * (E / 2);
#pragma unroll
        for (int q = 0; q < E / 2; ++q) {
            const float2 v = __ldg(p + q);
            x[2 * q + 0] = v.x; x[2 * q + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < E; ++k) x[k] = __ldg(emb + idx * E + k);
    }
}

// Orderable key: larger seediness wins, NaN is the maximum (torch.argmax), smaller index wins ties.  0 == "none".
__device__ __forceinline__ unsigned long long make_key(float s, unsigned int idx) {
    unsigned int o;
    if (s != s) {
        o = 0xFFFFFFFFu;
    } else {
        s = s + 0.0f;                                 // -0 -> +0 (argmax compares values)
        const unsigned int b = __float_as_uint(s);
        o = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
        if (o == 0xFFFFFFFFu) o = 0xFFFFFFFEu;        // keep the NaN code unique
    }
    return (static_cast<unsigned long long>(o) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - idx);
}

template <int THREADS = kThreads>
__device__ __forceinline__ void publish_key(unsigned long long key, unsigned long long* slot,
                                            unsigned long long* s_red) {
    key = warp_max_u64(key);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) s_red[warp] = key;
    __syncthreads();
    if (warp == 0) {
        key = lane < (THREADS / 32) ? s_red[lane] : 0ull;
        key = warp_max_u64(key);
        if (lane == 0 && key != 0ull) atomicMax(slot, key);
    }
}

constexpr int kUnroll = 4;      // independent points in flight per thread in the streaming variant

// R > 0: register-resident variant -- every thread keeps its (<= R) points (embedding, key, state) in registers for the
// whole call, so an iteration touches no global memory besides the winner slot: pure grid-barrier latency.  Used
// when N <= gridDim * kThreads * R (quarter-resolution point sets).  R == 0: streaming variant for large N
// (full-resolution clustering): kUnroll independent points per loop trip keep enough loads in flight.
template <int E, bool VEC, int R>
__global__ void __launch_bounds__(kThreads) seq_cluster_kernel(const ClusterArgs a) {
    long long n = a.n;
    if (a.n_dev != nullptr) {
        const long long nd = *a.n_dev;
        if (nd < n) n = nd;
    }
    __shared__ float s_center[kMaxI][E];
    __shared__ float s_bw[kMaxI][E];
    __shared__ unsigned long long s_red[kThreads / 32];

    const long long tid0 = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    constexpr int RR = R > 0 ? R : 1;
    float rx[RR][E];                      // resident points
    unsigned long long rkey[RR];          // their argmax keys (0 = not a valid point)
    int rst[RR];                          // -1 unassigned, else ordinal of the claiming cluster

    // pass 0: all points are unassigned (clusterers.py:96); winner of iteration 0
    {
        unsigned long long key = 0ull;
        if (R > 0) {
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                const long long idx = tid0 + r * stride;
                rkey[r] = 0ull;
                rst[r] = -1;
                if (idx < n) {
                    load_point<E, VEC>(a.emb, idx, rx[r]);
                    rkey[r] = make_key(__ldg(a.seed + idx), static_cast<unsigned int>(idx));
                    key = rkey[r] > key ? rkey[r] : key;
                }
            }
        } else {
            for (long long idx = tid0; idx < n; idx += stride) {
                a.primary[idx] = -1;
                const unsigned long long k = make_key(__ldg(a.seed + idx), static_cast<unsigned int>(idx));
                key = k > key ? k : key;
            }
        }
        publish_key(key, a.best + 0, s_red);
    }
    unsigned int generation = 1;
    grid_barrier(a.barrier, generation);

    int num_clusters = 0;
    int exit_reason = 0;
    for (int i = 0; i < a.max_inst; ++i) {                                  // clusterers.py:106
        const unsigned long long w = __ldcg(a.best + i);
        if (w == 0ull) { exit_reason = 1; break; }                          // clusterers.py:109-110
        const unsigned int widx = 0xFFFFFFFFu - static_cast<unsigned int>(w & 0xFFFFFFFFull);
        const float prob = __ldg(a.seed + widx);
        if (prob < a.min_seed) { exit_reason = 2; break; }                  // clusterers.py:116-117
        if (threadIdx.x < E) {                                              // clusterers.py:119,175
            const int k = threadIdx.x;
            s_center[i][k] = __ldg(a.emb + static_cast<long long>(widx) * E + k);
            s_bw[i][k] = k < a.v ? __ldg(a.bw + static_cast<long long>(widx) * a.v + k) : a.free_bw[k - a.v];
        }
        __syncthreads();
        num_clusters = i + 1;
        if (blockIdx.x == 0 && threadIdx.x == 0) a.meta[4 + i] = widx;

        float c[E], b[E];
#pragma unroll
        for (int k = 0; k < E; ++k) { c[k] = s_center[i][k]; b[k] = s_bw[i][k]; }

        unsigned long long key = 0ull;
        if (R > 0) {
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                if (rkey[r] != 0ull && rst[r] == -1) {                      // clusterers.py:107
                    const float d = mahalanobis<E>(rx[r], c, b);            // clusterers.py:129-130
                    if (d <= a.d1) rst[r] = i;                              // clusterers.py:136-143
                    else key = rkey[r] > key ? rkey[r] : key;
                }
            }
        } else {
            for (long long idx0 = tid0; idx0 < n; idx0 += stride * kUnroll) {
                int st[kUnroll];
                float x[kUnroll][E];
                float sd[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const long long idx = idx0 + u * stride;
                    st[u] = idx < n ? a.primary[idx] : 0;                 // clusterers.py:107
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const long long idx = idx0 + u * stride;
                    if (st[u] == -1) {
                        load_point<E, VEC>(a.emb, idx, x[u]);
                        sd[u] = __ldg(a.seed + idx);
                    }
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const long long idx = idx0 + u * stride;
                    if (st[u] == -1) {
                        const float d = mahalanobis<E>(x[u], c, b);         // clusterers.py:129-130
                        if (d <= a.d1) {                                    // clusterers.py:136-143
                            a.primary[idx] = i;
                        } else {
                            const unsigned long long k = make_key(sd[u], static_cast<unsigned int>(idx));
                            key = k > key ? k : key;
                        }
                    }
                }
            }
        }
        if (i + 1 < a.max_inst) {
            publish_key(key, a.best + i + 1, s_red);
            ++generation;
            grid_barrier(a.barrier, generation);
        }
    }

    // Secondary assignment (clusterers.py:148-159).  `avail` there is the mask taken at the top of the LAST
    // EXECUTED iteration: fresh after a break, stale (points claimed in the final iteration included) when the
    // loop ran out.  Such points were available in every iteration, so all their K distances are real.
    const bool exhausted = exit_reason == 0;
    const bool do_secondary = num_clusters >= 1 && exit_reason != 1;
    auto finish_point = [&](long long idx, int pl, const float (&x)[E]) {
        long long out = pl < 0 ? -1ll : static_cast<long long>(pl) + a.label_start;
        const bool avail = pl < 0 || (exhausted && pl == a.max_inst - 1);
        if (do_secondary && avail) {
            float dmax = 0.f;
            int kmax = 0;
            bool has_nan = false;
            for (int k = 0; k < num_clusters; ++k) {
                const float d = mahalanobis<E>(x, s_center[k], s_bw[k]);
                has_nan |= (d != d);
                if (k == 0 || d > dmax) { dmax = d; kmax = k; }            // first max wins (clusterers.py:153)
            }
            if (!has_nan && dmax <= a.d2) out = static_cast<long long>(kmax) + a.label_start;
        }
        a.labels[idx] = out;
    };
    if (R > 0) {
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const long long idx = tid0 + r * stride;
            if (idx < n) {
                a.primary[idx] = rst[r];
                finish_point(idx, rst[r], rx[r]);
            }
        }
    } else {
        for (long long idx = tid0; idx < n; idx += stride) {
            const int pl = a.primary[idx];
            const bool avail = pl < 0 || (exhausted && pl == a.max_inst - 1);
            float x[E];
            if (do_secondary && avail) load_point<E, VEC>(a.emb, idx, x);
            finish_point(idx, pl, x);
        }
    }

    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            a.meta[0] = static_cast<unsigned int>(num_clusters);
            a.meta[1] = static_cast<unsigned int>(exit_reason);
            a.meta[2] = static_cast<unsigned int>(n);      // points actually clustered
            a.meta[3] = 0u;
        }
        float* centers = reinterpret_cast<float*>(a.meta + 4 + a.max_inst);
        float* bws = centers + static_cast<size_t>(a.max_inst) * E;
        for (int q = threadIdx.x; q < num_clusters * E; q += kThreads) {
            centers[q] = s_center[q / E][q % E];
            bws[q] = s_bw[q / E][q % E];
        }
    }
}


// Streaming variant for point sets that do not fit the register-resident one (full-resolution clustering; HBM-bound).
// State is split in two: `primary` (int32 ordinal of the claiming cluster, written once per point, read by the final
// pass) and a 1-bit-per-point availability mask (N/8 bytes: L2-resident) that the iterations read instead -- an
// assigned point costs 1/8 byte per iteration and a fully assigned 32-point group is skipped without touching its
// embeddings.  Warp w owns the 32*UNROLL-point chunks w, w + W, ... (static mapping: only the owning warp ever reads or
// writes a mask word); per trip it loads UNROLL mask words, then issues all embedding / seediness loads of the
// available points (UNROLL x (E/4 + 1) independent 16-byte loads per lane) before the first distance is evaluated.
template <int E, bool VEC, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) seq_cluster_stream_kernel(const ClusterArgs a) {
    long long n = a.n;
    if (a.n_dev != nullptr) {
        const long long nd = *a.n_dev;
        if (nd < n) n = nd;
    }
    __shared__ float s_center[kMaxI][E];
    __shared__ float s_bw[kMaxI][E];
    __shared__ unsigned long long s_red[THREADS / 32];

    const int lane = threadIdx.x & 31;
    const long long warp_global = (static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x) >> 5;
    const long long total_warps = static_cast<long long>(gridDim.x) * (THREADS / 32);
    const long long n_groups = (n + 31) / 32;
    const long long n_chunks = (n_groups + UNROLL - 1) / UNROLL;

    {   // pass 0: everything unassigned (clusterers.py:96); winner of iteration 0
        unsigned long long key = 0ull;
        for (long long chunk = warp_global; chunk < n_chunks; chunk += total_warps) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const long long g = chunk * UNROLL + u;
                const long long idx = g * 32 + lane;
                const bool valid = idx < n;
                if (valid) {
                    a.primary[idx] = -1;
                    const unsigned long long k = make_key(__ldg(a.seed + idx), static_cast<unsigned int>(idx));
                    key = k > key ? k : key;
                }
                const unsigned int word = __ballot_sync(0xffffffffu, valid);
                if (lane == 0 && g < n_groups) a.avail[g] = word;
            }
        }
        publish_key<THREADS>(key, a.best + 0, s_red);
    }
    unsigned int generation = 1;
    grid_barrier(a.barrier, generation);

    int num_clusters = 0;
    int exit_reason = 0;
    for (int i = 0; i < a.max_inst; ++i) {                                  // clusterers.py:106
        const unsigned long long w = __ldcg(a.best + i);
        if (w == 0ull) { exit_reason = 1; break; }                          // clusterers.py:109-110
        const unsigned int widx = 0xFFFFFFFFu - static_cast<unsigned int>(w & 0xFFFFFFFFull);
        const float prob = __ldg(a.seed + widx);
        if (prob < a.min_seed) { exit_reason = 2; break; }                  // clusterers.py:116-117
        if (threadIdx.x < E) {                                              // clusterers.py:119,175
            const int k = threadIdx.x;
            s_center[i][k] = __ldg(a.emb + static_cast<long long>(widx) * E + k);
            s_bw[i][k] = k < a.v ? __ldg(a.bw + static_cast<long long>(widx) * a.v + k) : a.free_bw[k - a.v];
        }
        __syncthreads();
        num_clusters = i + 1;
        if (blockIdx.x == 0 && threadIdx.x == 0) a.meta[4 + i] = widx;

        float c[E], b[E];
#pragma unroll
        for (int k = 0; k < E; ++k) { c[k] = s_center[i][k]; b[k] = s_bw[i][k]; }

        unsigned long long key = 0ull;
        for (long long chunk = warp_global; chunk < n_chunks; chunk += total_warps) {
            unsigned int m[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const long long g = chunk * UNROLL + u;
                m[u] = g < n_groups ? __ldcg(a.avail + g) : 0u;              // clusterers.py:107
            }
            float x[UNROLL][E];
            float sd[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if ((m[u] >> lane) & 1u) {
                    const long long idx = (chunk * UNROLL + u) * 32 + lane;
                    load_point<E, VEC>(a.emb, idx, x[u]);
                    sd[u] = __ldg(a.seed + idx);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (m[u] == 0u) continue;                                   // warp-uniform
                const long long g = chunk * UNROLL + u;
                const long long idx = g * 32 + lane;
                bool claim = false;
                if ((m[u] >> lane) & 1u) {
                    const float d = mahalanobis<E>(x[u], c, b);             // clusterers.py:129-130
                    if (d <= a.d1) {                                        // clusterers.py:136-143
                        claim = true;
                        a.primary[idx] = i;
                    } else {
                        const unsigned long long k = make_key(sd[u], static_cast<unsigned int>(idx));
                        key = k > key ? k : key;
                    }
                }
                const unsigned int claimed = __ballot_sync(0xffffffffu, claim);
                if (claimed != 0u && lane == 0) a.avail[g] = m[u] & ~claimed;
            }
        }
        if (i + 1 < a.max_inst) {
            publish_key<THREADS>(key, a.best + i + 1, s_red);
            ++generation;
            grid_barrier(a.barrier, generation);
        }
    }

    // Secondary assignment (clusterers.py:148-159); see seq_cluster_kernel for the stale-mask rule.  Same chunk -> warp
    // mapping as the iterations: a thread only reads `primary` entries it wrote itself (the last iteration is not
    // followed by a grid barrier, so claims made there are not yet visible to other threads).
    const bool exhausted = exit_reason == 0;
    const bool do_secondary = num_clusters >= 1 && exit_reason != 1;
    for (long long chunk = warp_global; chunk < n_chunks; chunk += total_warps)
#pragma unroll 1
    for (int u = 0; u < UNROLL; ++u) {
        const long long idx = (chunk * UNROLL + u) * 32 + lane;
        if (idx >= n) continue;
        const int pl = a.primary[idx];
        long long out = pl < 0 ? -1ll : static_cast<long long>(pl) + a.label_start;
        const bool avail = pl < 0 || (exhausted && pl == a.max_inst - 1);
        if (do_secondary && avail) {
            float x[E];
            load_point<E, VEC>(a.emb, idx, x);
            float dmax = 0.f;
            int kmax = 0;
            bool has_nan = false;
            for (int k = 0; k < num_clusters; ++k) {
                const float d = mahalanobis<E>(x, s_center[k], s_bw[k]);
                has_nan |= (d != d);
                if (k == 0 || d > dmax) { dmax = d; kmax = k; }                // first max wins (clusterers.py:153)
            }
            if (!has_nan && dmax <= a.d2) out = static_cast<long long>(kmax) + a.label_start;
        }
        a.labels[idx] = out;
    }

    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            a.meta[0] = static_cast<unsigned int>(num_clusters);
            a.meta[1] = static_cast<unsigned int>(exit_reason);
            a.meta[2] = static_cast<unsigned int>(n);
            a.meta[3] = 0u;
        }
        float* centers = reinterpret_cast<float*>(a.meta + 4 + a.max_inst);
        float* bws = centers + static_cast<size_t>(a.max_inst) * E;
        for (int q = threadIdx.x; q < num_clusters * E; q += THREADS) {
            centers[q] = s_center[q / E][q % E];
            bws[q] = s_bw[q / E][q % E];
        }
    }
}

template <int E, bool VEC, int THREADS, int UNROLL>
int launch_cluster_stream(const ClusterArgs& args, cudaStream_t stream) {
    auto kernel = seq_cluster_stream_kernel<E, VEC, THREADS, UNROLL>;
    int per_sm = 0;
    SS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, 0));
    if (per_sm < 1) {
        set_error("seq_cluster: streaming kernel does not fit on an SM");
        return STEMSEG_ERR_CUDA;
    }
    long long blocks = static_cast<long long>(per_sm) * device_sm_count();
    const long long useful = (args.n + 32ll * UNROLL * (THREADS / 32) - 1) / (32ll * UNROLL * (THREADS / 32));
    if (blocks > useful) blocks = useful;
    void* kargs[] = {const_cast<ClusterArgs*>(&args)};
    SS_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), dim3(static_cast<unsigned>(blocks)),
                                           dim3(THREADS), kargs, 0, stream));
    return STEMSEG_OK;
}

// STEMSEG_CLUSTER_STREAM = "legacy" | "256x4" | "256x8" overrides the per-E choice made in launch_cluster
int stream_variant() {
    static int v = -2;
    if (v != -2) return v;
    const char* e = getenv("STEMSEG_CLUSTER_STREAM");
    v = -1;                                             // -1: choose by embedding size (launch_cluster)
    if (e != nullptr) {
        if (!strcmp(e, "legacy")) v = 0;
        else if (!strcmp(e, "256x4")) v = 1;
        else if (!strcmp(e, "256x8")) v = 2;
    }
    return v;
}

template <int E, bool VEC, int R>
int launch_cluster_r(const ClusterArgs& args, long long blocks, cudaStream_t stream) {
    auto kernel = seq_cluster_kernel<E, VEC, R>;
    int per_sm = 0;
    SS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0));
    const long long resident = static_cast<long long>(per_sm) * device_sm_count();
    if (per_sm < 1 || blocks > resident) {
        set_error("seq_cluster: %lld blocks cannot be co-resident (%lld fit)", blocks, resident);
        return STEMSEG_ERR_CUDA;
    }
    void* kargs[] = {const_cast<ClusterArgs*>(&args)};
    SS_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), dim3(static_cast<unsigned>(blocks)),
                                           dim3(kThreads), kargs, 0, stream));
    return STEMSEG_OK;
}

template <int E, bool VEC, int R>
long long resident_blocks() {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seq_cluster_kernel<E, VEC, R>, kThreads, 0) != cudaSuccess)
        return 0;
    return static_cast<long long>(per_sm) * device_sm_count();
}

template <int E, bool VEC>
int launch_cluster(const ClusterArgs& args, cudaStream_t stream) {
    const long long per_block = kThreads;
    auto blocks_for = [&](int r) { return (args.n + per_block * r - 1) / (per_block * r); };
    // register-resident variants, fewest blocks first (cheapest grid barrier), if all blocks can be co-resident
    if (blocks_for(4) <= resident_blocks<E, VEC, 4>()) return launch_cluster_r<E, VEC, 4>(args, blocks_for(4), stream);
    if (blocks_for(2) <= resident_blocks<E, VEC, 2>()) return launch_cluster_r<E, VEC, 2>(args, blocks_for(2), stream);
    if (blocks_for(1) <= resident_blocks<E, VEC, 1>()) return launch_cluster_r<E, VEC, 1>(args, blocks_for(1), stream);
    // streaming variant: fill the device.  Measured on B200 (profiles/r02_cluster_variants.txt): with >= 32 bytes of
    // embedding per point the bit-mask kernel with 8 groups in flight per warp wins (E=8, N=6.6M: 1.37 ms vs 1.71 ms);
    // with 16-byte points the state-array kernel below is faster (E=4, N=3.3M: 0.44 ms vs 0.50 ms)
    int variant = stream_variant();
    if (variant < 0) variant = E >= 6 ? 2 : 0;
    switch (variant) {
        case 1: return launch_cluster_stream<E, VEC, 256, 4>(args, stream);
        case 2: return launch_cluster_stream<E, VEC, 256, 8>(args, stream);
        default: break;
    }
    long long blocks = resident_blocks<E, VEC, 0>();
    if (blocks < 1) {
        set_error("seq_cluster: kernel does not fit on an SM");
        return STEMSEG_ERR_CUDA;
    }
    if (blocks > blocks_for(1)) blocks = blocks_for(1);
    return launch_cluster_r<E, VEC, 0>(args, blocks, stream);
}

template <int E>
int launch_cluster_e(const ClusterArgs& args, bool vec, cudaStream_t stream) {
    return vec ? launch_cluster<E, true>(args, stream) : launch_cluster<E, false>(args, stream);
}

}  // namespace
}  // namespace stemseg

using namespace stemseg;

static size_t cluster_fixed_ws_bytes(int max_instances) {
    return align_up(sizeof(unsigned long long) * (max_instances + 1) + sizeof(unsigned int), 256);
}

extern "C" size_t stemseg_seq_cluster_meta_words(int32_t e, int32_t max_instances) {
    return 4 + static_cast<size_t>(max_instances) * (1 + 2 * static_cast<size_t>(e));
}

extern "C" int32_t stemseg_seq_cluster_workspace_bytes(const StemsegClusterParams* p, size_t* bytes) {
    SS_REQUIRE(p != nullptr && bytes != nullptr, "seq_cluster_workspace_bytes: null argument");
    SS_REQUIRE(p->max_instances >= 0 && p->max_instances <= kMaxI, "max_instances %d out of range [0,%d]",
               p->max_instances, kMaxI);
    *bytes = cluster_fixed_ws_bytes(p->max_instances) + align_up(sizeof(unsigned int) * ((p->n_points + 31) / 32 + 8), 256);
    return STEMSEG_OK;
}

extern "C" float stemseg_prob_threshold_to_distance(double prob_threshold) {
    // Same definition as oracle/cluster_oracle.py:prob_threshold_to_distance (independent implementation).
    const float p = static_cast<float>(prob_threshold);
    auto prob = [](float d) -> float { return static_cast<float>(std::exp(static_cast<double>(-0.5f * d))); };
    if (p != p) return -1.0f;
    if (!(prob(0.0f) > p)) return -1.0f;
    if (0.0f > p) return INFINITY;
    uint32_t lo = 0u, hi = 0x7F800000u;
    while (hi - lo > 1u) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        float d;
        memcpy(&d, &mid, sizeof(d));
        if (prob(d) > p) lo = mid; else hi = mid;
    }
    float d;
    memcpy(&d, &lo, sizeof(d));
    return d;
}

extern "C" int32_t stemseg_seq_cluster(const float* embeddings, const float* bandwidths, const float* seediness,
                                       const StemsegClusterParams* p, const int32_t* n_points_dev, int64_t* labels,
                                       int32_t* primary, void* meta, void* workspace, size_t workspace_bytes,
                                       void* stream_) {
    SS_REQUIRE(p != nullptr, "seq_cluster: null params");
    SS_REQUIRE(p->n_points >= 1 && p->n_points < 0x7FFFFFFFll, "seq_cluster: n_points %lld out of range",
               static_cast<long long>(p->n_points));
    SS_REQUIRE(p->embedding_dims >= 1 && p->embedding_dims <= kMaxE, "seq_cluster: embedding_dims %d unsupported",
               p->embedding_dims);
    SS_REQUIRE(p->n_free_dims >= 0 && p->n_free_dims <= p->embedding_dims, "seq_cluster: bad n_free_dims");
    SS_REQUIRE(p->max_instances >= 0 && p->max_instances <= kMaxI, "seq_cluster: max_instances out of range");
    SS_REQUIRE(embeddings && seediness && labels && primary && meta && workspace, "seq_cluster: null pointer");
    SS_REQUIRE(bandwidths != nullptr || p->n_free_dims == p->embedding_dims, "seq_cluster: null bandwidths");
    size_t need = 0;
    stemseg_seq_cluster_workspace_bytes(p, &need);
    if (workspace_bytes < need) {
        set_error("seq_cluster: workspace %zu < %zu bytes", workspace_bytes, need);
        return STEMSEG_ERR_WORKSPACE;
    }
    int rc = require_sm100();
    if (rc != STEMSEG_OK) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);

    ClusterArgs a;
    a.emb = embeddings; a.bw = bandwidths; a.seed = seediness;
    a.n_dev = n_points_dev;
    a.n = p->n_points;
    a.v = p->embedding_dims - p->n_free_dims;
    for (int k = 0; k < kMaxE; ++k) a.free_bw[k] = k < p->n_free_dims ? p->free_dim_bandwidths[k] : 0.f;
    a.d1 = p->d_primary; a.d2 = p->d_secondary; a.min_seed = p->min_seediness_prob;
    a.max_inst = p->max_instances;
    a.label_start = p->cluster_label_start;
    a.labels = reinterpret_cast<long long*>(labels);
    a.primary = primary;
    a.meta = static_cast<unsigned int*>(meta);
    a.best = static_cast<unsigned long long*>(workspace);
    a.barrier = reinterpret_cast<unsigned int*>(a.best + p->max_instances + 1);
    a.avail = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(workspace) + cluster_fixed_ws_bytes(p->max_instances));
    SS_CUDA_OK(cudaMemsetAsync(workspace, 0, cluster_fixed_ws_bytes(p->max_instances), stream));

    const bool vec = (reinterpret_cast<uintptr_t>(embeddings) % 16u) == 0;
    switch (p->embedding_dims) {
#define SS_CASE(E) case E: return launch_cluster_e<E>(a, vec, stream);
        SS_CASE(1) SS_CASE(2) SS_CASE(3) SS_CASE(4) SS_CASE(5) SS_CASE(6) SS_CASE(7) SS_CASE(8)
        SS_CASE(9) SS_CASE(10) SS_CASE(11) SS_CASE(12) SS_CASE(13) SS_CASE(14) SS_CASE(15) SS_CASE(16)
#undef SS_CASE
    }
    set_error("seq_cluster: unreachable");
    return STEMSEG_ERR_UNSUPPORTED;
}
