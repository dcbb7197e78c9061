// Device side of the sub-clip stitch (sm_100a): label pair histograms and LUT relabelling.
//
// Replaces the per-pair mask reductions of OnlineChainer.associate_clusters (stemseg/inference/online_chainer.py:
// 315-328: K1*K2 times `(l1_active & l2_active).sum()` / `(l1_active | l2_active).sum()` over the overlap points),
// the per-association `torch.where` relabelling (online_chainer.py:219-224) and the per-frame `unique` / count loops of
// TrackContainer.get_track_mask_idxes (online_chainer.py:94-117).  Integer work, one pass over the label vectors;
// the <= 21 x 21 count table goes to the host for the Hungarian solve (scipy, microseconds).
#include "common.cuh"

namespace stemseg {
namespace {

constexpr int kHistThreads = 256;
constexpr int kSmemBins = 8192;      // 32 KB of int32 counters

// bin(v) = 0 for negative labels (outliers), v - base + 1 otherwise; values outside [0, nbins) are a caller error
__device__ __forceinline__ int label_bin(long long v, long long base, int nbins) {
    if (v < 0) return 0;
    const long long b = v - base + 1;
    return (b < 1 || b >= nbins) ? -1 : static_cast<int>(b);
}

__global__ void __launch_bounds__(kHistThreads) pair_histogram_kernel(const long long* __restrict__ a,
                                                                      const long long* __restrict__ b, long long n,
                                                                      long long a_base, long long b_base, int na, int nb,
                                                                      int* __restrict__ table, int* __restrict__ bad) {
    extern __shared__ int s_bins[];
    const int bins = na * nb;
    const bool use_smem = bins <= kSmemBins;
    if (use_smem) {
        for (int i = threadIdx.x; i < bins; i += kHistThreads) s_bins[i] = 0;
        __syncthreads();
    }
    for (long long i = blockIdx.x * 1ll * kHistThreads + threadIdx.x; i < n; i += 1ll * gridDim.x * kHistThreads) {
        const int ia = label_bin(a[i], a_base, na), ib = label_bin(b[i], b_base, nb);
        if (ia < 0 || ib < 0) {
            atomicAdd(bad, 1);
            continue;
        }
        if (use_smem) atomicAdd(&s_bins[ia * nb + ib], 1);
        else atomicAdd(&table[ia * nb + ib], 1);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < bins; i += kHistThreads)
            if (s_bins[i] != 0) atomicAdd(&table[i], s_bins[i]);
    }
}

__global__ void __launch_bounds__(256) relabel_lut_kernel(long long* __restrict__ labels, long long n, long long base,
                                                          const long long* __restrict__ lut, int nlut) {
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
        const long long v = labels[i];
        if (v < 0) continue;
        const long long k = v - base;
        if (k >= 0 && k < nlut) labels[i] = lut[k];
    }
}

}  // namespace
}  // namespace stemseg

using namespace stemseg;

extern "C" int32_t stemseg_label_pair_histogram(const int64_t* a, const int64_t* b, int64_t n, int64_t a_base,
                                                int64_t b_base, int32_t na, int32_t nb, int32_t* table,
                                                int32_t* out_of_range, void* stream_) {
    SS_REQUIRE(table && out_of_range, "label_pair_histogram: null output");
    SS_REQUIRE(na >= 1 && nb >= 1 && 1ll * na * nb <= (1 << 24), "label_pair_histogram: bad table size %d x %d", na, nb);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SS_CUDA_OK(cudaMemsetAsync(table, 0, sizeof(int32_t) * na * nb, stream));
    SS_CUDA_OK(cudaMemsetAsync(out_of_range, 0, sizeof(int32_t), stream));
    if (n == 0) return STEMSEG_OK;
    SS_REQUIRE(a && b && n > 0, "label_pair_histogram: null input");
    long long blocks = (n + kHistThreads * 8 - 1) / (kHistThreads * 8);
    const long long cap = 2ll * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const size_t smem = (1ll * na * nb <= kSmemBins) ? sizeof(int) * na * nb : 0;
    pair_histogram_kernel<<<static_cast<unsigned>(blocks), kHistThreads, smem, stream>>>(
        reinterpret_cast<const long long*>(a), reinterpret_cast<const long long*>(b), n, a_base, b_base, na, nb, table,
        out_of_range);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_relabel_lut(int64_t* labels, int64_t n, int64_t base, const int64_t* lut, int32_t nlut,
                                       void* stream_) {
    if (n == 0) return STEMSEG_OK;
    SS_REQUIRE(labels && lut && n > 0 && nlut >= 1, "relabel_lut: bad arguments");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    long long blocks = (n + 255) / 256;
    const long long cap = 8ll * device_sm_count();
    if (blocks > cap) blocks = cap;
    relabel_lut_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(reinterpret_cast<long long*>(labels), n, base,
                                                                          reinterpret_cast<const long long*>(lut), nlut);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}
