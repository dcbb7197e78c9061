// Foreground compaction + gather (sm_100a).
//
// Replaces masks_to_coord_list (stemseg/inference/online_chainer.py:11-22: one torch.nonzero + host sync per
// frame) and the per-frame permute / advanced-index / cat gather of cluster_subsequence (online_chainer.py:258-281).
// Pure byte/index work, HBM-bound: the mask is read twice (count, write), each map once.
// Output order = linear voxel order t*HW + y*W + x, which is exactly frame-major + torch.nonzero's row-major order.
#include "common.cuh"

namespace stemseg {
namespace {

constexpr int kThreads = 256;
constexpr int kPerThread = 16;                      // mask bytes per thread
constexpr int kChunk = kThreads * kPerThread;       // mask bytes per block

// foreground predicate sources: a uint8 mask (non-zero) or an fp32 map thresholded on the fly
// (fg = seediness > thr, stemseg/inference/main.py:93-103, or fg prob > 0.5, main.py:142-144)
struct MaskSrc {
    const uint8_t* m;
    __device__ __forceinline__ bool operator()(long long i) const { return m[i] != 0; }
};
struct ThresholdSrc {
    const float* v;
    float thr;
    __device__ __forceinline__ bool operator()(long long i) const { return v[i] > thr; }
};

template <class Src>
__device__ __forceinline__ int count_nonzero_run(const Src& m, long long base, long long begin, long long end) {
    int c = 0;
    for (long long i = begin; i < end; ++i) c += m(base + i) ? 1 : 0;
    return c;
}

// blockIdx.x = frame * blocks_per_frame + chunk
template <class Src>
__global__ void __launch_bounds__(kThreads) fg_count_kernel(const Src mask, long long hw, int blocks_per_frame,
                                                            int* __restrict__ block_counts) {
    const int frame = blockIdx.x / blocks_per_frame, chunk = blockIdx.x % blocks_per_frame;
    const long long base = static_cast<long long>(frame) * hw;
    long long begin = static_cast<long long>(chunk) * kChunk + static_cast<long long>(threadIdx.x) * kPerThread;
    long long end = begin + kPerThread;
    if (end > hw) end = hw;
    int c = begin < hw ? count_nonzero_run(mask, base, begin, end) : 0;
    __shared__ int s_warp[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kThreads / 32; ++w) t += s_warp[w];
        block_counts[blockIdx.x] = t;
    }
}

// single block: exclusive scan of block_counts -> block_offsets; per-frame counts + total
__global__ void __launch_bounds__(1024) fg_scan_kernel(const int* __restrict__ block_counts, int n_frames,
                                                       int blocks_per_frame, int* __restrict__ block_offsets,
                                                       int* __restrict__ frame_counts) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int total_blocks = n_frames * blocks_per_frame;
    for (int base = 0; base < total_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < total_blocks ? block_counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += n;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += n;
            }
            s_warp[threadIdx.x] = w;                // inclusive scan of warp totals
        }
        __syncthreads();
        const int warp_prefix = (threadIdx.x >> 5) == 0 ? 0 : s_warp[(threadIdx.x >> 5) - 1];
        const int carry = s_carry;
        if (i < total_blocks) block_offsets[i] = carry + warp_prefix + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_prefix + incl;
        __syncthreads();
    }
    // frame counts from the offsets just written (visible after the barrier above)
    for (int f = threadIdx.x; f < n_frames; f += 1024) {
        const int first = f * blocks_per_frame, last = first + blocks_per_frame - 1;
        frame_counts[f] = block_offsets[last] + block_counts[last] - block_offsets[first];
    }
    if (threadIdx.x == 0) frame_counts[n_frames] = s_carry;
}

template <class Src>
__global__ void __launch_bounds__(kThreads) fg_write_kernel(const Src mask, long long hw, int blocks_per_frame,
                                                            const int* __restrict__ block_offsets,
                                                            int* __restrict__ indices) {
    const int frame = blockIdx.x / blocks_per_frame, chunk = blockIdx.x % blocks_per_frame;
    const long long base = static_cast<long long>(frame) * hw;
    long long begin = static_cast<long long>(chunk) * kChunk + static_cast<long long>(threadIdx.x) * kPerThread;
    long long end = begin + kPerThread;
    if (end > hw) end = hw;
    const int c = begin < hw ? count_nonzero_run(mask, base, begin, end) : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += n;
    }
    __shared__ int s_warp[kThreads / 32];
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    int warp_prefix = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) warp_prefix += s_warp[w];
    int out = block_offsets[blockIdx.x] + warp_prefix + incl - c;
    for (long long i = begin; i < end; ++i)
        if (mask(base + i)) indices[out++] = static_cast<int>(base + i);
}

__global__ void __launch_bounds__(256) fg_gather_kernel(const float* __restrict__ src, long long channel_stride,
                                                        int channels, const int* __restrict__ indices, long long n,
                                                        const int* __restrict__ n_dev, int transform,
                                                        float* __restrict__ dst) {
    const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (n_dev != nullptr && *n_dev < n) n = *n_dev;        // device-side count (no host round trip)
    if (p >= n) return;
    const long long idx = indices[p];
    for (int c = 0; c < channels; ++c) {
        float v = __ldg(src + c * channel_stride + idx);
        if (transform == 1) v = expf(v) * 10.0f;          // bandwidths = exp(variance) * 10, inference_model.py:148
        dst[p * channels + c] = v;
    }
}

}  // namespace
}  // namespace stemseg

using namespace stemseg;

static inline int blocks_per_frame_for(int64_t hw) { return static_cast<int>((hw + kChunk - 1) / kChunk); }

extern "C" size_t stemseg_fg_compact_workspace_bytes(int64_t n_frames, int64_t frame_voxels) {
    if (n_frames <= 0 || frame_voxels <= 0) return 256;
    const size_t blocks = static_cast<size_t>(n_frames) * blocks_per_frame_for(frame_voxels);
    return align_up(2 * blocks * sizeof(int), 256);
}

template <class Src>
static int32_t fg_compact_impl(const Src mask, int64_t n_frames, int64_t frame_voxels, int32_t* indices,
                               int32_t* frame_counts, void* workspace, size_t workspace_bytes, void* stream_) {
    SS_REQUIRE(indices && frame_counts && workspace, "fg_compact: null pointer");
    SS_REQUIRE(n_frames >= 1 && frame_voxels >= 1, "fg_compact: empty mask");
    SS_REQUIRE(n_frames * frame_voxels < 0x7FFFFFFFll, "fg_compact: mask too large for int32 indices");
    const size_t need = stemseg_fg_compact_workspace_bytes(n_frames, frame_voxels);
    if (workspace_bytes < need) {
        set_error("fg_compact: workspace %zu < %zu bytes", workspace_bytes, need);
        return STEMSEG_ERR_WORKSPACE;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int bpf = blocks_per_frame_for(frame_voxels);
    const int blocks = static_cast<int>(n_frames) * bpf;
    int* block_counts = static_cast<int*>(workspace);
    int* block_offsets = block_counts + blocks;
    fg_count_kernel<Src><<<blocks, kThreads, 0, stream>>>(mask, frame_voxels, bpf, block_counts);
    fg_scan_kernel<<<1, 1024, 0, stream>>>(block_counts, static_cast<int>(n_frames), bpf, block_offsets,
                                           frame_counts);
    fg_write_kernel<Src><<<blocks, kThreads, 0, stream>>>(mask, frame_voxels, bpf, block_offsets, indices);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_fg_compact(const uint8_t* mask, int64_t n_frames, int64_t frame_voxels,
                                      int32_t* indices, int32_t* frame_counts, void* workspace,
                                      size_t workspace_bytes, void* stream_) {
    SS_REQUIRE(mask != nullptr, "fg_compact: null mask");
    return fg_compact_impl(MaskSrc{mask}, n_frames, frame_voxels, indices, frame_counts, workspace, workspace_bytes,
                           stream_);
}

extern "C" int32_t stemseg_fg_compact_threshold(const float* values, float threshold, int64_t n_frames,
                                                int64_t frame_voxels, int32_t* indices, int32_t* frame_counts,
                                                void* workspace, size_t workspace_bytes, void* stream_) {
    SS_REQUIRE(values != nullptr, "fg_compact_threshold: null values");
    return fg_compact_impl(ThresholdSrc{values, threshold}, n_frames, frame_voxels, indices, frame_counts, workspace,
                           workspace_bytes, stream_);
}

extern "C" int32_t stemseg_fg_gather(const float* src, int64_t channel_stride, int32_t channels,
                                     const int32_t* indices, int64_t n, const int32_t* n_dev, int32_t transform,
                                     float* dst, void* stream_) {
    SS_REQUIRE(transform == 0 || transform == 1, "fg_gather: transform must be 0 (none) or 1 (exp*10)");
    SS_REQUIRE(channels >= 1, "fg_gather: channels must be >= 1");
    if (n == 0) return STEMSEG_OK;
    SS_REQUIRE(src && indices && dst && n > 0, "fg_gather: bad arguments");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
    fg_gather_kernel<<<blocks, 256, 0, stream>>>(src, channel_stride, channels, indices, n, n_dev, transform, dst);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}
