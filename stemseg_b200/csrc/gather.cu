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

// ATen align_corners=False source tap for an integer up-sampling factor (scale = 1 / factor)
struct Tap1 {
    int i0, i1;
    float w0, w1;
};
__device__ __forceinline__ Tap1 up_tap(int dst, float inv_factor, int in_size) {
    float s = __fsub_rn(__fmul_rn(inv_factor, __fadd_rn(static_cast<float>(dst), 0.5f)), 0.5f);
    if (s < 0.f) s = 0.f;
    Tap1 t;
    t.i0 = static_cast<int>(s);
    if (t.i0 > in_size - 1) t.i0 = in_size - 1;
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    float lam = __fsub_rn(s, static_cast<float>(t.i0));
    lam = fminf(fmaxf(lam, 0.f), 1.f);
    t.w1 = lam;
    t.w0 = __fsub_rn(1.f, lam);
    return t;
}
// trilinear (1, f, f) interpolation of a [frames][h][w] plane at full-resolution voxel (frame, y, x); transform 1
// applies exp(v)*10 to the corner values first (the reference resizes the already activated bandwidths)
__device__ __forceinline__ float sample_up(const float* __restrict__ plane, int h, int w, int factor, int frame, int y,
                                           int x, int transform) {
    if (factor == 1) {
        float v = __ldg(plane + (static_cast<size_t>(frame) * h + y) * w + x);
        return transform == 1 ? expf(v) * 10.0f : v;
    }
    const float inv = 1.0f / static_cast<float>(factor);
    const Tap1 ty = up_tap(y, inv, h), tx = up_tap(x, inv, w);
    const float* p = plane + static_cast<size_t>(frame) * h * w;
    float v00 = __ldg(p + ty.i0 * w + tx.i0), v01 = __ldg(p + ty.i0 * w + tx.i1);
    float v10 = __ldg(p + ty.i1 * w + tx.i0), v11 = __ldg(p + ty.i1 * w + tx.i1);
    if (transform == 1) {
        v00 = expf(v00) * 10.0f; v01 = expf(v01) * 10.0f; v10 = expf(v10) * 10.0f; v11 = expf(v11) * 10.0f;
    }
    const float r0 = __fadd_rn(__fmul_rn(tx.w0, v00), __fmul_rn(tx.w1, v01));
    const float r1 = __fadd_rn(__fmul_rn(tx.w0, v10), __fmul_rn(tx.w1, v11));
    return __fadd_rn(__fmul_rn(ty.w0, r0), __fmul_rn(ty.w1, r1));
}

// foreground = up-sampled( sum / count[frame] ) > thr : the seediness (or foreground-logit) average over the sub-clips
// covering a frame (stemseg/inference/main.py:93-103, inference_model.py:126-128,207), optionally at `factor` x the
// map resolution (--resize_embeddings, inference_model.py:55-61)
struct MeanUpThresholdSrc {
    const float* sum;          // [frames][h][w]
    const float* count;        // [frames]
    float thr;
    int h, w, factor;
    __device__ __forceinline__ bool operator()(long long i) const {
        const int W = w * factor, H = h * factor;
        const int x = static_cast<int>(i % W);
        const int y = static_cast<int>((i / W) % H);
        const int f = static_cast<int>(i / (1ll * W * H));
        if (factor == 1) return __fdiv_rn(sum[i], count[f]) > thr;
        const float inv = 1.0f / static_cast<float>(factor);
        const Tap1 ty = up_tap(y, inv, h), tx = up_tap(x, inv, w);
        const float* p = sum + static_cast<size_t>(f) * h * w;
        const float c = count[f];
        const float v00 = __fdiv_rn(p[ty.i0 * w + tx.i0], c), v01 = __fdiv_rn(p[ty.i0 * w + tx.i1], c);
        const float v10 = __fdiv_rn(p[ty.i1 * w + tx.i0], c), v11 = __fdiv_rn(p[ty.i1 * w + tx.i1], c);
        const float r0 = __fadd_rn(__fmul_rn(tx.w0, v00), __fmul_rn(tx.w1, v01));
        const float r1 = __fadd_rn(__fmul_rn(tx.w0, v10), __fmul_rn(tx.w1, v11));
        return __fadd_rn(__fmul_rn(ty.w0, r0), __fmul_rn(ty.w1, r1)) > thr;
    }
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

// gather with on-the-fly (1, f, f) trilinear up-sampling: indices address the FULL-resolution grid [T][f*h][f*w]
__global__ void __launch_bounds__(256) fg_gather_up_kernel(const float* __restrict__ src, long long channel_stride,
                                                           int channels, int h, int w, int factor,
                                                           const int* __restrict__ indices, long long n,
                                                           const int* __restrict__ n_dev, int transform,
                                                           float* __restrict__ dst) {
    const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (n_dev != nullptr && *n_dev < n) n = *n_dev;
    if (p >= n) return;
    const long long idx = indices[p];
    const int W = w * factor, H = h * factor;
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const int f = static_cast<int>(idx / (1ll * W * H));
    for (int c = 0; c < channels; ++c)
        dst[p * channels + c] = sample_up(src + c * channel_stride, h, w, factor, f, y, x, transform);
}

// dst[frame_ids[j]][..] += src[j][..]; counts[frame_ids[j]] += 1   (per-frame running sums over sub-clips)
__global__ void __launch_bounds__(256) frame_accumulate_kernel(float* __restrict__ dst, float* __restrict__ counts,
                                                               const float* __restrict__ src,
                                                               const int* __restrict__ frame_ids, int frames,
                                                               long long plane) {
    const long long total = frames * plane;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
        const int j = static_cast<int>(i / plane);
        const long long p = i % plane;
        dst[frame_ids[j] * plane + p] += src[i];
        if (p == 0) counts[frame_ids[j]] += 1.0f;
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

extern "C" int32_t stemseg_fg_compact_mean_threshold(const float* sum, const float* count, float threshold,
                                                     int64_t n_frames, int32_t h, int32_t w, int32_t factor,
                                                     int32_t* indices, int32_t* frame_counts, void* workspace,
                                                     size_t workspace_bytes, void* stream_) {
    SS_REQUIRE(sum && count, "fg_compact_mean_threshold: null input");
    SS_REQUIRE(h >= 1 && w >= 1 && factor >= 1 && factor <= 8, "fg_compact_mean_threshold: bad shape");
    MeanUpThresholdSrc src{sum, count, threshold, h, w, factor};
    return fg_compact_impl(src, n_frames, static_cast<int64_t>(h) * factor * w * factor, indices, frame_counts,
                           workspace, workspace_bytes, stream_);
}

extern "C" int32_t stemseg_fg_gather_upsampled(const float* src, int64_t channel_stride, int32_t channels, int32_t h,
                                               int32_t w, int32_t factor, const int32_t* indices, int64_t n,
                                               const int32_t* n_dev, int32_t transform, float* dst, void* stream_) {
    SS_REQUIRE(channels >= 1 && h >= 1 && w >= 1 && factor >= 1 && factor <= 8, "fg_gather_upsampled: bad shape");
    SS_REQUIRE(transform == 0 || transform == 1, "fg_gather_upsampled: transform must be 0 or 1");
    if (n == 0) return STEMSEG_OK;
    SS_REQUIRE(src && indices && dst && n > 0, "fg_gather_upsampled: bad arguments");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    fg_gather_up_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(src, channel_stride, channels, h, w,
                                                                                     factor, indices, n, n_dev,
                                                                                     transform, dst);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_frame_accumulate(float* dst, float* counts, const float* src, const int32_t* frame_ids,
                                            int32_t frames, int64_t plane, void* stream_) {
    SS_REQUIRE(dst && counts && src && frame_ids && frames >= 1 && plane >= 1, "frame_accumulate: bad arguments");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    long long blocks = (frames * plane + 255) / 256;
    const long long cap = 16ll * device_sm_count();
    if (blocks > cap) blocks = cap;
    frame_accumulate_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(dst, counts, src, frame_ids, frames, plane);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}
