// Instance-mask writeback (sm_100a): per-point track labels -> full-resolution uint8 instance-id maps.
//
// Replaces, per frame, the reference's label scatter -> one-hot [K,h,w] -> x4 bilinear up-sampling -> crop of the zero
// padding -> bilinear resize to the image size -> `> 0.5` -> condensation loop (stemseg/inference/output_utils/
// davis.py:76-112; the same chain in youtube_vis.py:117-161 and kitti_mots.py:101-166).  The K separate interpolations
// collapse into one pass: bilinear weights are non-negative and sum to 1, so at most ONE instance can exceed 0.5 at a
// pixel; the kernel evaluates the two-stage interpolation only for the (<= 16) distinct instances whose points touch
// the pixel, in the reference's fp32 evaluation order (no FMA contraction), and writes that instance's rank.
// Pure bandwidth work: reads a uint8 rank map at 1/4 resolution (L2-resident), writes one byte per output pixel.
#include "common.cuh"

namespace stemseg {
namespace {

// rank_map[idx[i]] = lut[label[i]] (0 for labels outside the table / outliers); the map is zeroed beforehand
__global__ void __launch_bounds__(256) rank_scatter_kernel(const int* __restrict__ idx, const long long* __restrict__ labels,
                                                           long long n, const unsigned char* __restrict__ lut, int nlut,
                                                           unsigned char* __restrict__ rank_map) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= n) return;
    const long long l = labels[i];
    rank_map[idx[i]] = (l >= 0 && l < nlut) ? lut[l] : static_cast<unsigned char>(0);
}

struct Lin {
    int i0, i1;
    float w0, w1;
};
// ATen's align_corners=False source index / lambda computation (area_pixel_compute_source_index + guard)
__device__ __forceinline__ Lin lin_tap(int dst, float scale, int in_size) {
    float s = __fsub_rn(__fmul_rn(scale, __fadd_rn(static_cast<float>(dst), 0.5f)), 0.5f);
    if (s < 0.f) s = 0.f;
    Lin t;
    t.i0 = static_cast<int>(s);
    if (t.i0 > in_size - 1) t.i0 = in_size - 1;
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    float lam = __fsub_rn(s, static_cast<float>(t.i0));
    lam = fminf(fmaxf(lam, 0.f), 1.f);
    t.w1 = lam;
    t.w0 = __fsub_rn(1.f, lam);
    return t;
}

__global__ void __launch_bounds__(256) mask_writeback_kernel(const unsigned char* __restrict__ rank_map, int frames,
                                                             int h, int w, int up, int crop_h, int crop_w, int out_h,
                                                             int out_w, float scale_h, float scale_w,
                                                             unsigned char* __restrict__ out) {
    const long long total = 1ll * frames * out_h * out_w;
    const float inv_up = 1.0f / static_cast<float>(up);
    for (long long p = blockIdx.x * 256ll + threadIdx.x; p < total; p += 256ll * gridDim.x) {
        const int x = static_cast<int>(p % out_w);
        const int y = static_cast<int>((p / out_w) % out_h);
        const int f = static_cast<int>(p / (1ll * out_w * out_h));
        const unsigned char* m = rank_map + static_cast<size_t>(f) * h * w;
        // stage 2 taps: positions in the cropped, x`up` up-sampled map
        const Lin ty = lin_tap(y, scale_h, crop_h), tx = lin_tap(x, scale_w, crop_w);
        const int Y[2] = {ty.i0, ty.i1}, X[2] = {tx.i0, tx.i1};
        // stage 1 taps of each of them: positions in the 1/`up`-resolution map
        Lin sy[2], sx[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            sy[a] = up == 1 ? Lin{Y[a], Y[a], 1.f, 0.f} : lin_tap(Y[a], inv_up, h);
            sx[a] = up == 1 ? Lin{X[a], X[a], 1.f, 0.f} : lin_tap(X[a], inv_up, w);
        }
        unsigned char lab[2][2][2][2];       // [stage-2 y][stage-2 x][stage-1 y][stage-1 x]
        bool uniform = true;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int ys[2] = {sy[a].i0, sy[a].i1}, xs[2] = {sx[b].i0, sx[b].i1};
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int d = 0; d < 2; ++d) {
                        lab[a][b][c][d] = m[ys[c] * w + xs[d]];
                        uniform &= lab[a][b][c][d] == lab[0][0][0][0];
                    }
            }
        unsigned char result = 0;
        if (uniform) {
            result = lab[0][0][0][0];        // every tap belongs to the same instance (or to none): weight ~1 > 0.5
        } else {
            // evaluate the reference's two-stage expression for every distinct non-zero candidate
#pragma unroll 1
            for (int cand_i = 0; cand_i < 16 && result == 0; ++cand_i) {
                const unsigned char cand = (&lab[0][0][0][0])[cand_i];
                if (cand == 0) continue;
                bool seen = false;
                for (int k = 0; k < cand_i; ++k) seen |= (&lab[0][0][0][0])[k] == cand;
                if (seen) continue;
                float v[2][2];
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const float m00 = lab[a][b][0][0] == cand ? 1.f : 0.f, m01 = lab[a][b][0][1] == cand ? 1.f : 0.f;
                        const float m10 = lab[a][b][1][0] == cand ? 1.f : 0.f, m11 = lab[a][b][1][1] == cand ? 1.f : 0.f;
                        const float r0 = __fadd_rn(__fmul_rn(sx[b].w0, m00), __fmul_rn(sx[b].w1, m01));
                        const float r1 = __fadd_rn(__fmul_rn(sx[b].w0, m10), __fmul_rn(sx[b].w1, m11));
                        v[a][b] = __fadd_rn(__fmul_rn(sy[a].w0, r0), __fmul_rn(sy[a].w1, r1));
                    }
                const float r0 = __fadd_rn(__fmul_rn(tx.w0, v[0][0]), __fmul_rn(tx.w1, v[0][1]));
                const float r1 = __fadd_rn(__fmul_rn(tx.w0, v[1][0]), __fmul_rn(tx.w1, v[1][1]));
                const float val = __fadd_rn(__fmul_rn(ty.w0, r0), __fmul_rn(ty.w1, r1));
                if (val > 0.5f) result = cand;
            }
        }
        out[p] = result;
    }
}

}  // namespace
}  // namespace stemseg

using namespace stemseg;

extern "C" int32_t stemseg_rank_map_scatter(const int32_t* indices, const int64_t* labels, int64_t n, const uint8_t* lut,
                                            int32_t nlut, uint8_t* rank_map, int64_t map_elems, void* stream_) {
    SS_REQUIRE(rank_map && map_elems >= 1 && lut && nlut >= 1, "rank_map_scatter: bad arguments");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SS_CUDA_OK(cudaMemsetAsync(rank_map, 0, static_cast<size_t>(map_elems), stream));
    if (n == 0) return STEMSEG_OK;
    SS_REQUIRE(indices && labels && n > 0, "rank_map_scatter: null input");
    rank_scatter_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
        indices, reinterpret_cast<const long long*>(labels), n, lut, nlut, rank_map);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_mask_writeback(const uint8_t* rank_map, int32_t frames, int32_t h, int32_t w,
                                          int32_t upscale, int32_t crop_h, int32_t crop_w, int32_t out_h, int32_t out_w,
                                          uint8_t* out, void* stream_) {
    SS_REQUIRE(rank_map && out, "mask_writeback: null pointer");
    SS_REQUIRE(frames >= 1 && h >= 1 && w >= 1 && upscale >= 1 && out_h >= 1 && out_w >= 1, "mask_writeback: bad shape");
    SS_REQUIRE(crop_h >= 1 && crop_w >= 1 && crop_h <= h * upscale && crop_w <= w * upscale,
               "mask_writeback: crop %dx%d exceeds the up-sampled map %dx%d", crop_h, crop_w, h * upscale, w * upscale);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    // ATen: scale = (float)input_size / output_size when only `size` is given (align_corners=False)
    const float scale_h = static_cast<float>(crop_h) / static_cast<float>(out_h);
    const float scale_w = static_cast<float>(crop_w) / static_cast<float>(out_w);
    const long long total = 1ll * frames * out_h * out_w;
    long long blocks = (total + 255) / 256;
    const long long cap = 16ll * device_sm_count();
    if (blocks > cap) blocks = cap;
    mask_writeback_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(rank_map, frames, h, w, upscale, crop_h,
                                                                              crop_w, out_h, out_w, scale_h, scale_w, out);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}
