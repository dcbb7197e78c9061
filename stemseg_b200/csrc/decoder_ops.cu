// Decoder-head kernels around the tcgen05 convolution (sm_100a): operand packing, GroupNorm statistics,
// GroupNorm+ReLU(+AvgPool3d) apply, trilinear upsample-add, and the fused output heads.
//
// All of these are HBM-bound elementwise / small-stencil passes over NDHWC tensors: vectorised (float4 / bf16x4)
// accesses along the channel axis, no shared-memory staging needed except for the NCTHW -> NDHWC transpose.
#include "common.cuh"
#include "trilinear.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

namespace stemseg {
namespace {

constexpr int kPlanesFp16 = STEMSEG_PLANES_FP16;     // one fp16 plane stored in the bf16-typed buffer (same 2 bytes)
__host__ __device__ __forceinline__ bool valid_planes(int planes, bool allow_fp16) {
    return planes == 1 || planes == 2 || (allow_fp16 && planes == kPlanesFp16);
}
__device__ __forceinline__ __nv_bfloat16 half_bits_as_bf16(float x) {
    const __half h = __float2half_rn(x);
    return *reinterpret_cast<const __nv_bfloat16*>(&h);
}

// ---- fp32 -> bf16 planes: x ~= hi + lo (lo = bf16(x - hi)), round-to-nearest-even both times ------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

struct alignas(8) bf16x4 {
    __nv_bfloat16 v[4];
};

__device__ __forceinline__ void store_planes4(__nv_bfloat16* hi_plane, size_t plane_elems, int planes, size_t off,
                                              const float (&x)[4]) {
    bf16x4 h, l;
    if (planes == kPlanesFp16) {
#pragma unroll
        for (int k = 0; k < 4; ++k) h.v[k] = half_bits_as_bf16(x[k]);
        *reinterpret_cast<bf16x4*>(hi_plane + off) = h;
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) split_bf16(x[k], h.v[k], l.v[k]);
    *reinterpret_cast<bf16x4*>(hi_plane + off) = h;
    if (planes == 2) *reinterpret_cast<bf16x4*>(hi_plane + plane_elems + off) = l;
}

// ---------------------------------------------------------------------------------------------------------------
// K0: [N][C][T][HW] fp32 (strides sn, sc, st; HW contiguous) -> planes [P][N][T][HW][C] bf16
// Replaces the permute / stack "temporal fusion" copies (model_builder.py:84-99, inference_model.py:112-119).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPackC = 64, kPackS = 32;

__global__ void __launch_bounds__(256) pack_ncthw_kernel(const float* __restrict__ src, long long sn, long long sc,
                                                         long long st, int c, int t, int hw,
                                                         __nv_bfloat16* __restrict__ dst, size_t plane_elems,
                                                         int planes) {
    __shared__ float tile[kPackC][kPackS + 1];
    const int s0 = blockIdx.x * kPackS, c0 = blockIdx.y * kPackC;
    const int nt = blockIdx.z, n = nt / t, tt = nt % t;
    const float* base = src + n * sn + tt * st;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
#pragma unroll
    for (int i = 0; i < kPackC / 8; ++i) {
        const int cc = c0 + ty + 8 * i, s = s0 + tx;
        tile[ty + 8 * i][tx] = (cc < c && s < hw) ? __ldg(base + cc * sc + s) : 0.f;
    }
    __syncthreads();
    // each thread writes 2 channels of one voxel: 32 lanes x 2 = 64 channels = one 128-byte row segment
#pragma unroll
    for (int i = 0; i < kPackS / 8; ++i) {
        const int s = s0 + ty + 8 * i, cc = c0 + 2 * tx;
        if (s < hw && cc < c) {
            const size_t off = ((static_cast<size_t>(nt) * hw + s) * c) + cc;
            __nv_bfloat16 h0, l0, h1, l1;
            if (planes == kPlanesFp16) {
                h0 = half_bits_as_bf16(tile[2 * tx][ty + 8 * i]);
                h1 = half_bits_as_bf16(tile[2 * tx + 1][ty + 8 * i]);
                *reinterpret_cast<__nv_bfloat162*>(dst + off) = __nv_bfloat162(h0, h1);
                continue;
            }
            split_bf16(tile[2 * tx][ty + 8 * i], h0, l0);
            split_bf16(tile[2 * tx + 1][ty + 8 * i], h1, l1);
            *reinterpret_cast<__nv_bfloat162*>(dst + off) = __nv_bfloat162(h0, h1);
            if (planes == 2) *reinterpret_cast<__nv_bfloat162*>(dst + plane_elems + off) = __nv_bfloat162(l0, l1);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing: torch [Cout][Cin_total][taps] fp32 -> planes [P][rows_total][taps][cin_count] bf16 (K-major)
// ---------------------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ src, int cout, int cin_total, int cin_begin,
                                   int cin_count, int taps, __nv_bfloat16* __restrict__ dst, int row_begin,
                                   size_t plane_elems, int planes) {
    const long long total = 1ll * cout * taps * cin_count;
    for (long long i = blockIdx.x * 1ll * blockDim.x + threadIdx.x; i < total; i += 1ll * gridDim.x * blockDim.x) {
        const int ci = static_cast<int>(i % cin_count);
        const int tap = static_cast<int>((i / cin_count) % taps);
        const int co = static_cast<int>(i / (1ll * cin_count * taps));
        const float x = src[(static_cast<size_t>(co) * cin_total + cin_begin + ci) * taps + tap];
        __nv_bfloat16 h, l;
        split_bf16(x, h, l);
        const size_t off = (static_cast<size_t>(row_begin + co) * taps + tap) * cin_count + ci;
        dst[off] = planes == kPlanesFp16 ? half_bits_as_bf16(x) : h;
        if (planes == 2) dst[plane_elems + off] = l;
    }
}

// dgrad weights: conv(dy, W') with W'[ci][taps-1-tap][co] = W[co][ci][tap] (flipped taps, swapped channels)
__global__ void pack_weight_dgrad_kernel(const float* __restrict__ src, int cout, int cin_total, int cin_begin,
                                         int cin_count, int taps, __nv_bfloat16* __restrict__ dst, size_t plane_elems,
                                         int planes) {
    const long long total = 1ll * cin_count * taps * cout;
    for (long long i = blockIdx.x * 1ll * blockDim.x + threadIdx.x; i < total; i += 1ll * gridDim.x * blockDim.x) {
        const int co = static_cast<int>(i % cout);
        const int tapf = static_cast<int>((i / cout) % taps);
        const int ci = static_cast<int>(i / (1ll * cout * taps));
        const float x = src[(static_cast<size_t>(co) * cin_total + cin_begin + ci) * taps + (taps - 1 - tapf)];
        __nv_bfloat16 h, l;
        split_bf16(x, h, l);
        dst[i] = h;
        if (planes == 2) dst[plane_elems + i] = l;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K2a: per-channel partial sums of an NDHWC fp32 tensor; K2b: finalize per (sample, group) mean / rstd.
// Replaces the statistics pass of nn.GroupNorm(32, C) (model_builder.py:34; embedding_decoder.py:22,...).
// Deterministic: fixed chunking, fixed reduction order, final combination in double.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStatsMaxChunk = 256;   // voxels per block (upper bound; small layers use smaller chunks)

// x may be stored as `slices` split-K partial sums [slices][n][spatial][row_stride] that are added (fixed order) on
// read; the statistics kernel writes the sum back into slice 0 so that later passes read one slice only
__device__ __forceinline__ float4 load_sum_slices(const float4* p, size_t slice_stride4, int slices) {
    float4 a = __ldg(p);
    for (int s = 1; s < slices; ++s) {
        const float4 b = __ldg(p + s * slice_stride4);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    return a;
}

__global__ void __launch_bounds__(256) gn_partial_kernel(float* __restrict__ x, long long spatial, int c,
                                                         int row_stride, int slices, size_t slice_stride, int chunk_voxels,
                                                         float* __restrict__ partial /*[n][c][chunks][2]*/,
                                                         int chunks) {
    extern __shared__ float s_acc[];                 // [rows][c][2]
    const int quads = c / 4;
    const int rows = blockDim.x / quads;
    const int q = threadIdx.x % quads, r = threadIdx.x / quads;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const long long v0 = 1ll * chunk * chunk_voxels;
    long long v1 = v0 + chunk_voxels;
    if (v1 > spatial) v1 = spatial;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    float4* base = reinterpret_cast<float4*>(x + (static_cast<size_t>(n) * spatial) * row_stride) + q;
    const int row_quads = row_stride / 4;
    for (long long v = v0 + r; v < v1; v += rows) {
        float4* p = base + v * row_quads;
        float4 a;
        if (slices > 1) {
            a = *p;
            for (int sl = 1; sl < slices; ++sl) {
                const float4 b = __ldcs(p + sl * (slice_stride / 4));
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            *p = a;                                   // reduced value replaces slice 0 (this thread owns the element)
        } else {
            a = *p;
        }
        s[0] += a.x; ss[0] += a.x * a.x;
        s[1] += a.y; ss[1] += a.y * a.y;
        s[2] += a.z; ss[2] += a.z * a.z;
        s[3] += a.w; ss[3] += a.w * a.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s_acc[(r * c + 4 * q + k) * 2 + 0] = s[k];
        s_acc[(r * c + 4 * q + k) * 2 + 1] = ss[k];
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int rr = 0; rr < rows; ++rr) {
            a += s_acc[(rr * c + ch) * 2 + 0];
            b += s_acc[(rr * c + ch) * 2 + 1];
        }
        float* out = partial + ((static_cast<size_t>(n) * c + ch) * chunks + chunk) * 2;   // [n][c][chunks][2]
        out[0] = a;
        out[1] = b;
    }
}

// one block per (group, sample): mean / rstd of the group and the per-channel affine table
//   scale[ch] = rstd * gamma[ch],  shift[ch] = beta[ch] - mean * rstd * gamma[ch]
__global__ void __launch_bounds__(512) gn_finalize_kernel(const float* __restrict__ partial, size_t sample_stride,
                                                          int chunks, int c, int cpg, long long spatial, float eps,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ scale_shift /*[n][c][2]*/,
                                                          float* __restrict__ mean_rstd /*[n][groups][2] or null*/) {
    const int g = blockIdx.x, n = blockIdx.y;
    double s = 0.0, ss = 0.0;
    // the cpg channels of a group are contiguous in the [n][c][chunks][2] layout: one coalesced stream
    const float2* base = reinterpret_cast<const float2*>(partial + n * sample_stride) + static_cast<size_t>(g) * cpg * chunks;
    for (int i = threadIdx.x; i < chunks * cpg; i += blockDim.x) {
        const float2 v = __ldg(base + i);
        s += v.x;
        ss += v.y;
    }
    __shared__ double sh[2][512];
    sh[0][threadIdx.x] = s;
    sh[1][threadIdx.x] = ss;
    __syncthreads();
    for (int o = 256; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x < cpg) {
        const double cnt = static_cast<double>(spatial) * cpg;
        const double mean = sh[0][0] / cnt;
        double var = sh[1][0] / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
        const float m = static_cast<float>(mean);
        const int ch = g * cpg + threadIdx.x;
        float* o = scale_shift + (static_cast<size_t>(n) * c + ch) * 2;
        const float sc = rstd * gamma[ch];
        o[0] = sc;
        o[1] = beta[ch] - m * sc;
        if (mean_rstd != nullptr && threadIdx.x == 0) {
            float* mr = mean_rstd + (static_cast<size_t>(n) * (c / cpg) + g) * 2;
            mr[0] = m;
            mr[1] = rstd;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K2c: y = relu((x - mean) * rstd * gamma + beta) [-> AvgPool3d(3, stride (2,1,1), pad 1, divisor 27)] -> bf16 planes
// Replaces GroupNorm apply + ReLU + AvgPool3d (embedding_decoder.py:22-24; common.py:8-24).
// ---------------------------------------------------------------------------------------------------------------
// POOL: 0 = none, 1 = AvgPool3d(3, stride (2,1,1), padding 1) (zero padding counted: always /27), 2 = MaxPool3d with the
// same window (padding never wins: the pooled values are post-ReLU, >= 0, and every window holds a real voxel)
template <int POOL>
__global__ void __launch_bounds__(256) gn_relu_pool_kernel(const float* __restrict__ x, const float* __restrict__ scale_shift,
                                                           int n, int t, int h, int w, int c, int t_out,
                                                           int row_stride, int slices, size_t slice_stride,
                                                           __nv_bfloat16* __restrict__ dst, size_t plane_elems,
                                                           int planes) {
    const int quads = c / 4;
    const long long total = 1ll * n * t_out * h * w * quads;
    for (long long i = blockIdx.x * 1ll * blockDim.x + threadIdx.x; i < total; i += 1ll * gridDim.x * blockDim.x) {
        const int q = static_cast<int>(i % quads);
        long long v = i / quads;
        const int ww = static_cast<int>(v % w); v /= w;
        const int hh = static_cast<int>(v % h); v /= h;
        const int to = static_cast<int>(v % t_out);
        const int nn = static_cast<int>(v / t_out);
        float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
        if (scale_shift) {
            const float4* tab = reinterpret_cast<const float4*>(scale_shift + (static_cast<size_t>(nn) * c + 4 * q) * 2);
            const float4 t0 = __ldg(tab), t1 = __ldg(tab + 1);
            sc[0] = t0.x; sh[0] = t0.y; sc[1] = t0.z; sh[1] = t0.w;
            sc[2] = t1.x; sh[2] = t1.y; sc[3] = t1.z; sh[3] = t1.w;
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (POOL != 0) {
            for (int dt = -1; dt <= 1; ++dt) {
                const int ti = 2 * to + dt;
                if (ti < 0 || ti >= t) continue;
                for (int dh = -1; dh <= 1; ++dh) {
                    const int hi = hh + dh;
                    if (hi < 0 || hi >= h) continue;
                    for (int dw = -1; dw <= 1; ++dw) {
                        const int wi = ww + dw;
                        if (wi < 0 || wi >= w) continue;
                        const float4 a = load_sum_slices(
                            reinterpret_cast<const float4*>(x + (((static_cast<size_t>(nn) * t + ti) * h + hi) * w + wi) * row_stride) + q,
                            slice_stride / 4, slices);
                        const float v0 = fmaxf(fmaf(a.x, sc[0], sh[0]), 0.f), v1 = fmaxf(fmaf(a.y, sc[1], sh[1]), 0.f);
                        const float v2 = fmaxf(fmaf(a.z, sc[2], sh[2]), 0.f), v3 = fmaxf(fmaf(a.w, sc[3], sh[3]), 0.f);
                        if (POOL == 2) {
                            acc[0] = fmaxf(acc[0], v0); acc[1] = fmaxf(acc[1], v1);
                            acc[2] = fmaxf(acc[2], v2); acc[3] = fmaxf(acc[3], v3);
                        } else {
                            acc[0] += v0; acc[1] += v1; acc[2] += v2; acc[3] += v3;
                        }
                    }
                }
            }
            if (POOL == 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] *= (1.0f / 27.0f);
            }
        } else {
            const float4 a = load_sum_slices(
                reinterpret_cast<const float4*>(x + (((static_cast<size_t>(nn) * t + to) * h + hh) * w + ww) * row_stride) + q,
                slice_stride / 4, slices);
            acc[0] = fmaxf(fmaf(a.x, sc[0], sh[0]), 0.f);
            acc[1] = fmaxf(fmaf(a.y, sc[1], sh[1]), 0.f);
            acc[2] = fmaxf(fmaf(a.z, sc[2], sh[2]), 0.f);
            acc[3] = fmaxf(fmaf(a.w, sc[3], sh[3]), 0.f);
        }
        const size_t off = ((((static_cast<size_t>(nn) * t_out + to) * h + hh) * w + ww) * c) + 4 * q;
        store_planes4(dst, plane_elems, planes, off, acc);
    }
}

// 4 consecutive channels of a conv output row: fp32 (16 bytes) or bf16 (8 bytes, StemsegConvShape.out_bf16)
template <typename XT>
__device__ __forceinline__ float4 load_quad(const XT* row, int q);
template <>
__device__ __forceinline__ float4 load_quad<float>(const float* row, int q) {
    return __ldg(reinterpret_cast<const float4*>(row) + q);
}
template <>
__device__ __forceinline__ float4 load_quad<__nv_bfloat16>(const __nv_bfloat16* row, int q) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(row) + q);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// Pooled variant with column reuse: one thread produces kPoolW consecutive outputs along W for one channel quad; the
// (t,h)-summed columns are shared between neighbouring outputs (54 loads per 4 outputs instead of 108).
constexpr int kPoolW = 4;

template <bool MAXPOOL, typename XT>
__global__ void __launch_bounds__(256) gn_relu_pool_w4_kernel(const XT* __restrict__ x, const float* __restrict__ scale_shift,
                                                              int n, int t, int h, int w, int c, int t_out, int row_stride,
                                                              __nv_bfloat16* __restrict__ dst, size_t plane_elems,
                                                              int planes) {
    const int quads = c / 4;
    const int wgroups = (w + kPoolW - 1) / kPoolW;
    const long long total = 1ll * n * t_out * h * wgroups * quads;
    for (long long i = blockIdx.x * 1ll * blockDim.x + threadIdx.x; i < total; i += 1ll * gridDim.x * blockDim.x) {
        const int q = static_cast<int>(i % quads);
        long long v = i / quads;
        const int wg = static_cast<int>(v % wgroups); v /= wgroups;
        const int hh = static_cast<int>(v % h); v /= h;
        const int to = static_cast<int>(v % t_out);
        const int nn = static_cast<int>(v / t_out);
        const int w0 = wg * kPoolW;
        float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
        if (scale_shift) {
            const float4* tab = reinterpret_cast<const float4*>(scale_shift + (static_cast<size_t>(nn) * c + 4 * q) * 2);
            const float4 t0 = __ldg(tab), t1 = __ldg(tab + 1);
            sc[0] = t0.x; sh[0] = t0.y; sc[1] = t0.z; sh[1] = t0.w;
            sc[2] = t1.x; sh[2] = t1.y; sc[3] = t1.z; sh[3] = t1.w;
        }
        float col[kPoolW + 2][4];
#pragma unroll
        for (int k = 0; k < kPoolW + 2; ++k) col[k][0] = col[k][1] = col[k][2] = col[k][3] = 0.f;
        for (int dt = -1; dt <= 1; ++dt) {
            const int ti = 2 * to + dt;
            if (ti < 0 || ti >= t) continue;
            for (int dh = -1; dh <= 1; ++dh) {
                const int hi = hh + dh;
                if (hi < 0 || hi >= h) continue;
                const XT* rowp = x + ((static_cast<size_t>(nn) * t + ti) * h + hi) * static_cast<size_t>(w) * row_stride;
#pragma unroll
                for (int k = 0; k < kPoolW + 2; ++k) {
                    const int wi = w0 + k - 1;
                    if (wi < 0 || wi >= w) continue;
                    const float4 a = load_quad<XT>(rowp + static_cast<size_t>(wi) * row_stride, q);
                    const float v0 = fmaxf(fmaf(a.x, sc[0], sh[0]), 0.f), v1 = fmaxf(fmaf(a.y, sc[1], sh[1]), 0.f);
                    const float v2 = fmaxf(fmaf(a.z, sc[2], sh[2]), 0.f), v3 = fmaxf(fmaf(a.w, sc[3], sh[3]), 0.f);
                    if (MAXPOOL) {
                        col[k][0] = fmaxf(col[k][0], v0); col[k][1] = fmaxf(col[k][1], v1);
                        col[k][2] = fmaxf(col[k][2], v2); col[k][3] = fmaxf(col[k][3], v3);
                    } else {
                        col[k][0] += v0; col[k][1] += v1; col[k][2] += v2; col[k][3] += v3;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kPoolW; ++k) {
            const int wo = w0 + k;
            if (wo >= w) break;
            float r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                r[e] = MAXPOOL ? fmaxf(fmaxf(col[k][e], col[k + 1][e]), col[k + 2][e])
                               : (col[k][e] + col[k + 1][e] + col[k + 2][e]) * (1.0f / 27.0f);
            const size_t off = ((((static_cast<size_t>(nn) * t_out + to) * h + hh) * w + wo) * c) + 4 * q;
            store_planes4(dst, plane_elems, planes, off, r);
        }
    }
}

// flat variant of gn_relu_pool_kernel<false> for row_stride == c and one slice (the big 4x layer): no index
// decomposition, two independent 16-byte loads in flight per thread
template <typename XT>
__global__ void __launch_bounds__(256) gn_relu_flat_kernel(const XT* __restrict__ x, const float* __restrict__ scale_shift,
                                                           long long quads_per_sample, int quads, int row_quads,
                                                           long long total_quads, __nv_bfloat16* __restrict__ dst,
                                                           size_t plane_elems, int planes) {
    const long long stride = 1ll * gridDim.x * blockDim.x;
    auto apply = [&](long long i, const float4 a) {
        float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
        if (scale_shift) {
            const long long nn = i / quads_per_sample;
            const int q = static_cast<int>(i % quads);
            const float4* tab = reinterpret_cast<const float4*>(scale_shift) + (nn * quads + q) * 2;
            const float4 t0 = __ldg(tab), t1 = __ldg(tab + 1);
            sc[0] = t0.x; sh[0] = t0.y; sc[1] = t0.z; sh[1] = t0.w;
            sc[2] = t1.x; sh[2] = t1.y; sc[3] = t1.z; sh[3] = t1.w;
        }
        float r[4];
        r[0] = fmaxf(fmaf(a.x, sc[0], sh[0]), 0.f);
        r[1] = fmaxf(fmaf(a.y, sc[1], sh[1]), 0.f);
        r[2] = fmaxf(fmaf(a.z, sc[2], sh[2]), 0.f);
        r[3] = fmaxf(fmaf(a.w, sc[3], sh[3]), 0.f);
        store_planes4(dst, plane_elems, planes, static_cast<size_t>(i) * 4, r);
    };
    if constexpr (std::is_same<XT, __nv_bfloat16>::value) {
        // bf16 rows: 8 channels (one 16-byte load) per item, two independent items in flight per thread; the host
        // guarantees an even number of quads per row
        const long long total_octs = total_quads / 2;
        const int octs = quads / 2, row_octs = row_quads / 2;
        for (long long o0 = blockIdx.x * 1ll * blockDim.x + threadIdx.x; o0 < total_octs; o0 += 2 * stride) {
            const long long o1 = o0 + stride;
            const bool has1 = o1 < total_octs;
            const uint4 r0 = __ldcs(reinterpret_cast<const uint4*>(x) + (o0 / octs) * row_octs + (o0 % octs));
            uint4 r1 = make_uint4(0u, 0u, 0u, 0u);
            if (has1) r1 = __ldcs(reinterpret_cast<const uint4*>(x) + (o1 / octs) * row_octs + (o1 % octs));
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !has1) break;
                const uint4 raw = u == 0 ? r0 : r1;
                const long long o = u == 0 ? o0 : o1;
                const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
                const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
                const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.z));
                const float2 f3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.w));
                apply(2 * o, make_float4(f0.x, f0.y, f1.x, f1.y));
                apply(2 * o + 1, make_float4(f2.x, f2.y, f3.x, f3.y));
            }
        }
    } else {
        for (long long i0 = blockIdx.x * 1ll * blockDim.x + threadIdx.x; i0 < total_quads; i0 += 2 * stride) {
            const long long i1 = i0 + stride;
            const bool has1 = i1 < total_quads;
            // source rows may be wider than the normalised slice (channel slice of a multi-head conv output)
            const float4 a0 = __ldcs(reinterpret_cast<const float4*>(x) + (i0 / quads) * row_quads + (i0 % quads));
            float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has1) a1 = __ldcs(reinterpret_cast<const float4*>(x) + (i1 / quads) * row_quads + (i1 % quads));
            apply(i0, a0);
            if (has1) apply(i1, a1);
        }
    }
}

// K3b: out = z + upsample(y_low)  -> bf16 planes (input of the next 1x1 merge GEMM)
// Uses conv1x1(cat(up(x), f)) == up(W_a x) + W_b f (1x1 conv and trilinear interpolation commute).
__global__ void __launch_bounds__(256) upsample_add_kernel(const float* __restrict__ z, const float* __restrict__ ylow,
                                                           int n, int t, int h, int w, int c, int st, int tl, int hl,
                                                           int wl, __nv_bfloat16* __restrict__ dst, size_t plane_elems,
                                                           int planes) {
    const int quads = c / 4;
    const long long total = 1ll * n * t * h * w * quads;
    for (long long i = blockIdx.x * 1ll * blockDim.x + threadIdx.x; i < total; i += 1ll * gridDim.x * blockDim.x) {
        const int q = static_cast<int>(i % quads);
        long long v = i / quads;
        const int wo = static_cast<int>(v % w); v /= w;
        const int ho = static_cast<int>(v % h); v /= h;
        const int to = static_cast<int>(v % t);
        const int nn = static_cast<int>(v / t);
        const Tri tr = make_tri(nn, to, ho, wo, st, tl, hl, wl, c);
        const size_t off = ((((static_cast<size_t>(nn) * t + to) * h + ho) * w + wo) * c) + 4 * q;
        const float4 zz = __ldg(reinterpret_cast<const float4*>(z + off));
        float acc[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (tr.wgt[k] != 0.f) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(ylow + tr.off[k]) + q);
                acc[0] = fmaf(tr.wgt[k], a.x, acc[0]);
                acc[1] = fmaf(tr.wgt[k], a.y, acc[1]);
                acc[2] = fmaf(tr.wgt[k], a.z, acc[2]);
                acc[3] = fmaf(tr.wgt[k], a.w, acc[3]);
            }
        }
        store_planes4(dst, plane_elems, planes, off, acc);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K4: output heads.  x = z + upsample(y_low) (the conv_4 merge), then J 1x1x1 outputs with per-output activation:
//   act 0: identity (variance / semseg logits)          embedding_decoder.py:137, semseg_decoder.py:116
//   act 1: tanh(0.25 v)                                 embedding_decoder.py:132-133
//   act 2: sigmoid                                      embedding_decoder.py:140, seediness_decoder.py:112
//   coord 0 none / 1 t / 2 y / 3 x added after the activation (embedding_utils.py:29-120)
// Output is channels-first [N][J][T][H][W] fp32 -- the layout the callers index (inference_model.py:137-146).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kHeadMaxC = 256;
constexpr int kHeadJChunk = 8;

// One warp per group of 32 consecutive voxels: for each voxel the 32 lanes read the channel row cooperatively
// (coalesced 16-byte loads), every output is a warp-shuffle reduction, lane i keeps the outputs of voxel i so the
// channels-first stores are coalesced along W.
constexpr int kHeadMaxOut = 64;

__global__ void __launch_bounds__(256) head_out_kernel(const float* __restrict__ z, const float* __restrict__ ylow,
                                                       int n, int t, int h, int w, int c, int st, int tl, int hl,
                                                       int wl, const float* __restrict__ wout /*[J][c]*/,
                                                       const float* __restrict__ bout /*[J] or null*/,
                                                       const int* __restrict__ act, const int* __restrict__ coord,
                                                       int j_total, float x_abs, float y_abs, float t_abs,
                                                       float* __restrict__ out) {
    extern __shared__ float s_w[];          // [J][c] + [J] bias
    float* s_b = s_w + j_total * c;
    for (int i = threadIdx.x; i < j_total * c; i += blockDim.x) s_w[i] = wout[i];
    for (int i = threadIdx.x; i < j_total; i += blockDim.x) s_b[i] = bout ? bout[i] : 0.f;
    __syncthreads();
    const long long spatial = 1ll * t * h * w;
    const long long total = 1ll * n * spatial;
    const int quads = c / 4;                               // <= 64
    const int lane = threadIdx.x & 31;
    const long long warps = 1ll * gridDim.x * (blockDim.x >> 5);
    const long long groups = (total + 31) / 32;
    for (long long g = 1ll * blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < groups; g += warps) {
        const long long vbase = g * 32;
        for (int j0 = 0; j0 < j_total; j0 += kHeadJChunk) {
            float mine[kHeadJChunk];
#pragma unroll
            for (int j = 0; j < kHeadJChunk; ++j) mine[j] = 0.f;
            for (int i = 0; i < 32; ++i) {
                const long long v = vbase + i;
                if (v >= total) break;                       // warp-uniform
                long long r = v;
                const int wo = static_cast<int>(r % w); r /= w;
                const int ho = static_cast<int>(r % h); r /= h;
                const int to = static_cast<int>(r % t);
                const int nn = static_cast<int>(r / t);
                const Tri tr = make_tri(nn, to, ho, wo, st, tl, hl, wl, c);
                float acc[kHeadJChunk];
#pragma unroll
                for (int j = 0; j < kHeadJChunk; ++j) acc[j] = 0.f;
                for (int q = lane; q < quads; q += 32) {
                    const float4 zz = __ldg(reinterpret_cast<const float4*>(z + static_cast<size_t>(v) * c) + q);
                    float xv[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (tr.wgt[k] != 0.f) {
                            const float4 a = __ldg(reinterpret_cast<const float4*>(ylow + tr.off[k]) + q);
                            xv[0] = fmaf(tr.wgt[k], a.x, xv[0]);
                            xv[1] = fmaf(tr.wgt[k], a.y, xv[1]);
                            xv[2] = fmaf(tr.wgt[k], a.z, xv[2]);
                            xv[3] = fmaf(tr.wgt[k], a.w, xv[3]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < kHeadJChunk; ++j) {
                        if (j0 + j < j_total) {
                            const float4 wr = *reinterpret_cast<const float4*>(s_w + (j0 + j) * c + 4 * q);
                            acc[j] = fmaf(xv[0], wr.x, acc[j]);
                            acc[j] = fmaf(xv[1], wr.y, acc[j]);
                            acc[j] = fmaf(xv[2], wr.z, acc[j]);
                            acc[j] = fmaf(xv[3], wr.w, acc[j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < kHeadJChunk; ++j) {
                    if (j0 + j < j_total) {
                        float a = acc[j];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                        if (lane == i) mine[j] = a;
                    }
                }
            }
            const long long v = vbase + lane;
            if (v < total) {
                long long r = v;
                const int wo = static_cast<int>(r % w); r /= w;
                const int ho = static_cast<int>(r % h); r /= h;
                const int to = static_cast<int>(r % t);
                const int nn = static_cast<int>(r / t);
#pragma unroll
                for (int j = 0; j < kHeadJChunk; ++j) {
                    const int jj = j0 + j;
                    if (jj < j_total) {
                        float val = mine[j] + s_b[jj];
                        const int a = act[jj];
                        if (a == 1) val = tanhf(0.25f * val);
                        else if (a == 2) val = 1.0f / (1.0f + expf(-val));
                        const int cd = coord[jj];
                        if (cd == 1) val += linspace_value(t_abs, t, to);
                        else if (cd == 2) val += linspace_value(y_abs, h, ho);
                        else if (cd == 3) val += linspace_value(x_abs, w, wo);
                        out[(static_cast<size_t>(nn) * j_total + jj) * spatial + (v - static_cast<long long>(nn) * spatial)] = val;
                    }
                }
            }
        }
    }
}

// p_low[v][j] = sum_c W_out[j][c] * y_low[v][c]: the output convs applied at the LOW resolution (they commute with the
// trilinear upsampling), consumed by the fused head epilogue of the conv_4 GEMM (conv_tc.cu, epi_mode 1).
__global__ void __launch_bounds__(256) head_lowres_kernel(const float* __restrict__ ylow, long long voxels, int c,
                                                          const float* __restrict__ wout, int j_total,
                                                          float* __restrict__ p_low) {
    const long long total = voxels * j_total;
    const int quads = c / 4;
    for (long long i = blockIdx.x * 1ll * blockDim.x + threadIdx.x; i < total; i += 1ll * gridDim.x * blockDim.x) {
        const int j = static_cast<int>(i % j_total);
        const long long v = i / j_total;
        const float4* row = reinterpret_cast<const float4*>(ylow + static_cast<size_t>(v) * c);
        const float4* wr = reinterpret_cast<const float4*>(wout + static_cast<size_t>(j) * c);
        float acc = 0.f;
        for (int q = 0; q < quads; ++q) {
            const float4 a = __ldg(row + q), b = __ldg(wr + q);
            acc = fmaf(a.x, b.x, acc);
            acc = fmaf(a.y, b.y, acc);
            acc = fmaf(a.z, b.z, acc);
            acc = fmaf(a.w, b.w, acc);
        }
        p_low[i] = acc;
    }
}

inline unsigned grid_for(long long total, int block, int waves = 8) {
    long long blocks = (total + block - 1) / block;
    const long long cap = static_cast<long long>(device_sm_count()) * waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<unsigned>(blocks);
}

}  // namespace
}  // namespace stemseg

using namespace stemseg;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int32_t stemseg_pack_activation(const float* src, int64_t stride_n, int64_t stride_c, int64_t stride_t,
                                           int32_t n, int32_t c, int32_t t, int32_t hw, void* dst_planes,
                                           int32_t planes, void* stream_) {
    SS_REQUIRE(src && dst_planes, "pack_activation: null pointer");
    SS_REQUIRE(valid_planes(planes, true), "pack_activation: planes must be 1, 2 or STEMSEG_PLANES_FP16");
    SS_REQUIRE(n >= 1 && c >= 2 && c % 2 == 0 && t >= 1 && hw >= 1, "pack_activation: bad shape");
    SS_REQUIRE(1ll * n * t <= 65535, "pack_activation: n*t too large");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t plane_elems = static_cast<size_t>(n) * t * hw * c;
    dim3 grid((hw + kPackS - 1) / kPackS, (c + kPackC - 1) / kPackC, n * t);
    pack_ncthw_kernel<<<grid, 256, 0, stream>>>(src, stride_n, stride_c, stride_t, c, t, hw,
                                                static_cast<__nv_bfloat16*>(dst_planes), plane_elems, planes);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_pack_conv_weight(const float* src, int32_t cout, int32_t cin_total, int32_t cin_begin,
                                            int32_t cin_count, int32_t taps, void* dst_planes, int32_t row_begin,
                                            int32_t rows_total, int32_t planes, void* stream_) {
    SS_REQUIRE(src && dst_planes, "pack_conv_weight: null pointer");
    SS_REQUIRE(valid_planes(planes, true), "pack_conv_weight: planes must be 1, 2 or STEMSEG_PLANES_FP16");
    SS_REQUIRE(cout >= 1 && cin_count >= 1 && cin_begin >= 0 && cin_begin + cin_count <= cin_total && taps >= 1 &&
                   row_begin >= 0 && row_begin + cout <= rows_total,
               "pack_conv_weight: bad shape");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t plane_elems = static_cast<size_t>(rows_total) * taps * cin_count;
    const long long total = 1ll * cout * taps * cin_count;
    pack_weight_kernel<<<grid_for(total, 256), 256, 0, stream>>>(src, cout, cin_total, cin_begin, cin_count, taps,
                                                                 static_cast<__nv_bfloat16*>(dst_planes), row_begin,
                                                                 plane_elems, planes);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

static int stats_chunk_voxels(int64_t spatial) {
    // enough blocks to fill the device even for the few-thousand-voxel layers
    long long chunk = (spatial + 4ll * device_sm_count() - 1) / (4ll * device_sm_count());
    if (chunk < 8) chunk = 8;
    if (chunk > kStatsMaxChunk) chunk = kStatsMaxChunk;
    return static_cast<int>(chunk);
}

extern "C" int32_t stemseg_pack_conv_weight_dgrad(const float* src, int32_t cout, int32_t cin_total, int32_t cin_begin,
                                                  int32_t cin_count, int32_t taps, void* dst_planes, int32_t planes,
                                                  void* stream_) {
    SS_REQUIRE(src && dst_planes, "pack_conv_weight_dgrad: null pointer");
    SS_REQUIRE(planes == 1 || planes == 2, "pack_conv_weight_dgrad: planes must be 1 or 2");
    SS_REQUIRE(cout >= 1 && cin_count >= 1 && cin_begin >= 0 && cin_begin + cin_count <= cin_total && taps >= 1,
               "pack_conv_weight_dgrad: bad shape");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t plane_elems = static_cast<size_t>(cin_count) * taps * cout;
    pack_weight_dgrad_kernel<<<grid_for(static_cast<long long>(plane_elems), 256), 256, 0, stream>>>(
        src, cout, cin_total, cin_begin, cin_count, taps, static_cast<__nv_bfloat16*>(dst_planes), plane_elems, planes);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" size_t stemseg_group_norm_workspace_bytes(int32_t n, int64_t spatial, int32_t c) {
    const long long chunks = (spatial + 7) / 8;          // upper bound over every chunk size that may be chosen
    return align_up(static_cast<size_t>(n) * chunks * c * 2 * sizeof(float), 256);
}

extern "C" int32_t stemseg_group_norm_stats(float* x, int32_t row_stride, int32_t slices, int32_t n,
                                            int64_t spatial, int32_t c, int32_t channels_per_group, float eps,
                                            const float* gamma, const float* beta, float* scale_shift,
                                            float* mean_rstd, void* workspace, size_t workspace_bytes, void* stream_) {
    SS_REQUIRE(x && scale_shift && gamma && beta && workspace, "group_norm_stats: null pointer");
    SS_REQUIRE(n >= 1 && spatial >= 1 && c >= 4 && c % 4 == 0 && c <= 1024, "group_norm_stats: bad shape");
    SS_REQUIRE(channels_per_group >= 1 && channels_per_group <= 512 && c % channels_per_group == 0,
               "group_norm_stats: bad group size");
    SS_REQUIRE(slices >= 1 && slices <= 27, "group_norm_stats: slices out of range");
    SS_REQUIRE(row_stride >= c && row_stride % 4 == 0, "group_norm_stats: bad row stride");
    SS_REQUIRE(aligned16(x), "group_norm_stats: x must be 16-byte aligned");
    const int chunk_voxels = stats_chunk_voxels(spatial);
    const int chunks = static_cast<int>((spatial + chunk_voxels - 1) / chunk_voxels);
    const size_t need = align_up(static_cast<size_t>(n) * chunks * c * 2 * sizeof(float), 256);
    if (workspace_bytes < need) {
        set_error("group_norm_stats: workspace %zu < %zu bytes", workspace_bytes, need);
        return STEMSEG_ERR_WORKSPACE;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int quads = c / 4;
    const int rows = 256 / quads >= 1 ? 256 / quads : 1;
    const int threads = quads * rows;
    const size_t smem = static_cast<size_t>(rows) * c * 2 * sizeof(float);
    SS_REQUIRE(threads <= 1024 && smem <= 48 * 1024, "group_norm_stats: channel count %d unsupported", c);
    gn_partial_kernel<<<dim3(chunks, n), threads, smem, stream>>>(x, spatial, c, row_stride, slices,
                                                                  static_cast<size_t>(n) * spatial * row_stride,
                                                                  chunk_voxels, static_cast<float*>(workspace), chunks);
    gn_finalize_kernel<<<dim3(c / channels_per_group, n), 512, 0, stream>>>(
        static_cast<const float*>(workspace), static_cast<size_t>(c) * chunks * 2, chunks, c, channels_per_group,
        spatial, eps, gamma, beta, scale_shift, mean_rstd);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_group_norm_finalize(const float* partial, int64_t partial_sample_stride, int32_t chunks,
                                               int32_t n, int64_t spatial, int32_t c, int32_t channels_per_group,
                                               float eps, const float* gamma, const float* beta, float* scale_shift,
                                               float* mean_rstd, void* stream_) {
    SS_REQUIRE(partial && gamma && beta && scale_shift, "group_norm_finalize: null pointer");
    SS_REQUIRE(n >= 1 && spatial >= 1 && chunks >= 1 && c >= 1, "group_norm_finalize: bad shape");
    SS_REQUIRE(channels_per_group >= 1 && channels_per_group <= 512 && c % channels_per_group == 0,
               "group_norm_finalize: bad group size");
    SS_REQUIRE((reinterpret_cast<uintptr_t>(partial) & 7) == 0, "group_norm_finalize: partial must be 8-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    gn_finalize_kernel<<<dim3(c / channels_per_group, n), 512, 0, stream>>>(
        partial, static_cast<size_t>(partial_sample_stride), chunks, c, channels_per_group, spatial, eps, gamma, beta,
        scale_shift, mean_rstd);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

template <typename XT>
static int32_t norm_relu_pool_impl(const XT* x, int32_t row_stride, int32_t slices, const float* scale_shift,
                                          int32_t n, int32_t t, int32_t h, int32_t w, int32_t c, int32_t pool,
                                          void* dst_planes, int32_t planes, void* stream_) {
    SS_REQUIRE(x && dst_planes, "norm_relu_pool: null pointer");
    SS_REQUIRE(valid_planes(planes, true), "norm_relu_pool: planes must be 1, 2 or STEMSEG_PLANES_FP16");
    SS_REQUIRE(slices >= 1 && slices <= 27, "norm_relu_pool: slices out of range");
    SS_REQUIRE(pool >= 0 && pool <= 2, "norm_relu_pool: pool must be 0 (none), 1 (average) or 2 (max)");
    SS_REQUIRE(row_stride >= c && row_stride % 4 == 0, "norm_relu_pool: bad row stride");
    SS_REQUIRE(n >= 1 && t >= 1 && h >= 1 && w >= 1 && c >= 4 && c % 4 == 0, "norm_relu_pool: bad shape");
    SS_REQUIRE(aligned16(x) && aligned16(dst_planes) && aligned16(scale_shift),
               "norm_relu_pool: pointers must be 16-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int t_out = pool ? (t - 1) / 2 + 1 : t;
    const size_t plane_elems = static_cast<size_t>(n) * t_out * h * w * c;
    const size_t slice_stride = static_cast<size_t>(n) * t * h * w * row_stride;
    const long long total = 1ll * n * t_out * h * w * (c / 4);
    auto* dst = static_cast<__nv_bfloat16*>(dst_planes);
    if (pool == 1 && slices == 1)
        gn_relu_pool_w4_kernel<false, XT><<<grid_for((total + kPoolW - 1) / kPoolW, 256, 16), 256, 0, stream>>>(
            x, scale_shift, n, t, h, w, c, t_out, row_stride, dst, plane_elems, planes);
    else if (pool == 2 && slices == 1)
        gn_relu_pool_w4_kernel<true, XT><<<grid_for((total + kPoolW - 1) / kPoolW, 256, 16), 256, 0, stream>>>(
            x, scale_shift, n, t, h, w, c, t_out, row_stride, dst, plane_elems, planes);
    else if (slices == 1) {
        if (!std::is_same<XT, float>::value && (c % 8 != 0 || row_stride % 8 != 0)) {
            set_error("norm_relu_pool: bf16 rows need channel counts that are multiples of 8");
            return STEMSEG_ERR_INVALID_ARGUMENT;
        }
        const long long items = std::is_same<XT, float>::value ? total : total / 2;
        gn_relu_flat_kernel<XT><<<grid_for((items + 1) / 2, 256, 8), 256, 0, stream>>>(
            x, scale_shift, 1ll * t * h * w * (c / 4), c / 4, row_stride / 4, total, dst, plane_elems, planes);
    }
    else if constexpr (!std::is_same<XT, float>::value) {
        set_error("norm_relu_pool: a bf16 conv output must be a single slice");
        return STEMSEG_ERR_INVALID_ARGUMENT;
    } else if (pool == 1)
        gn_relu_pool_kernel<1><<<grid_for(total, 256, 16), 256, 0, stream>>>(
            x, scale_shift, n, t, h, w, c, t_out, row_stride, slices, slice_stride, dst, plane_elems, planes);
    else if (pool == 2)
        gn_relu_pool_kernel<2><<<grid_for(total, 256, 16), 256, 0, stream>>>(
            x, scale_shift, n, t, h, w, c, t_out, row_stride, slices, slice_stride, dst, plane_elems, planes);
    else
        gn_relu_pool_kernel<0><<<grid_for(total, 256, 16), 256, 0, stream>>>(
            x, scale_shift, n, t, h, w, c, t_out, row_stride, slices, slice_stride, dst, plane_elems, planes);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_norm_relu_pool(const float* x, int32_t row_stride, int32_t slices, const float* scale_shift,
                                          int32_t n, int32_t t, int32_t h, int32_t w, int32_t c, int32_t pool,
                                          void* dst_planes, int32_t planes, void* stream_) {
    return norm_relu_pool_impl<float>(x, row_stride, slices, scale_shift, n, t, h, w, c, pool, dst_planes, planes, stream_);
}

extern "C" int32_t stemseg_norm_relu_pool_bf16in(const void* x, int32_t row_stride, int32_t slices,
                                                 const float* scale_shift, int32_t n, int32_t t, int32_t h, int32_t w,
                                                 int32_t c, int32_t pool, void* dst_planes, int32_t planes, void* stream_) {
    SS_REQUIRE(slices == 1, "norm_relu_pool_bf16in: slices must be 1");
    return norm_relu_pool_impl<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(x), row_stride, slices, scale_shift, n, t, h,
                                              w, c, pool, dst_planes, planes, stream_);
}

extern "C" int32_t stemseg_upsample_add(const float* z, const float* y_low, int32_t n, int32_t t, int32_t h,
                                        int32_t w, int32_t c, int32_t t_scale, void* dst_planes, int32_t planes,
                                        void* stream_) {
    SS_REQUIRE(z && y_low && dst_planes, "upsample_add: null pointer");
    SS_REQUIRE(planes == 1 || planes == 2, "upsample_add: planes must be 1 or 2");
    SS_REQUIRE(t_scale == 1 || t_scale == 2, "upsample_add: temporal scale must be 1 or 2");
    SS_REQUIRE(n >= 1 && t >= 1 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && t % t_scale == 0 && c >= 4 &&
                   c % 4 == 0,
               "upsample_add: bad shape");
    SS_REQUIRE(aligned16(z) && aligned16(y_low) && aligned16(dst_planes), "upsample_add: pointers must be 16-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t plane_elems = static_cast<size_t>(n) * t * h * w * c;
    const long long total = 1ll * n * t * h * w * (c / 4);
    upsample_add_kernel<<<grid_for(total, 256, 16), 256, 0, stream>>>(
        z, y_low, n, t, h, w, c, t_scale, t / t_scale, h / 2, w / 2, static_cast<__nv_bfloat16*>(dst_planes),
        plane_elems, planes);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_head_output(const float* z, const float* y_low, int32_t n, int32_t t, int32_t h,
                                       int32_t w, int32_t c, int32_t t_scale, const float* out_weight,
                                       const float* out_bias, const int32_t* activation, const int32_t* coordinate,
                                       int32_t n_out, float time_scale, float* out, void* stream_) {
    SS_REQUIRE(z && y_low && out_weight && activation && coordinate && out, "head_output: null pointer");
    SS_REQUIRE(t_scale == 1 || t_scale == 2, "head_output: temporal scale must be 1 or 2");
    SS_REQUIRE(n >= 1 && t >= 1 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && t % t_scale == 0 && c >= 4 &&
                   c % 4 == 0 && c <= kHeadMaxC,
               "head_output: bad shape");
    SS_REQUIRE(n_out >= 1 && n_out <= 64, "head_output: n_out %d out of range [1,64]", n_out);
    SS_REQUIRE(aligned16(z) && aligned16(y_low), "head_output: pointers must be 16-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t smem = (static_cast<size_t>(n_out) * c + n_out) * sizeof(float);
    SS_CUDA_OK(cudaFuncSetAttribute(head_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    // embedding_utils.py:31-37: x in [-max(1, W/H), +], y in [-max(1, H/W), +], t in [-time_scale, +]
    const float x_abs = fmaxf(1.0f, static_cast<float>(static_cast<double>(w) / static_cast<double>(h)));
    const float y_abs = fmaxf(1.0f, static_cast<float>(static_cast<double>(h) / static_cast<double>(w)));
    const long long total = 1ll * n * t * h * w;
    head_out_kernel<<<grid_for((total + 31) / 32 * 32 / 8, 32, 8), 256, smem, stream>>>(z, y_low, n, t, h, w, c, t_scale, t / t_scale,
                                                                      h / 2, w / 2, out_weight, out_bias, activation,
                                                                      coordinate, n_out, x_abs, y_abs, time_scale,
                                                                      out);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_head_lowres(const float* y_low, int64_t voxels, int32_t c, const float* out_weight,
                                       int32_t n_out, float* p_low, void* stream_) {
    SS_REQUIRE(y_low && out_weight && p_low, "head_lowres: null pointer");
    SS_REQUIRE(voxels >= 1 && c >= 4 && c % 4 == 0 && n_out >= 1 && n_out <= 64, "head_lowres: bad shape");
    SS_REQUIRE(aligned16(y_low) && aligned16(out_weight), "head_lowres: pointers must be 16-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    head_lowres_kernel<<<grid_for(voxels * n_out, 256, 16), 256, 0, stream>>>(y_low, voxels, c, out_weight, n_out, p_low);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}
