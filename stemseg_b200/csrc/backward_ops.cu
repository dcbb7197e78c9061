// Backward-pass kernels of the decoder heads (sm_100a) -- SURVEY.md §8f rank 3.
//
// The GEMM-shaped parts of the backward pass reuse the tcgen05 convolution kernel (conv_tc.cu):
//   * dgrad of a stride-1 / pad-1 convolution = the same convolution of dy with the flipped, transposed weights;
//   * wgrad: dW[co][tap][ci] = sum_p dyT[co][p] * xT[ci][p + delta(tap)] on zero-padded, TRANSPOSED planes
//     ([C][(T+2)(H+2)(W+2)] bf16), i.e. 27 plain K-major GEMMs whose B operand is read at a constant offset.
// This file holds the HBM-bound rest: output-head backward, the adjoint of the trilinear up-sampling, AvgPool/ReLU
// backward, GroupNorm backward (reductions + apply), fp32 -> bf16 plane conversion and the padded transposes.
// All reductions are deterministic (fixed partitioning, fixed order).
#include "common.cuh"
#include "trilinear.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace stemseg {
namespace {

__device__ __forceinline__ void split_bf16_b(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

inline unsigned grid_cap(long long total, int block, int waves = 16) {
    long long blocks = (total + block - 1) / block;
    const long long cap = static_cast<long long>(device_sm_count()) * waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<unsigned>(blocks);
}

// ---------------------------------------------------------------------------------------------------------------
// Output heads backward.  Forward (decoder_ops.cu head_out / conv_tc epilogue):
//   x = z + up(y_low);  pre_j = W_j . x + b_j;  out_j = act_j(pre_j) + coord_j
// Given g = dL/d out [n][J][t][h][w]:  dpre_j = g_j act_j'(pre_j);  dx = sum_j dpre_j W_j;
//   dW_j = sum_v dpre_j x;  db_j = sum_v dpre_j.
// One warp per group of 32 voxels, lanes own channel quads (c <= 128: one quad per lane; c = 256: two).
// Per-warp partial dW / db are reduced per block in shared memory and written as [blocks][J][c+1] partials.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kHbJ = 8;          // outputs handled per pass over the voxels
constexpr int kHbWarps = 8;

__global__ void __launch_bounds__(kHbWarps * 32) head_backward_kernel(
    const float* __restrict__ z, const float* __restrict__ ylow, int n, int t, int h, int w, int c, int st, int tl, int hl,
    int wl, const float* __restrict__ wout, const float* __restrict__ bout, const int* __restrict__ act,
    const float* __restrict__ g, int j_total, int j0, float* __restrict__ dx /*[n][t][h][w][c], accumulated over passes*/,
    float* __restrict__ partial /*[gridDim.x][kHbJ][c + 1]*/) {
    extern __shared__ float s_mem[];
    float* s_w = s_mem;                               // [kHbJ][c]
    float* s_acc = s_mem + kHbJ * c;                  // [kHbWarps][kHbJ][c + 1]
    const int jn = min(kHbJ, j_total - j0);
    for (int i = threadIdx.x; i < jn * c; i += blockDim.x) s_w[i] = wout[(j0 + i / c) * c + i % c];
    __syncthreads();
    const long long spatial = 1ll * t * h * w;
    const long long total = 1ll * n * spatial;
    const int quads = c / 4;                          // <= 64
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float accw[kHbJ][2][4];                           // dW partials for this lane's (<= 2) channel quads
    float accb[kHbJ];
#pragma unroll
    for (int j = 0; j < kHbJ; ++j) {
        accb[j] = 0.f;
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) accw[j][qq][0] = accw[j][qq][1] = accw[j][qq][2] = accw[j][qq][3] = 0.f;
    }
    const long long warps = 1ll * gridDim.x * kHbWarps;
    for (long long v = 1ll * blockIdx.x * kHbWarps + warp; v < total; v += warps) {
        long long r = v;
        const int wo = static_cast<int>(r % w); r /= w;
        const int ho = static_cast<int>(r % h); r /= h;
        const int to = static_cast<int>(r % t);
        const int nn = static_cast<int>(r / t);
        const Tri tr = make_tri(nn, to, ho, wo, st, tl, hl, wl, c);
        float xq[2][4];
        float dot[kHbJ];
#pragma unroll
        for (int j = 0; j < kHbJ; ++j) dot[j] = 0.f;
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
            const int q = lane + 32 * qq;
            xq[qq][0] = xq[qq][1] = xq[qq][2] = xq[qq][3] = 0.f;
            if (q < quads) {
                const float4 zz = __ldg(reinterpret_cast<const float4*>(z + static_cast<size_t>(v) * c) + q);
                xq[qq][0] = zz.x; xq[qq][1] = zz.y; xq[qq][2] = zz.z; xq[qq][3] = zz.w;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (tr.wgt[k] != 0.f) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(ylow + tr.off[k]) + q);
                        xq[qq][0] = fmaf(tr.wgt[k], a.x, xq[qq][0]);
                        xq[qq][1] = fmaf(tr.wgt[k], a.y, xq[qq][1]);
                        xq[qq][2] = fmaf(tr.wgt[k], a.z, xq[qq][2]);
                        xq[qq][3] = fmaf(tr.wgt[k], a.w, xq[qq][3]);
                    }
                }
#pragma unroll
                for (int j = 0; j < kHbJ; ++j) {
                    if (j < jn) {
                        const float4 wr = *reinterpret_cast<const float4*>(s_w + j * c + 4 * q);
                        dot[j] = fmaf(xq[qq][0], wr.x, dot[j]);
                        dot[j] = fmaf(xq[qq][1], wr.y, dot[j]);
                        dot[j] = fmaf(xq[qq][2], wr.z, dot[j]);
                        dot[j] = fmaf(xq[qq][3], wr.w, dot[j]);
                    }
                }
            }
        }
        float dpre[kHbJ];
#pragma unroll
        for (int j = 0; j < kHbJ; ++j) {
            dpre[j] = 0.f;
            if (j < jn) {
                float a = dot[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                const int jj = j0 + j;
                const float pre = a + (bout ? bout[jj] : 0.f);
                const float gg = __ldg(g + (static_cast<size_t>(nn) * j_total + jj) * spatial +
                                       (v - static_cast<long long>(nn) * spatial));
                float d = gg;
                const int ac = act[jj];
                if (ac == 1) {
                    const float th = tanhf(0.25f * pre);
                    d = gg * 0.25f * (1.f - th * th);
                } else if (ac == 2) {
                    const float sg = 1.0f / (1.0f + expf(-pre));
                    d = gg * sg * (1.f - sg);
                }
                dpre[j] = d;
                accb[j] += d;                     // identical in every lane; lane 0's copy is used
            }
        }
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
            const int q = lane + 32 * qq;
            if (q < quads) {
                float4* dst = reinterpret_cast<float4*>(dx + static_cast<size_t>(v) * c) + q;
                float4 o = j0 == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : *dst;
#pragma unroll
                for (int j = 0; j < kHbJ; ++j) {
                    if (j < jn) {
                        const float4 wr = *reinterpret_cast<const float4*>(s_w + j * c + 4 * q);
                        o.x = fmaf(dpre[j], wr.x, o.x); o.y = fmaf(dpre[j], wr.y, o.y);
                        o.z = fmaf(dpre[j], wr.z, o.z); o.w = fmaf(dpre[j], wr.w, o.w);
                        accw[j][qq][0] = fmaf(dpre[j], xq[qq][0], accw[j][qq][0]);
                        accw[j][qq][1] = fmaf(dpre[j], xq[qq][1], accw[j][qq][1]);
                        accw[j][qq][2] = fmaf(dpre[j], xq[qq][2], accw[j][qq][2]);
                        accw[j][qq][3] = fmaf(dpre[j], xq[qq][3], accw[j][qq][3]);
                    }
                }
                *dst = o;
            }
        }
    }
    // block reduction of the dW / db partials (fixed order over the warps)
    const int stride = c + 1;
#pragma unroll
    for (int j = 0; j < kHbJ; ++j) {
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
            const int q = lane + 32 * qq;
            if (q < quads)
#pragma unroll
                for (int e = 0; e < 4; ++e) s_acc[(warp * kHbJ + j) * stride + 4 * q + e] = accw[j][qq][e];
        }
        if (lane == 0) s_acc[(warp * kHbJ + j) * stride + c] = accb[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHbJ * stride; i += blockDim.x) {
        float a = 0.f;
        for (int wv = 0; wv < kHbWarps; ++wv) a += s_acc[wv * kHbJ * stride + i];
        partial[static_cast<size_t>(blockIdx.x) * kHbJ * stride + i] = a;
    }
}

// out[i] = sum_b partial[b][i] (fixed order), double accumulation
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ partial, int rows, long long cols,
                                                          long long row_stride, float* __restrict__ out, int accumulate) {
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < cols; i += 256ll * gridDim.x) {
        double a = 0.0;
        for (int b = 0; b < rows; ++b) a += partial[static_cast<size_t>(b) * row_stride + i];
        out[i] = accumulate ? out[i] + static_cast<float>(a) : static_cast<float>(a);
    }
}

// same result, parallel over rows as well: block = 32 columns x 8 row groups; partial sums of the row groups are added
// in a fixed order (deterministic)
__global__ void __launch_bounds__(256) reduce_rows_tiled_kernel(const float* __restrict__ partial, int rows, long long cols,
                                                                long long row_stride, float* __restrict__ out) {
    __shared__ double s[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long col = blockIdx.x * 32ll + tx;
    double a = 0.0;
    if (col < cols)
        for (int b = ty; b < rows; b += 8) a += static_cast<double>(partial[static_cast<size_t>(b) * row_stride + col]);
    s[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && col < cols) {
        double tot = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) tot += s[r][tx];
        out[col] = static_cast<float>(tot);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Adjoint of the trilinear up-sampling: dlow[i] = sum_o w(o -> i) dhigh[o]  (gather form, deterministic)
// ---------------------------------------------------------------------------------------------------------------
struct AxisIn {
    int o[4];
    float w[4];
    int cnt;
};
// high-resolution positions (and weights) that read low-resolution index i along one axis
__device__ __forceinline__ AxisIn axis_adjoint(int i, int scale, int low_size) {
    AxisIn r;
    r.cnt = 0;
    if (scale == 1) {
        r.o[0] = i; r.w[0] = 1.f; r.cnt = 1;
        return r;
    }
    const int high = low_size * 2;
    for (int o = 2 * i - 2; o <= 2 * i + 3; ++o) {
        if (o < 0 || o >= high) continue;
        const Tap tp = axis_tap(o, 2, low_size);
        float wsum = 0.f;
        if (tp.i0 == i) wsum += tp.w0;
        if (tp.i1 == i) wsum += tp.w1;
        if (wsum != 0.f && r.cnt < 4) { r.o[r.cnt] = o; r.w[r.cnt] = wsum; ++r.cnt; }
    }
    return r;
}

__global__ void __launch_bounds__(256) upsample_transpose_kernel(const float* __restrict__ dhigh, int n, int t, int h,
                                                                 int w, int c, int st, int tl, int hl, int wl,
                                                                 float* __restrict__ dlow) {
    const int quads = c / 4;
    const long long total = 1ll * n * tl * hl * wl * quads;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
        const int q = static_cast<int>(i % quads);
        long long v = i / quads;
        const int wi = static_cast<int>(v % wl); v /= wl;
        const int hi = static_cast<int>(v % hl); v /= hl;
        const int ti = static_cast<int>(v % tl);
        const int nn = static_cast<int>(v / tl);
        const AxisIn at = axis_adjoint(ti, st, tl), ah = axis_adjoint(hi, 2, hl), aw = axis_adjoint(wi, 2, wl);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int a = 0; a < at.cnt; ++a)
            for (int b = 0; b < ah.cnt; ++b)
                for (int d = 0; d < aw.cnt; ++d) {
                    const float wgt = at.w[a] * ah.w[b] * aw.w[d];
                    const float4 x = __ldg(reinterpret_cast<const float4*>(
                                               dhigh + (((static_cast<size_t>(nn) * t + at.o[a]) * h + ah.o[b]) * w + aw.o[d]) * c) + q);
                    acc.x = fmaf(wgt, x.x, acc.x); acc.y = fmaf(wgt, x.y, acc.y);
                    acc.z = fmaf(wgt, x.z, acc.z); acc.w = fmaf(wgt, x.w, acc.w);
                }
        *(reinterpret_cast<float4*>(dlow + (((static_cast<size_t>(nn) * tl + ti) * hl + hi) * wl + wi) * c) + q) = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// AvgPool3d(3, (2,1,1), 1) + ReLU backward: dn = [scale*y + shift > 0] * (1/27) sum_{outputs covering the voxel} dp
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_relu_backward_kernel(const float* __restrict__ dp, const float* __restrict__ y,
                                                                 const float* __restrict__ scale_shift, int n, int t,
                                                                 int h, int w, int c, int t_out, int pool,
                                                                 float* __restrict__ dn) {
    const int quads = c / 4;
    const long long total = 1ll * n * t * h * w * quads;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
        const int q = static_cast<int>(i % quads);
        long long v = i / quads;
        const int ww = static_cast<int>(v % w); v /= w;
        const int hh = static_cast<int>(v % h); v /= h;
        const int tt = static_cast<int>(v % t);
        const int nn = static_cast<int>(v / t);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pool) {
            for (int to = 0; to < t_out; ++to) {
                if (2 * to - tt > 1 || 2 * to - tt < -1) continue;
                for (int dh = -1; dh <= 1; ++dh) {
                    const int ho = hh + dh;
                    if (ho < 0 || ho >= h) continue;
                    for (int dw = -1; dw <= 1; ++dw) {
                        const int wo = ww + dw;
                        if (wo < 0 || wo >= w) continue;
                        const float4 a = __ldg(reinterpret_cast<const float4*>(
                                                   dp + (((static_cast<size_t>(nn) * t_out + to) * h + ho) * w + wo) * c) + q);
                        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
                    }
                }
            }
            acc.x *= (1.0f / 27.0f); acc.y *= (1.0f / 27.0f); acc.z *= (1.0f / 27.0f); acc.w *= (1.0f / 27.0f);
        } else {
            acc = __ldg(reinterpret_cast<const float4*>(dp) + i);
        }
        const float4 yy = __ldg(reinterpret_cast<const float4*>(y) + i);
        float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
        if (scale_shift) {
            const float4* tab = reinterpret_cast<const float4*>(scale_shift + (static_cast<size_t>(nn) * c + 4 * q) * 2);
            const float4 t0 = __ldg(tab), t1 = __ldg(tab + 1);
            sc[0] = t0.x; sh[0] = t0.y; sc[1] = t0.z; sh[1] = t0.w;
            sc[2] = t1.x; sh[2] = t1.y; sc[3] = t1.z; sh[3] = t1.w;
        }
        float4 o;
        o.x = fmaf(yy.x, sc[0], sh[0]) > 0.f ? acc.x : 0.f;
        o.y = fmaf(yy.y, sc[1], sh[1]) > 0.f ? acc.y : 0.f;
        o.z = fmaf(yy.z, sc[2], sh[2]) > 0.f ? acc.z : 0.f;
        o.w = fmaf(yy.w, sc[3], sh[3]) > 0.f ? acc.w : 0.f;
        reinterpret_cast<float4*>(dn)[i] = o;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm backward.  x_hat = (y - mean) rstd;  per channel: dbeta = sum dn, dgamma = sum dn x_hat;
// per group: A = sum_c gamma dbeta / m,  B = sum_c gamma dgamma / m  (m = cpg * voxels);
//   dy = rstd (gamma dn - A - x_hat B).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGbChunk = 128;

__global__ void __launch_bounds__(256) gn_backward_partial_kernel(const float* __restrict__ dn, const float* __restrict__ y,
                                                                  const float* __restrict__ mean_rstd, long long spatial,
                                                                  int c, int cpg, int chunk_voxels,
                                                                  float* __restrict__ partial /*[n][c][chunks][2]*/,
                                                                  int chunks) {
    extern __shared__ float s_acc[];
    const int quads = c / 4;
    const int rows = blockDim.x / quads;
    const int q = threadIdx.x % quads, r = threadIdx.x / quads;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const long long v0 = 1ll * chunk * chunk_voxels;
    long long v1 = v0 + chunk_voxels;
    if (v1 > spatial) v1 = spatial;
    const int groups = c / cpg;
    float mu[4], rs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float* mr = mean_rstd + (static_cast<size_t>(n) * groups + (4 * q + k) / cpg) * 2;
        mu[k] = mr[0]; rs[k] = mr[1];
    }
    float sb[4] = {0.f, 0.f, 0.f, 0.f}, sg[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* dbase = reinterpret_cast<const float4*>(dn + (static_cast<size_t>(n) * spatial) * c) + q;
    const float4* ybase = reinterpret_cast<const float4*>(y + (static_cast<size_t>(n) * spatial) * c) + q;
    for (long long v = v0 + r; v < v1; v += rows) {
        const float4 d = __ldg(dbase + v * quads), yy = __ldg(ybase + v * quads);
        sb[0] += d.x; sg[0] += d.x * ((yy.x - mu[0]) * rs[0]);
        sb[1] += d.y; sg[1] += d.y * ((yy.y - mu[1]) * rs[1]);
        sb[2] += d.z; sg[2] += d.z * ((yy.z - mu[2]) * rs[2]);
        sb[3] += d.w; sg[3] += d.w * ((yy.w - mu[3]) * rs[3]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s_acc[(r * c + 4 * q + k) * 2 + 0] = sb[k];
        s_acc[(r * c + 4 * q + k) * 2 + 1] = sg[k];
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int rr = 0; rr < rows; ++rr) {
            a += s_acc[(rr * c + ch) * 2 + 0];
            b += s_acc[(rr * c + ch) * 2 + 1];
        }
        float* out = partial + ((static_cast<size_t>(n) * c + ch) * chunks + chunk) * 2;
        out[0] = a;
        out[1] = b;
    }
}

// one block per (group, sample): per-channel dbeta / dgamma of this sample and the two group terms
__global__ void __launch_bounds__(256) gn_backward_finalize_kernel(const float* __restrict__ partial, int chunks, int c,
                                                                   int cpg, long long spatial,
                                                                   const float* __restrict__ gamma,
                                                                   float* __restrict__ dgamma_dbeta /*[n][c][2]*/,
                                                                   float* __restrict__ group_terms /*[n][groups][2]*/) {
    const int g = blockIdx.x, n = blockIdx.y, groups = c / cpg;
    __shared__ double sh[2][256];
    __shared__ double s_ch[2][512];
    for (int ch = 0; ch < cpg; ++ch) {
        const float2* base = reinterpret_cast<const float2*>(partial) +
                             (static_cast<size_t>(n) * c + g * cpg + ch) * chunks;
        double a = 0.0, b = 0.0;
        for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
            const float2 v = __ldg(base + i);
            a += v.x;
            b += v.y;
        }
        sh[0][threadIdx.x] = a;
        sh[1][threadIdx.x] = b;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
                sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            s_ch[0][ch] = sh[0][0];
            s_ch[1][ch] = sh[1][0];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double A = 0.0, B = 0.0;
        for (int ch = 0; ch < cpg; ++ch) {
            const int cc = g * cpg + ch;
            float* o = dgamma_dbeta + (static_cast<size_t>(n) * c + cc) * 2;
            o[0] = static_cast<float>(s_ch[1][ch]);      // dgamma
            o[1] = static_cast<float>(s_ch[0][ch]);      // dbeta
            A += static_cast<double>(gamma[cc]) * s_ch[0][ch];
            B += static_cast<double>(gamma[cc]) * s_ch[1][ch];
        }
        const double m = static_cast<double>(spatial) * cpg;
        float* gt = group_terms + (static_cast<size_t>(n) * groups + g) * 2;
        gt[0] = static_cast<float>(A / m);
        gt[1] = static_cast<float>(B / m);
    }
}

// dy = rstd (gamma dn - A - x_hat B), written in place over dn (fp32)
__global__ void __launch_bounds__(256) gn_backward_apply_kernel(float* __restrict__ dn, const float* __restrict__ y,
                                                                const float* __restrict__ mean_rstd,
                                                                const float* __restrict__ group_terms,
                                                                const float* __restrict__ gamma, long long spatial, int c,
                                                                int cpg, long long total_quads) {
    const int quads = c / 4, groups = c / cpg;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total_quads; i += 256ll * gridDim.x) {
        const int q = static_cast<int>(i % quads);
        const int nn = static_cast<int>(i / (spatial * quads));
        float4 d = reinterpret_cast<float4*>(dn)[i];
        const float4 yy = __ldg(reinterpret_cast<const float4*>(y) + i);
        float dv[4] = {d.x, d.y, d.z, d.w};
        const float yv[4] = {yy.x, yy.y, yy.z, yy.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ch = 4 * q + k, g = ch / cpg;
            const float* mr = mean_rstd + (static_cast<size_t>(nn) * groups + g) * 2;
            const float* gt = group_terms + (static_cast<size_t>(nn) * groups + g) * 2;
            const float xh = (yv[k] - mr[0]) * mr[1];
            dv[k] = mr[1] * (gamma[ch] * dv[k] - gt[0] - xh * gt[1]);
        }
        reinterpret_cast<float4*>(dn)[i] = make_float4(dv[0], dv[1], dv[2], dv[3]);
    }
}

// Fused tail of the GroupNorm backward: dy = rstd (gamma dn - A - x_hat B) is written ONLY as bf16 planes (hi / lo) --
// the form both consumers want (dgrad convolution and direct wgrad) -- and summed per channel on the way (the conv
// bias gradient).  dy is never materialised in fp32: one read of dn and y, one 4-byte-per-element write, instead of
// apply (read 8 B, write 4 B) + channel sum (read 4 B) + plane conversion (read 4 B, write 4 B).
// Block = one chunk of kApChunk voxels; bias partials [chunks][c] are reduced in a fixed order afterwards.
constexpr int kApChunk = 64;

__global__ void __launch_bounds__(256) gn_backward_apply_planes_kernel(
    const float* __restrict__ dn, const float* __restrict__ y, const float* __restrict__ mean_rstd,
    const float* __restrict__ group_terms, const float* __restrict__ gamma, long long spatial, long long rows_total, int c,
    int cpg, __nv_bfloat16* __restrict__ dst, size_t plane_elems, int planes, float* __restrict__ bias_partial) {
    extern __shared__ float s_acc[];
    const int quads = c / 4, groups = c / cpg;
    const int rows = blockDim.x / quads;
    const int q = threadIdx.x % quads, r = threadIdx.x / quads;
    const long long v0 = 1ll * blockIdx.x * kApChunk;
    long long v1 = v0 + kApChunk;
    if (v1 > rows_total) v1 = rows_total;
    // per-channel constants of dy = rstd (gamma dn - A - x_hat B) = a1 dn + a2 y + a3 are hoisted out of the voxel loop
    // whenever the chunk lies inside one sample (always for batch 1)
    const int n_first = static_cast<int>(v0 / spatial);
    const bool one_sample = (v1 - 1) / spatial == n_first;
    float g4[4], mu[4], rs[4], ga[4], gb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ch = 4 * q + k, g = ch / cpg;
        g4[k] = gamma[ch];
        const float* mr = mean_rstd + (static_cast<size_t>(n_first) * groups + g) * 2;
        const float* gt = group_terms + (static_cast<size_t>(n_first) * groups + g) * 2;
        mu[k] = mr[0]; rs[k] = mr[1]; ga[k] = gt[0]; gb[k] = gt[1];
    }
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long v = v0 + r; v < v1; v += rows) {
        if (!one_sample) {
            const int nn = static_cast<int>(v / spatial);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int g = (4 * q + k) / cpg;
                const float* mr = mean_rstd + (static_cast<size_t>(nn) * groups + g) * 2;
                const float* gt = group_terms + (static_cast<size_t>(nn) * groups + g) * 2;
                mu[k] = mr[0]; rs[k] = mr[1]; ga[k] = gt[0]; gb[k] = gt[1];
            }
        }
        const long long i = v * quads + q;
        const float4 d = __ldg(reinterpret_cast<const float4*>(dn) + i);
        const float4 yy = __ldg(reinterpret_cast<const float4*>(y) + i);
        const float dv[4] = {d.x, d.y, d.z, d.w};
        const float yv[4] = {yy.x, yy.y, yy.z, yy.w};
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float xh = (yv[k] - mu[k]) * rs[k];
            const float o = rs[k] * (g4[k] * dv[k] - ga[k] - xh * gb[k]);       // same expression as gn_backward_apply
            s[k] += o;
            split_bf16_b(o, hi[k], lo[k]);
        }
        *reinterpret_cast<uint2*>(dst + i * 4) = *reinterpret_cast<uint2*>(hi);
        if (planes == 2) *reinterpret_cast<uint2*>(dst + plane_elems + i * 4) = *reinterpret_cast<uint2*>(lo);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) s_acc[r * c + 4 * q + k] = s[k];
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float a = 0.f;
        for (int rr = 0; rr < rows; ++rr) a += s_acc[rr * c + ch];
        bias_partial[static_cast<size_t>(blockIdx.x) * c + ch] = a;
    }
}

// per-channel sums of an NDHWC fp32 tensor (bias gradients): partial [chunks][c], then reduce_rows
__global__ void __launch_bounds__(256) channel_sum_partial_kernel(const float* __restrict__ x, long long rows_total, int c,
                                                                  int chunk_rows, float* __restrict__ partial) {
    extern __shared__ float s_acc[];
    const int quads = c / 4;
    const int rows = blockDim.x / quads;
    const int q = threadIdx.x % quads, r = threadIdx.x / quads;
    const long long v0 = 1ll * blockIdx.x * chunk_rows;
    long long v1 = v0 + chunk_rows;
    if (v1 > rows_total) v1 = rows_total;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long v = v0 + r; v < v1; v += rows) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + v * quads + q);
        s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) s_acc[r * c + 4 * q + k] = s[k];
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float a = 0.f;
        for (int rr = 0; rr < rows; ++rr) a += s_acc[rr * c + ch];
        partial[static_cast<size_t>(blockIdx.x) * c + ch] = a;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 NDHWC -> bf16 planes (no activation), and zero-padded TRANSPOSED planes for the wgrad GEMMs:
//   dst[p][ch][ ((t+pad)*(h+2pad) + (y+pad))*(w+2pad) + (x+pad) ]  with row length k_pad (zeros elsewhere)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) to_planes_kernel(const float* __restrict__ x, long long total_quads,
                                                        __nv_bfloat16* __restrict__ dst, size_t plane_elems, int planes) {
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total_quads; i += 256ll * gridDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
        const float v[4] = {a.x, a.y, a.z, a.w};
        __nv_bfloat16 hi[4], lo[4];
        if (planes == STEMSEG_PLANES_FP16) {                 // one fp16 plane (same 2-byte storage)
            __half hh[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) hh[k] = __float2half_rn(v[k]);
            *reinterpret_cast<uint2*>(dst + i * 4) = *reinterpret_cast<uint2*>(hh);
            continue;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) split_bf16_b(v[k], hi[k], lo[k]);
        *reinterpret_cast<uint2*>(dst + i * 4) = *reinterpret_cast<uint2*>(hi);
        if (planes == 2) *reinterpret_cast<uint2*>(dst + plane_elems + i * 4) = *reinterpret_cast<uint2*>(lo);
    }
}

constexpr int kTrC = 32, kTrV = 32;
// pitch of one padded row of the transposed planes: a multiple of 8 bf16 so that every (dt, dh) tap offset is a
// multiple of 16 bytes -- a tiled TMA load faults ("illegal instruction") when the start coordinate of the contiguous
// dimension is not 16-byte aligned (measured on B200, scripts/tma_probe.cu)
__host__ __device__ __forceinline__ int padded_row_pitch(int w, int pad) { return pad ? (w + 2 * pad + 7) / 8 * 8 : w; }
// SRC_BF16 = false: src fp32 [n=1][t][h][w][c];  true: src bf16 planes [P][t*h*w][c]
template <bool SRC_BF16>
__global__ void __launch_bounds__(256) transpose_pad_kernel(const void* __restrict__ src, int t, int h, int w, int c,
                                                            int pad, int shifts, long long k_pad,
                                                            size_t src_plane_elems, __nv_bfloat16* __restrict__ dst,
                                                            int planes) {
    __shared__ float tile_hi[kTrV][kTrC + 1];
    __shared__ float tile_lo[kTrV][kTrC + 1];
    const long long spatial = 1ll * t * h * w;
    const long long v0 = 1ll * blockIdx.x * kTrV;
    const int c0 = blockIdx.y * kTrC;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
#pragma unroll
    for (int i = 0; i < kTrV / 8; ++i) {
        const long long v = v0 + ty + 8 * i;
        const int cc = c0 + tx;
        float hi = 0.f, lo = 0.f;
        if (v < spatial && cc < c) {
            if (SRC_BF16) {
                const __nv_bfloat16* s = static_cast<const __nv_bfloat16*>(src);
                hi = __bfloat162float(s[v * c + cc]);
                if (planes == 2) lo = __bfloat162float(s[src_plane_elems + v * c + cc]);
            } else {
                const float xv = static_cast<const float*>(src)[v * c + cc];
                __nv_bfloat16 bh, bl;
                split_bf16_b(xv, bh, bl);
                hi = __bfloat162float(bh);
                lo = __bfloat162float(bl);
            }
        }
        tile_hi[ty + 8 * i][tx] = hi;
        tile_lo[ty + 8 * i][tx] = lo;
    }
    __syncthreads();
    const int hp = h + 2 * pad, wp = padded_row_pitch(w, pad);
    const size_t plane_stride = static_cast<size_t>(shifts) * c * k_pad;
#pragma unroll
    for (int i = 0; i < kTrC / 8; ++i) {
        const int cc = c0 + ty + 8 * i;
        const long long v = v0 + tx;
        if (cc < c && v < spatial) {
            const int xw = static_cast<int>(v % w);
            const int yh = static_cast<int>((v / w) % h);
            const int tt = static_cast<int>(v / (1ll * w * h));
            const long long p = (1ll * (tt + pad) * hp + (yh + pad)) * wp + (xw + pad);
            const __nv_bfloat16 bh = __float2bfloat16_rn(tile_hi[tx][ty + 8 * i]);
            const __nv_bfloat16 bl = __float2bfloat16_rn(tile_lo[tx][ty + 8 * i]);
            for (int sft = 0; sft < shifts; ++sft) {
                // copy sft of a 3-shift set holds x shifted by dw = sft - 1 along the row: copy[q] = x[q + dw]
                const long long q = shifts == 3 ? p - (sft - 1) : p;
                const size_t o = (static_cast<size_t>(sft) * c + cc) * k_pad + q;
                dst[o] = bh;
                if (planes == 2) dst[plane_stride + o] = bl;
            }
        }
    }
}

// dW partial slices [S][ntaps][cout][cin] (fp32) -> torch layout dst[cout][cin_total][ntaps] at channel offset cin_begin.
// One block per (cout, 32 input channels): reads are coalesced along cin, the (cin, tap) tile is transposed through
// shared memory and written as one contiguous run of 32*ntaps floats.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ slices, int s_count, int cout, int ntaps,
                                                           int cin, float* __restrict__ dst, int cin_total, int cin_begin,
                                                           int accumulate) {
    __shared__ float s_tile[32 * 27];
    const int tiles_ci = (cin + 31) / 32;
    const int co = blockIdx.x / tiles_ci;
    const int ci0 = (blockIdx.x % tiles_ci) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int ci = ci0 + tx;
    for (int tap = ty; tap < ntaps; tap += 8) {
        double a = 0.0;
        if (ci < cin)
            for (int s = 0; s < s_count; ++s)
                a += slices[((static_cast<size_t>(s) * ntaps + tap) * cout + co) * cin + ci];
        s_tile[tx * ntaps + tap] = static_cast<float>(a);
    }
    __syncthreads();
    const int n_ci = cin - ci0 < 32 ? cin - ci0 : 32;
    float* o = dst + (static_cast<size_t>(co) * cin_total + cin_begin + ci0) * ntaps;
    for (int i = threadIdx.x; i < n_ci * ntaps; i += 256) o[i] = accumulate ? o[i] + s_tile[i] : s_tile[i];
}

}  // namespace
}  // namespace stemseg

using namespace stemseg;

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" size_t stemseg_head_backward_workspace_bytes(int32_t c) {
    const size_t blocks = static_cast<size_t>(device_sm_count()) * 4;
    return align_up(blocks * kHbJ * (c + 1) * sizeof(float), 256);
}

extern "C" int32_t stemseg_head_backward(const float* z, const float* y_low, int32_t n, int32_t t, int32_t h, int32_t w,
                                         int32_t c, int32_t t_scale, const float* out_weight, const float* out_bias,
                                         const int32_t* activation, int32_t n_out, const float* grad_out, float* dx,
                                         float* d_weight, float* d_bias, void* workspace, size_t workspace_bytes,
                                         void* stream_) {
    SS_REQUIRE(z && y_low && out_weight && activation && grad_out && dx && d_weight && d_bias && workspace,
               "head_backward: null pointer");
    SS_REQUIRE(t_scale == 1 || t_scale == 2, "head_backward: temporal scale must be 1 or 2");
    SS_REQUIRE(n >= 1 && t >= 1 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && t % t_scale == 0 && c >= 4 &&
                   c % 4 == 0 && c <= 256,
               "head_backward: bad shape");
    SS_REQUIRE(n_out >= 1 && n_out <= 64, "head_backward: n_out out of range");
    SS_REQUIRE(al16(z) && al16(y_low) && al16(dx) && al16(out_weight), "head_backward: pointers must be 16-byte aligned");
    const size_t need = stemseg_head_backward_workspace_bytes(c);
    if (workspace_bytes < need) {
        set_error("head_backward: workspace %zu < %zu bytes", workspace_bytes, need);
        return STEMSEG_ERR_WORKSPACE;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int blocks = device_sm_count() * 4;
    const size_t smem = (static_cast<size_t>(kHbJ) * c + static_cast<size_t>(kHbWarps) * kHbJ * (c + 1)) * sizeof(float);
    SS_CUDA_OK(cudaFuncSetAttribute(head_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    SS_REQUIRE(smem <= 100 * 1024, "head_backward: shared memory");
    float* partial = static_cast<float*>(workspace);
    for (int j0 = 0; j0 < n_out; j0 += kHbJ) {
        const int jn = n_out - j0 < kHbJ ? n_out - j0 : kHbJ;
        head_backward_kernel<<<blocks, kHbWarps * 32, smem, stream>>>(z, y_low, n, t, h, w, c, t_scale, t / t_scale, h / 2,
                                                                       w / 2, out_weight, out_bias, activation, grad_out,
                                                                       n_out, j0, dx, partial);
        // partial rows are [kHbJ][c + 1]: weights then the bias column
        for (int j = 0; j < jn; ++j) {
            const long long rs = static_cast<long long>(kHbJ) * (c + 1);
            reduce_rows_kernel<<<1, 256, 0, stream>>>(partial + static_cast<size_t>(j) * (c + 1), blocks, c, rs,
                                                      d_weight + static_cast<size_t>(j0 + j) * c, 0);
            reduce_rows_kernel<<<1, 256, 0, stream>>>(partial + static_cast<size_t>(j) * (c + 1) + c, blocks, 1, rs,
                                                      d_bias + j0 + j, 0);
        }
    }
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_upsample_transpose(const float* d_high, int32_t n, int32_t t, int32_t h, int32_t w, int32_t c,
                                              int32_t t_scale, float* d_low, void* stream_) {
    SS_REQUIRE(d_high && d_low, "upsample_transpose: null pointer");
    SS_REQUIRE(t_scale == 1 || t_scale == 2, "upsample_transpose: temporal scale must be 1 or 2");
    SS_REQUIRE(n >= 1 && t >= 1 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && t % t_scale == 0 && c >= 4 && c % 4 == 0,
               "upsample_transpose: bad shape");
    SS_REQUIRE(al16(d_high) && al16(d_low), "upsample_transpose: pointers must be 16-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long total = 1ll * n * (t / t_scale) * (h / 2) * (w / 2) * (c / 4);
    upsample_transpose_kernel<<<grid_cap(total, 256), 256, 0, stream>>>(d_high, n, t, h, w, c, t_scale, t / t_scale, h / 2,
                                                                        w / 2, d_low);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_pool_relu_backward(const float* d_out, const float* y, const float* scale_shift, int32_t n,
                                              int32_t t, int32_t h, int32_t w, int32_t c, int32_t pool, float* d_norm,
                                              void* stream_) {
    SS_REQUIRE(d_out && y && d_norm, "pool_relu_backward: null pointer");
    SS_REQUIRE(n >= 1 && t >= 1 && h >= 1 && w >= 1 && c >= 4 && c % 4 == 0, "pool_relu_backward: bad shape");
    SS_REQUIRE(al16(d_out) && al16(y) && al16(d_norm) && al16(scale_shift), "pool_relu_backward: alignment");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int t_out = pool ? (t - 1) / 2 + 1 : t;
    const long long total = 1ll * n * t * h * w * (c / 4);
    pool_relu_backward_kernel<<<grid_cap(total, 256), 256, 0, stream>>>(d_out, y, scale_shift, n, t, h, w, c, t_out,
                                                                        pool ? 1 : 0, d_norm);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" size_t stemseg_group_norm_backward_workspace_bytes(int32_t n, int64_t spatial, int32_t c) {
    const long long chunks = (spatial + kGbChunk - 1) / kGbChunk;
    return align_up(static_cast<size_t>(n) * chunks * c * 2 * sizeof(float), 256);
}

extern "C" int32_t stemseg_group_norm_backward(float* d_norm_to_dy, const float* y, const float* mean_rstd,
                                               const float* gamma, int32_t n, int64_t spatial, int32_t c,
                                               int32_t channels_per_group, float* dgamma_dbeta, float* group_terms,
                                               void* workspace, size_t workspace_bytes, void* stream_) {
    SS_REQUIRE(d_norm_to_dy && y && mean_rstd && gamma && dgamma_dbeta && group_terms && workspace,
               "group_norm_backward: null pointer");
    SS_REQUIRE(n >= 1 && spatial >= 1 && c >= 4 && c % 4 == 0 && c <= 1024, "group_norm_backward: bad shape");
    SS_REQUIRE(channels_per_group >= 1 && channels_per_group <= 512 && c % channels_per_group == 0,
               "group_norm_backward: bad group size");
    SS_REQUIRE(al16(d_norm_to_dy) && al16(y), "group_norm_backward: alignment");
    const size_t need = stemseg_group_norm_backward_workspace_bytes(n, spatial, c);
    if (workspace_bytes < need) {
        set_error("group_norm_backward: workspace %zu < %zu bytes", workspace_bytes, need);
        return STEMSEG_ERR_WORKSPACE;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int chunks = static_cast<int>((spatial + kGbChunk - 1) / kGbChunk);
    const int quads = c / 4;
    const int rows = 256 / quads >= 1 ? 256 / quads : 1;
    const int threads = quads * rows;
    const size_t smem = static_cast<size_t>(rows) * c * 2 * sizeof(float);
    SS_REQUIRE(threads <= 1024 && smem <= 48 * 1024, "group_norm_backward: channel count %d unsupported", c);
    gn_backward_partial_kernel<<<dim3(chunks, n), threads, smem, stream>>>(
        d_norm_to_dy, y, mean_rstd, spatial, c, channels_per_group, kGbChunk, static_cast<float*>(workspace), chunks);
    gn_backward_finalize_kernel<<<dim3(c / channels_per_group, n), 256, 0, stream>>>(
        static_cast<const float*>(workspace), chunks, c, channels_per_group, spatial, gamma, dgamma_dbeta, group_terms);
    const long long total_quads = 1ll * n * spatial * quads;
    gn_backward_apply_kernel<<<grid_cap(total_quads, 256), 256, 0, stream>>>(d_norm_to_dy, y, mean_rstd, group_terms, gamma,
                                                                            spatial, c, channels_per_group, total_quads);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" size_t stemseg_group_norm_backward_planes_workspace_bytes(int32_t n, int64_t spatial, int32_t c) {
    const size_t a = stemseg_group_norm_backward_workspace_bytes(n, spatial, c);
    const long long chunks = (static_cast<long long>(n) * spatial + kApChunk - 1) / kApChunk;
    return a + align_up(static_cast<size_t>(chunks) * c * sizeof(float), 256);
}

extern "C" int32_t stemseg_group_norm_backward_planes(const float* d_norm, const float* y, const float* mean_rstd,
                                                      const float* gamma, int32_t n, int64_t spatial, int32_t c,
                                                      int32_t channels_per_group, float* dgamma_dbeta, float* group_terms,
                                                      void* dy_planes, int32_t planes, float* d_bias, void* workspace,
                                                      size_t workspace_bytes, void* stream_) {
    SS_REQUIRE(d_norm && y && mean_rstd && gamma && dgamma_dbeta && group_terms && dy_planes && d_bias && workspace,
               "group_norm_backward_planes: null pointer");
    SS_REQUIRE(n >= 1 && spatial >= 1 && c >= 4 && c % 4 == 0 && c <= 1024, "group_norm_backward_planes: bad shape");
    SS_REQUIRE(channels_per_group >= 1 && channels_per_group <= 512 && c % channels_per_group == 0,
               "group_norm_backward_planes: bad group size");
    SS_REQUIRE(planes == 1 || planes == 2, "group_norm_backward_planes: planes must be 1 or 2");
    SS_REQUIRE(al16(d_norm) && al16(y) && al16(dy_planes), "group_norm_backward_planes: alignment");
    const size_t need = stemseg_group_norm_backward_planes_workspace_bytes(n, spatial, c);
    if (workspace_bytes < need) {
        set_error("group_norm_backward_planes: workspace %zu < %zu bytes", workspace_bytes, need);
        return STEMSEG_ERR_WORKSPACE;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int chunks = static_cast<int>((spatial + kGbChunk - 1) / kGbChunk);
    const int quads = c / 4;
    int rows = 256 / quads;
    if (rows < 1) rows = 1;
    const int threads = quads * rows;
    const size_t smem = static_cast<size_t>(rows) * c * 2 * sizeof(float);
    SS_REQUIRE(threads <= 1024 && smem <= 48 * 1024, "group_norm_backward_planes: channel count %d unsupported", c);
    float* gn_ws = static_cast<float*>(workspace);
    float* bias_ws = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) +
                                              stemseg_group_norm_backward_workspace_bytes(n, spatial, c));
    gn_backward_partial_kernel<<<dim3(chunks, n), threads, smem, stream>>>(d_norm, y, mean_rstd, spatial, c,
                                                                           channels_per_group, kGbChunk, gn_ws, chunks);
    gn_backward_finalize_kernel<<<dim3(c / channels_per_group, n), 256, 0, stream>>>(
        gn_ws, chunks, c, channels_per_group, spatial, gamma, dgamma_dbeta, group_terms);
    const long long rows_total = static_cast<long long>(n) * spatial;
    const int ap_chunks = static_cast<int>((rows_total + kApChunk - 1) / kApChunk);
    gn_backward_apply_planes_kernel<<<ap_chunks, threads, static_cast<size_t>(rows) * c * sizeof(float), stream>>>(
        d_norm, y, mean_rstd, group_terms, gamma, spatial, rows_total, c, channels_per_group,
        static_cast<__nv_bfloat16*>(dy_planes), static_cast<size_t>(rows_total) * c, planes, bias_ws);
    reduce_rows_tiled_kernel<<<(c + 31) / 32, 256, 0, stream>>>(bias_ws, ap_chunks, c, c, d_bias);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" size_t stemseg_channel_sum_workspace_bytes(int64_t rows, int32_t c) {
    const long long chunks = (rows + 255) / 256;
    return align_up(static_cast<size_t>(chunks) * c * sizeof(float), 256);
}

extern "C" int32_t stemseg_channel_sum(const float* x, int64_t rows, int32_t c, float* out, void* workspace,
                                       size_t workspace_bytes, void* stream_) {
    SS_REQUIRE(x && out && workspace && rows >= 1 && c >= 4 && c % 4 == 0 && c <= 1024, "channel_sum: bad arguments");
    const size_t need = stemseg_channel_sum_workspace_bytes(rows, c);
    if (workspace_bytes < need) {
        set_error("channel_sum: workspace %zu < %zu bytes", workspace_bytes, need);
        return STEMSEG_ERR_WORKSPACE;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int chunks = static_cast<int>((rows + 255) / 256);
    const int quads = c / 4;
    const int rws = 256 / quads >= 1 ? 256 / quads : 1;
    const size_t smem = static_cast<size_t>(rws) * c * sizeof(float);
    channel_sum_partial_kernel<<<chunks, quads * rws, smem, stream>>>(x, rows, c, 256, static_cast<float*>(workspace));
    reduce_rows_tiled_kernel<<<(c + 31) / 32, 256, 0, stream>>>(static_cast<const float*>(workspace), chunks, c, c, out);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_to_planes(const float* x, int64_t elems, void* dst_planes, int32_t planes, void* stream_) {
    SS_REQUIRE(x && dst_planes && elems >= 4 && elems % 4 == 0, "to_planes: bad arguments");
    SS_REQUIRE(planes == 1 || planes == 2 || planes == STEMSEG_PLANES_FP16, "to_planes: planes must be 1, 2 or STEMSEG_PLANES_FP16");
    SS_REQUIRE(al16(x) && al16(dst_planes), "to_planes: alignment");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    to_planes_kernel<<<grid_cap(elems / 4, 256), 256, 0, stream>>>(x, elems / 4, static_cast<__nv_bfloat16*>(dst_planes),
                                                                   static_cast<size_t>(elems), planes);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int64_t stemseg_transposed_row_length(int32_t t, int32_t h, int32_t w, int32_t pad) {
    const long long k = 1ll * (t + 2 * pad) * (h + 2 * pad) * padded_row_pitch(w, pad);
    return (k + 63) / 64 * 64;          // multiple of the K chunk; the tail stays zero
}

extern "C" int32_t stemseg_transpose_pad(const void* src, int32_t src_is_planes, int32_t t, int32_t h, int32_t w,
                                         int32_t c, int32_t pad, int32_t shifts, void* dst_planes, int32_t planes,
                                         void* stream_) {
    SS_REQUIRE(src && dst_planes, "transpose_pad: null pointer");
    SS_REQUIRE(shifts == 1 || (shifts == 3 && pad == 1), "transpose_pad: shifts must be 1, or 3 with pad 1");
    SS_REQUIRE(t >= 1 && h >= 1 && w >= 1 && c >= 1 && (pad == 0 || pad == 1), "transpose_pad: bad shape");
    SS_REQUIRE(planes == 1 || planes == 2, "transpose_pad: planes must be 1 or 2");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long k_pad = stemseg_transposed_row_length(t, h, w, pad);
    SS_CUDA_OK(cudaMemsetAsync(dst_planes, 0, static_cast<size_t>(planes) * shifts * c * k_pad * 2, stream));
    const long long spatial = 1ll * t * h * w;
    dim3 grid(static_cast<unsigned>((spatial + kTrV - 1) / kTrV), (c + kTrC - 1) / kTrC);
    if (src_is_planes)
        transpose_pad_kernel<true><<<grid, 256, 0, stream>>>(src, t, h, w, c, pad, shifts, k_pad,
                                                             static_cast<size_t>(spatial) * c,
                                                             static_cast<__nv_bfloat16*>(dst_planes), planes);
    else
        transpose_pad_kernel<false><<<grid, 256, 0, stream>>>(src, t, h, w, c, pad, shifts, k_pad, 0,
                                                              static_cast<__nv_bfloat16*>(dst_planes), planes);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_wgrad_reduce(const float* slices, int32_t n_slices, int32_t cout, int32_t ntaps, int32_t cin,
                                        float* dst, int32_t cin_total, int32_t cin_begin, int32_t accumulate,
                                        void* stream_) {
    SS_REQUIRE(slices && dst && n_slices >= 1 && cout >= 1 && ntaps >= 1 && cin >= 1 && cin_begin >= 0 &&
                   cin_begin + cin <= cin_total,
               "wgrad_reduce: bad arguments");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SS_REQUIRE(ntaps <= 27, "wgrad_reduce: at most 27 taps");
    const long long blocks = 1ll * cout * ((cin + 31) / 32);
    SS_REQUIRE(blocks < 0x7FFFFFFFll, "wgrad_reduce: too many blocks");
    wgrad_reduce_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(slices, n_slices, cout, ntaps, cin, dst, cin_total,
                                                                           cin_begin, accumulate);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}
