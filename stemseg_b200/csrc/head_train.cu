// Output heads of the TRAINING path (sm_100a): forward and backward over the merged feature x = W_b f + up(W_a x_8)
// kept in fp32 (the backward needs it), HBM-bound streaming kernels.
//
// Replaces, for training, the 1x1x1 output convs + activations + coordinate offsets of
//   embedding_decoder.py:90-96,131-145, seediness_decoder.py:80,112, semseg_decoder.py:86-87,116
// and torch autograd through them.  (Inference fuses the same arithmetic into the conv_4 GEMM epilogue, conv_tc.cu.)
//
//   stemseg_upsample_add_f32 : z <- z + up(y_low) in place (trilinear, align_corners=False), once per step
//   stemseg_head_output_x    : out[n][j][v] = act_j(W_j . x[v] + b_j) + coord_j
//   stemseg_head_backward_x  : dpre_j = g_j act_j'(pre_j); dx = sum_j dpre_j W_j; dW_j = sum_v dpre_j x[v]; db_j = sum_v dpre_j
//
// Work decomposition of both head kernels: a warp handles FOUR consecutive voxels per iteration; lane l owns the
// channel quads l, l+32.  The 4 x 8 = 32 partial dot products (4 voxels x up to 8 outputs per launch) are reduced
// with ONE 31-shuffle butterfly (warp_column_sums), after which lane (voxel*8 + j) owns pre[voxel][j] and applies the
// activation / its derivative.  Four independent 512-byte row loads per lane are in flight per iteration.
// dW / db partials: registers -> shared (fixed warp order) -> [blocks][8][c+1] -> reduce_cols (fixed order):
// deterministic.
#include "common.cuh"
#include "trilinear.cuh"

namespace stemseg {
namespace {

constexpr int kJ = 8;             // outputs per launch
constexpr int kVox = 4;           // voxels per warp iteration
constexpr int kWarps = 8;
constexpr int kMaxC = 256;

__device__ __forceinline__ void column_sums32(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
}

struct HeadGeom {
    int n, t, h, w, c;
    float x_abs, y_abs, t_abs;
};

__global__ void __launch_bounds__(256) upsample_add_f32_kernel(float* __restrict__ z, const float* __restrict__ ylow, int n,
                                                               int t, int h, int w, int c, int st, int tl, int hl, int wl) {
    const int quads = c / 4;
    const long long total = 1ll * n * t * h * w * quads;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
        const int q = static_cast<int>(i % quads);
        long long r = i / quads;
        const int wo = static_cast<int>(r % w); r /= w;
        const int ho = static_cast<int>(r % h); r /= h;
        const int to = static_cast<int>(r % t);
        const int nn = static_cast<int>(r / t);
        const Tri tr = make_tri(nn, to, ho, wo, st, tl, hl, wl, c);
        float4 a = reinterpret_cast<float4*>(z)[i];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (tr.wgt[k] != 0.f) {
                const float4 y = __ldg(reinterpret_cast<const float4*>(ylow + tr.off[k]) + q);
                a.x = fmaf(tr.wgt[k], y.x, a.x);
                a.y = fmaf(tr.wgt[k], y.y, a.y);
                a.z = fmaf(tr.wgt[k], y.z, a.z);
                a.w = fmaf(tr.wgt[k], y.w, a.w);
            }
        }
        reinterpret_cast<float4*>(z)[i] = a;
    }
}

// loads the rows of 4 consecutive voxels and forms the 32 partial dot products part[vox*8 + j]
template <int QQ>
__device__ __forceinline__ void load_and_dot(const float* __restrict__ x, long long vbase, long long total, int c, int quads,
                                             int lane, const float* s_w, int jn, float (&xq)[kVox][QQ][4],
                                             float (&part)[32]) {
#pragma unroll
    for (int vv = 0; vv < kVox; ++vv) {
#pragma unroll
        for (int qq = 0; qq < QQ; ++qq) {
            const int q = lane + 32 * qq;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < quads && vbase + vv < total)
                a = __ldg(reinterpret_cast<const float4*>(x + static_cast<size_t>(vbase + vv) * c) + q);
            xq[vv][qq][0] = a.x; xq[vv][qq][1] = a.y; xq[vv][qq][2] = a.z; xq[vv][qq][3] = a.w;
        }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) part[i] = 0.f;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
        if (j < jn) {                                   // block-uniform
#pragma unroll
            for (int qq = 0; qq < QQ; ++qq) {
                const int q = lane + 32 * qq;
                if (q < quads) {
                    const float4 wr = *reinterpret_cast<const float4*>(s_w + j * c + 4 * q);
#pragma unroll
                    for (int vv = 0; vv < kVox; ++vv) {
                        float a = part[vv * kJ + j];
                        a = fmaf(xq[vv][qq][0], wr.x, a);
                        a = fmaf(xq[vv][qq][1], wr.y, a);
                        a = fmaf(xq[vv][qq][2], wr.z, a);
                        a = fmaf(xq[vv][qq][3], wr.w, a);
                        part[vv * kJ + j] = a;
                    }
                }
            }
        }
    }
}

template <int QQ>
__global__ void __launch_bounds__(kWarps * 32) head_x_forward_kernel(
    const float* __restrict__ x, HeadGeom gm, const float* __restrict__ wout, const float* __restrict__ bout,
    const int* __restrict__ act, const int* __restrict__ coord, int j_total, int j0, float* __restrict__ out) {
    extern __shared__ float s_mem[];
    float* s_w = s_mem;                               // [kJ][c]
    const int c = gm.c;
    const int jn = min(kJ, j_total - j0);
    for (int i = threadIdx.x; i < jn * c; i += blockDim.x) s_w[i] = wout[(j0 + i / c) * c + i % c];
    __syncthreads();
    const long long spatial = 1ll * gm.t * gm.h * gm.w;
    const long long total = 1ll * gm.n * spatial;
    const int quads = c / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long groups = (total + kVox - 1) / kVox;
    const long long warps = 1ll * gridDim.x * kWarps;
    const int my_vox = lane >> 3, my_j = lane & 7;
    for (long long g = 1ll * blockIdx.x * kWarps + warp; g < groups; g += warps) {
        const long long vbase = g * kVox;
        float xq[kVox][QQ][4];
        float part[32];
        load_and_dot<QQ>(x, vbase, total, c, quads, lane, s_w, jn, xq, part);
        column_sums32(part, lane);                      // lane (vox*8 + j) now holds the full dot product
        const long long v = vbase + my_vox;
        if (my_j < jn && v < total) {
            const int jj = j0 + my_j;
            long long r = v;
            const int wo = static_cast<int>(r % gm.w); r /= gm.w;
            const int ho = static_cast<int>(r % gm.h); r /= gm.h;
            const int to = static_cast<int>(r % gm.t);
            const int nn = static_cast<int>(r / gm.t);
            float val = part[0] + (bout ? bout[jj] : 0.f);
            const int a = act[jj];
            if (a == 1) val = tanhf(0.25f * val);
            else if (a == 2) val = 1.0f / (1.0f + expf(-val));
            const int cd = coord[jj];
            if (cd == 1) val += linspace_value(gm.t_abs, gm.t, to);
            else if (cd == 2) val += linspace_value(gm.y_abs, gm.h, ho);
            else if (cd == 3) val += linspace_value(gm.x_abs, gm.w, wo);
            out[(static_cast<size_t>(nn) * j_total + jj) * spatial + (v - nn * spatial)] = val;
        }
    }
}

template <int QQ>
__global__ void __launch_bounds__(kWarps * 32) head_x_backward_kernel(
    const float* __restrict__ x, HeadGeom gm, const float* __restrict__ wout, const float* __restrict__ bout,
    const int* __restrict__ act, const float* __restrict__ g, int j_total, int j0,
    float* __restrict__ dx /*[n][t][h][w][c], accumulated over the j0 passes*/,
    float* __restrict__ partial /*[gridDim.x][kJ][c + 1]*/) {
    extern __shared__ float s_mem[];
    const int c = gm.c;
    float* s_w = s_mem;                               // [kJ][c]
    float* s_acc = s_mem + kJ * c;                    // [kWarps][kJ][c + 1]
    const int jn = min(kJ, j_total - j0);
    for (int i = threadIdx.x; i < jn * c; i += blockDim.x) s_w[i] = wout[(j0 + i / c) * c + i % c];
    __syncthreads();
    const long long spatial = 1ll * gm.t * gm.h * gm.w;
    const long long total = 1ll * gm.n * spatial;
    const int quads = c / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long groups = (total + kVox - 1) / kVox;
    const long long warps = 1ll * gridDim.x * kWarps;
    const int my_vox = lane >> 3, my_j = lane & 7;
    float accw[kJ][QQ][4];
    float accb = 0.f;                                 // lane (vox, j): partial of db_j over this lane's voxel slot
#pragma unroll
    for (int j = 0; j < kJ; ++j)
#pragma unroll
        for (int qq = 0; qq < QQ; ++qq) accw[j][qq][0] = accw[j][qq][1] = accw[j][qq][2] = accw[j][qq][3] = 0.f;
    for (long long gi = 1ll * blockIdx.x * kWarps + warp; gi < groups; gi += warps) {
        const long long vbase = gi * kVox;
        float xq[kVox][QQ][4];
        float part[32];
        // the upstream gradient of this lane's (voxel, output) does not depend on the dot products: load it first
        const long long v = vbase + my_vox;
        float gg = 0.f;
        const bool mine = my_j < jn && v < total;
        if (mine) {
            const int nn = static_cast<int>(v / spatial);
            gg = __ldg(g + (static_cast<size_t>(nn) * j_total + j0 + my_j) * spatial + (v - nn * spatial));
        }
        load_and_dot<QQ>(x, vbase, total, c, quads, lane, s_w, jn, xq, part);
        column_sums32(part, lane);
        float d = 0.f;
        if (mine) {
            const int jj = j0 + my_j;
            const float pre = part[0] + (bout ? bout[jj] : 0.f);
            d = gg;
            const int ac = act[jj];
            if (ac == 1) {
                const float th = tanhf(0.25f * pre);
                d = gg * 0.25f * (1.f - th * th);
            } else if (ac == 2) {
                const float sg = 1.0f / (1.0f + expf(-pre));
                d = gg * sg * (1.f - sg);
            }
        }
        accb += d;
        // every lane needs dpre of all (voxel, output) pairs
#pragma unroll
        for (int i = 0; i < 32; ++i) part[i] = __shfl_sync(0xffffffffu, d, i);
#pragma unroll
        for (int qq = 0; qq < QQ; ++qq) {
            const int q = lane + 32 * qq;
            if (q < quads) {
                float o[kVox][4];
#pragma unroll
                for (int vv = 0; vv < kVox; ++vv) {
                    o[vv][0] = o[vv][1] = o[vv][2] = o[vv][3] = 0.f;
                    if (j0 != 0 && vbase + vv < total) {
                        const float4 prev = *(reinterpret_cast<const float4*>(dx + static_cast<size_t>(vbase + vv) * c) + q);
                        o[vv][0] = prev.x; o[vv][1] = prev.y; o[vv][2] = prev.z; o[vv][3] = prev.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < kJ; ++j) {
                    if (j < jn) {
                        const float4 wr = *reinterpret_cast<const float4*>(s_w + j * c + 4 * q);
#pragma unroll
                        for (int vv = 0; vv < kVox; ++vv) {
                            const float dp = part[vv * kJ + j];
                            o[vv][0] = fmaf(dp, wr.x, o[vv][0]);
                            o[vv][1] = fmaf(dp, wr.y, o[vv][1]);
                            o[vv][2] = fmaf(dp, wr.z, o[vv][2]);
                            o[vv][3] = fmaf(dp, wr.w, o[vv][3]);
                            accw[j][qq][0] = fmaf(dp, xq[vv][qq][0], accw[j][qq][0]);
                            accw[j][qq][1] = fmaf(dp, xq[vv][qq][1], accw[j][qq][1]);
                            accw[j][qq][2] = fmaf(dp, xq[vv][qq][2], accw[j][qq][2]);
                            accw[j][qq][3] = fmaf(dp, xq[vv][qq][3], accw[j][qq][3]);
                        }
                    }
                }
#pragma unroll
                for (int vv = 0; vv < kVox; ++vv)
                    if (vbase + vv < total)
                        *(reinterpret_cast<float4*>(dx + static_cast<size_t>(vbase + vv) * c) + q) =
                            make_float4(o[vv][0], o[vv][1], o[vv][2], o[vv][3]);
            }
        }
    }
    // db_j: sum the four voxel slots (lanes j, j+8, j+16, j+24), fixed order
    accb += __shfl_xor_sync(0xffffffffu, accb, 8);
    accb += __shfl_xor_sync(0xffffffffu, accb, 16);
    const int stride = c + 1;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
#pragma unroll
        for (int qq = 0; qq < QQ; ++qq) {
            const int q = lane + 32 * qq;
            if (q < quads)
#pragma unroll
                for (int e = 0; e < 4; ++e) s_acc[(warp * kJ + j) * stride + 4 * q + e] = accw[j][qq][e];
        }
    }
    if (lane < kJ) s_acc[(warp * kJ + lane) * stride + c] = accb;
    __syncthreads();
    for (int i = threadIdx.x; i < kJ * stride; i += blockDim.x) {
        float a = 0.f;
        for (int wv = 0; wv < kWarps; ++wv) a += s_acc[wv * kJ * stride + i];
        partial[static_cast<size_t>(blockIdx.x) * kJ * stride + i] = a;
    }
}

// d_weight[j0 + j][k] / d_bias[j0 + j] = sum_b partial[b][j][k] (k = c is the bias column); fixed order, double
__global__ void __launch_bounds__(256) head_reduce_cols_kernel(const float* __restrict__ partial, int blocks, int c, int jn,
                                                               int j0, float* __restrict__ d_weight,
                                                               float* __restrict__ d_bias) {
    __shared__ double s[8][33];
    const int stride = c + 1;
    const int cols = jn * stride;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    double a = 0.0;
    if (col < cols)
        for (int b = ty; b < blocks; b += 8) a += static_cast<double>(partial[static_cast<size_t>(b) * kJ * stride + col]);
    s[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && col < cols) {
        double tot = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) tot += s[r][tx];
        const int j = col / stride, k = col % stride;
        if (k < c) d_weight[static_cast<size_t>(j0 + j) * c + k] = static_cast<float>(tot);
        else d_bias[j0 + j] = static_cast<float>(tot);
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int head_blocks() { return device_sm_count() * 2; }

HeadGeom make_geom(int n, int t, int h, int w, int c, float time_scale) {
    HeadGeom gm;
    gm.n = n; gm.t = t; gm.h = h; gm.w = w; gm.c = c;
    // embedding_utils.py:31-37: x in [-max(1, W/H), +], y in [-max(1, H/W), +], t in [-time_scale, +]
    gm.x_abs = fmaxf(1.0f, static_cast<float>(static_cast<double>(w) / static_cast<double>(h)));
    gm.y_abs = fmaxf(1.0f, static_cast<float>(static_cast<double>(h) / static_cast<double>(w)));
    gm.t_abs = time_scale;
    return gm;
}

}  // namespace
}  // namespace stemseg

using namespace stemseg;

extern "C" int32_t stemseg_upsample_add_f32(float* z, const float* y_low, int32_t n, int32_t t, int32_t h, int32_t w,
                                            int32_t c, int32_t t_scale, void* stream_) {
    SS_REQUIRE(z && y_low, "upsample_add_f32: null pointer");
    SS_REQUIRE(t_scale == 1 || t_scale == 2, "upsample_add_f32: temporal scale must be 1 or 2");
    SS_REQUIRE(n >= 1 && t >= 1 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && t % t_scale == 0 && c >= 4 && c % 4 == 0,
               "upsample_add_f32: bad shape");
    SS_REQUIRE(al16(z) && al16(y_low), "upsample_add_f32: pointers must be 16-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long total = 1ll * n * t * h * w * (c / 4);
    long long blocks = (total + 255) / 256;
    const long long cap = 16ll * device_sm_count();
    if (blocks > cap) blocks = cap;
    upsample_add_f32_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(z, y_low, n, t, h, w, c, t_scale, t / t_scale,
                                                                              h / 2, w / 2);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_head_output_x(const float* x, int32_t n, int32_t t, int32_t h, int32_t w, int32_t c,
                                         const float* out_weight, const float* out_bias, const int32_t* activation,
                                         const int32_t* coordinate, int32_t n_out, float time_scale, float* out,
                                         void* stream_) {
    SS_REQUIRE(x && out_weight && activation && coordinate && out, "head_output_x: null pointer");
    SS_REQUIRE(n >= 1 && t >= 1 && h >= 1 && w >= 1 && c >= 4 && c % 4 == 0 && c <= kMaxC, "head_output_x: bad shape");
    SS_REQUIRE(n_out >= 1 && n_out <= 64, "head_output_x: n_out %d out of range [1,64]", n_out);
    SS_REQUIRE(al16(x) && al16(out_weight), "head_output_x: pointers must be 16-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const HeadGeom gm = make_geom(n, t, h, w, c, time_scale);
    const size_t smem = static_cast<size_t>(kJ) * c * sizeof(float);
    const int blocks = head_blocks();
    for (int j0 = 0; j0 < n_out; j0 += kJ) {
        if (c <= 128)
            head_x_forward_kernel<1><<<blocks, kWarps * 32, smem, stream>>>(x, gm, out_weight, out_bias, activation,
                                                                            coordinate, n_out, j0, out);
        else
            head_x_forward_kernel<2><<<blocks, kWarps * 32, smem, stream>>>(x, gm, out_weight, out_bias, activation,
                                                                            coordinate, n_out, j0, out);
    }
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" size_t stemseg_head_backward_x_workspace_bytes(int32_t c) {
    return align_up(static_cast<size_t>(head_blocks()) * kJ * (c + 1) * sizeof(float), 256);
}

extern "C" int32_t stemseg_head_backward_x(const float* x, int32_t n, int32_t t, int32_t h, int32_t w, int32_t c,
                                           const float* out_weight, const float* out_bias, const int32_t* activation,
                                           int32_t n_out, const float* grad_out, float* dx, float* d_weight,
                                           float* d_bias, void* workspace, size_t workspace_bytes, void* stream_) {
    SS_REQUIRE(x && out_weight && activation && grad_out && dx && d_weight && d_bias && workspace,
               "head_backward_x: null pointer");
    SS_REQUIRE(n >= 1 && t >= 1 && h >= 1 && w >= 1 && c >= 4 && c % 4 == 0 && c <= kMaxC, "head_backward_x: bad shape");
    SS_REQUIRE(n_out >= 1 && n_out <= 64, "head_backward_x: n_out out of range");
    SS_REQUIRE(al16(x) && al16(dx) && al16(out_weight), "head_backward_x: pointers must be 16-byte aligned");
    const size_t need = stemseg_head_backward_x_workspace_bytes(c);
    if (workspace_bytes < need) {
        set_error("head_backward_x: workspace %zu < %zu bytes", workspace_bytes, need);
        return STEMSEG_ERR_WORKSPACE;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const HeadGeom gm = make_geom(n, t, h, w, c, 1.0f);
    const int blocks = head_blocks();
    const size_t smem = (static_cast<size_t>(kJ) * c + static_cast<size_t>(kWarps) * kJ * (c + 1)) * sizeof(float);
    SS_REQUIRE(smem <= 100 * 1024, "head_backward_x: shared memory");
    SS_CUDA_OK(cudaFuncSetAttribute(head_x_backward_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    SS_CUDA_OK(cudaFuncSetAttribute(head_x_backward_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    float* partial = static_cast<float*>(workspace);
    for (int j0 = 0; j0 < n_out; j0 += kJ) {
        const int jn = n_out - j0 < kJ ? n_out - j0 : kJ;
        if (c <= 128)
            head_x_backward_kernel<1><<<blocks, kWarps * 32, smem, stream>>>(x, gm, out_weight, out_bias, activation, grad_out,
                                                                             n_out, j0, dx, partial);
        else
            head_x_backward_kernel<2><<<blocks, kWarps * 32, smem, stream>>>(x, gm, out_weight, out_bias, activation, grad_out,
                                                                             n_out, j0, dx, partial);
        const int cols = jn * (c + 1);
        head_reduce_cols_kernel<<<(cols + 31) / 32, 256, 0, stream>>>(partial, blocks, c, jn, j0, d_weight, d_bias);
    }
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}
