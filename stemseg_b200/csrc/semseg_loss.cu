// Losses of the semantic-segmentation head (YouTube-VIS / KITTI-MOTS configs), forward AND gradient, one sequence per
// call (sm_100a).  HBM-bound streaming over channels-first logits [C][M] (M = T*H*W voxels, the heads' API layout).
//
// Replaces (paths relative to the reference root)
//   CrossEntropyLoss.forward           stemseg/modeling/losses/cross_entropy.py:13-49   (F.cross_entropy over the classes)
//   TrainingModel.compute_fg_loss      stemseg/modeling/model_builder.py:210-244         (BCE-with-logits, fg = class id > 0)
// and torch autograd through them.
//   class term : ce = mean_v [ logsumexp_c x[c][v] - x[gt_v][v] ]; the reference multiplies this SCALAR by the
//                non-ignore mask and divides by the mask sum (cross_entropy.py:36-42), i.e. ce * (S / S) with S the number
//                of non-ignored voxels: ignored voxels still count in the class loss, and S = 0 gives NaN -- both kept.
//   fg term    : sum_v bce(x_fg[v], gt_v > 0) nonignore_v / S
// Two passes: (1) per-voxel log-sum-exp / BCE, block sums -> double atomics; (2) gradient
//   d x[c][v] = w_semseg (softmax_c - [c == gt_v]) (S / S) / M,   d x_fg[v] = w_fg (sigmoid(x_fg) - [gt_v > 0]) nonignore_v / S.
// Either term can be switched off with a null pointer, so the reference's two call sites map to two calls and the
// trainer's fused path to one.
#include "common.cuh"

namespace stemseg {
namespace {

struct SemsegAcc {
    double ce_sum, bce_sum, nonignore;
};

__device__ __forceinline__ double warp_sum_d(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

__global__ void __launch_bounds__(256) semseg_loss_reduce_kernel(const float* __restrict__ cls, long long cls_stride,
                                                                 int n_classes, const float* __restrict__ fg,
                                                                 const long long* __restrict__ ids,
                                                                 const uint8_t* __restrict__ ignore, long long m,
                                                                 SemsegAcc* __restrict__ acc) {
    __shared__ double s_red[8][3];
    double ce = 0.0, bce = 0.0, cnt = 0.0;
    const long long stride = 1ll * gridDim.x * blockDim.x;
    for (long long v = 1ll * blockIdx.x * blockDim.x + threadIdx.x; v < m; v += stride) {
        const long long gt = ids[v];
        const float keep = (ignore == nullptr || ignore[v] == 0) ? 1.f : 0.f;
        cnt += keep;
        if (cls != nullptr) {
            float mx = -INFINITY;
            for (int c = 0; c < n_classes; ++c) mx = fmaxf(mx, cls[c * cls_stride + v]);
            float se = 0.f;
            for (int c = 0; c < n_classes; ++c) se += expf(cls[c * cls_stride + v] - mx);
            const float xg = (gt >= 0 && gt < n_classes) ? cls[gt * cls_stride + v] : mx + logf(se);
            ce += static_cast<double>(logf(se) + mx - xg);
        }
        if (fg != nullptr) {
            const float x = fg[v], y = gt > 0 ? 1.f : 0.f;
            // binary_cross_entropy_with_logits: max(x, 0) - x y + log(1 + exp(-|x|))
            const float l = fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
            bce += static_cast<double>(l * keep);
        }
    }
    ce = warp_sum_d(ce); bce = warp_sum_d(bce); cnt = warp_sum_d(cnt);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[warp][0] = ce; s_red[warp][1] = bce; s_red[warp][2] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int w = 0; w < 8; ++w) { a += s_red[w][0]; b += s_red[w][1]; c += s_red[w][2]; }
        atomicAdd(&acc->ce_sum, a);
        atomicAdd(&acc->bce_sum, b);
        atomicAdd(&acc->nonignore, c);
    }
}

__global__ void __launch_bounds__(256) semseg_loss_grad_kernel(const float* __restrict__ cls, long long cls_stride,
                                                               int n_classes, const float* __restrict__ fg,
                                                               const long long* __restrict__ ids,
                                                               const uint8_t* __restrict__ ignore, long long m,
                                                               const SemsegAcc* __restrict__ acc, float w_semseg,
                                                               float w_fg, float* __restrict__ d_cls,
                                                               long long d_cls_stride, float* __restrict__ d_fg,
                                                               float* __restrict__ losses) {
    const double s = acc->nonignore;
    const float unit = static_cast<float>(s / s);                     // 1, or NaN when every voxel is ignored
    const float ce_scale = w_semseg * unit / static_cast<float>(m);
    const float fg_scale = w_fg / static_cast<float>(s);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        losses[0] = cls != nullptr ? static_cast<float>(acc->ce_sum / static_cast<double>(m) * (s / s)) : 0.f;
        losses[1] = fg != nullptr ? static_cast<float>(acc->bce_sum / s) : 0.f;
    }
    const long long stride = 1ll * gridDim.x * blockDim.x;
    for (long long v = 1ll * blockIdx.x * blockDim.x + threadIdx.x; v < m; v += stride) {
        const long long gt = ids[v];
        if (cls != nullptr) {
            float mx = -INFINITY;
            for (int c = 0; c < n_classes; ++c) mx = fmaxf(mx, cls[c * cls_stride + v]);
            float se = 0.f;
            for (int c = 0; c < n_classes; ++c) se += expf(cls[c * cls_stride + v] - mx);
            const float inv = 1.0f / se;
            for (int c = 0; c < n_classes; ++c) {
                const float p = expf(cls[c * cls_stride + v] - mx) * inv;
                d_cls[c * d_cls_stride + v] = (p - (c == gt ? 1.f : 0.f)) * ce_scale;
            }
        }
        if (fg != nullptr) {
            const float keep = (ignore == nullptr || ignore[v] == 0) ? 1.f : 0.f;
            const float x = fg[v], y = gt > 0 ? 1.f : 0.f;
            const float sg = 1.0f / (1.0f + expf(-x));
            d_fg[v] = (sg - y) * keep * fg_scale;
        }
    }
}

}  // namespace
}  // namespace stemseg

using namespace stemseg;

extern "C" size_t stemseg_semseg_loss_workspace_bytes(void) { return align_up(sizeof(SemsegAcc), 256); }

extern "C" int32_t stemseg_semseg_loss(const float* class_logits, int64_t class_stride, int32_t n_classes,
                                       const float* fg_logits, const int64_t* class_ids, const uint8_t* ignore,
                                       int64_t voxels, float w_semseg, float w_foreground, float* losses, float* d_class,
                                       int64_t d_class_stride, float* d_fg, void* workspace, size_t workspace_bytes,
                                       void* stream_) {
    SS_REQUIRE(class_ids && losses && workspace, "semseg_loss: null pointer");
    SS_REQUIRE(class_logits != nullptr || fg_logits != nullptr, "semseg_loss: neither class nor foreground logits given");
    SS_REQUIRE(class_logits == nullptr || (d_class != nullptr && n_classes >= 1 && n_classes <= 4096 &&
                                           class_stride >= voxels && d_class_stride >= voxels),
               "semseg_loss: bad class-logit arguments");
    SS_REQUIRE(fg_logits == nullptr || d_fg != nullptr, "semseg_loss: d_fg is null");
    SS_REQUIRE(voxels >= 1, "semseg_loss: voxels must be positive");
    if (workspace_bytes < sizeof(SemsegAcc)) {
        set_error("semseg_loss: workspace %zu < %zu bytes", workspace_bytes, sizeof(SemsegAcc));
        return STEMSEG_ERR_WORKSPACE;
    }
    int rc = require_sm100();
    if (rc != STEMSEG_OK) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SemsegAcc* acc = static_cast<SemsegAcc*>(workspace);
    SS_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(SemsegAcc), stream));
    long long blocks = (voxels + 255) / 256;
    const long long cap = 8ll * device_sm_count();
    if (blocks > cap) blocks = cap;
    const long long* ids = reinterpret_cast<const long long*>(class_ids);
    semseg_loss_reduce_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(class_logits, class_stride, n_classes,
                                                                                 fg_logits, ids, ignore, voxels, acc);
    semseg_loss_grad_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(class_logits, class_stride, n_classes, fg_logits,
                                                                               ids, ignore, voxels, acc, w_semseg,
                                                                               w_foreground, d_class, d_class_stride, d_fg,
                                                                               losses);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}
