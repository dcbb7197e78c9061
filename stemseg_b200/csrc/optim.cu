// Fused SGD (momentum / Nesterov / weight decay) step over a flat fp32 parameter buffer (sm_100a).
//
// Replaces torch.optim.SGD(model.parameters(), lr, momentum, weight_decay=..., nesterov=...).step()
// (stemseg/training/utils.py:199-202, called at stemseg/training/main.py:209) for the decoder heads' parameters, which
// stemseg_b200.training keeps as views into one flat buffer.  One pass: read p, g, buf; write p, buf (20 B / parameter,
// HBM-bound).  The gradient scale folds the data-parallel average (sum over ranks -> mean) into the same pass.
//   g' = g * grad_scale + wd * p;  buf = momentum * buf + g'  (buf starts at 0, which equals torch's first-step copy);
//   step = nesterov ? g' + momentum * buf : buf;  p -= lr * step
#include "common.cuh"

namespace stemseg {
namespace {

__device__ __forceinline__ void sgd_update(float& p, float g, float& buf, float lr, float momentum, float wd,
                                           float grad_scale, int nesterov) {
    const float gs = __fadd_rn(__fmul_rn(g, grad_scale), __fmul_rn(wd, p));
    const float b = __fadd_rn(__fmul_rn(momentum, buf), gs);
    const float step = nesterov ? __fadd_rn(gs, __fmul_rn(momentum, b)) : b;
    buf = b;
    p = __fsub_rn(p, __fmul_rn(lr, step));
}

// hyper: optional device array {lr, momentum, weight_decay, grad_scale}; when given it overrides the by-value arguments,
// so that a CUDA graph that captured this launch follows a learning-rate schedule (the reference steps its scheduler
// every iteration, stemseg/training/main.py:209-210) without being re-captured
__global__ void __launch_bounds__(256) sgd_nesterov_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                                           float* __restrict__ buf, long long n, float lr, float momentum,
                                                           float wd, float grad_scale, int nesterov,
                                                           const float* __restrict__ hyper) {
    if (hyper != nullptr) {
        lr = __ldg(hyper + 0);
        momentum = __ldg(hyper + 1);
        wd = __ldg(hyper + 2);
        grad_scale = __ldg(hyper + 3);
    }
    const long long stride = 1ll * gridDim.x * blockDim.x;
    const long long n4 = n / 4;
    float4* p4 = reinterpret_cast<float4*>(param);
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    float4* b4 = reinterpret_cast<float4*>(buf);
    for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 p = p4[i], b = b4[i];
        const float4 g = g4[i];
        sgd_update(p.x, g.x, b.x, lr, momentum, wd, grad_scale, nesterov);
        sgd_update(p.y, g.y, b.y, lr, momentum, wd, grad_scale, nesterov);
        sgd_update(p.z, g.z, b.z, lr, momentum, wd, grad_scale, nesterov);
        sgd_update(p.w, g.w, b.w, lr, momentum, wd, grad_scale, nesterov);
        p4[i] = p;
        b4[i] = b;
    }
    for (long long i = n4 * 4 + 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        sgd_update(param[i], grad[i], buf[i], lr, momentum, wd, grad_scale, nesterov);
}

}  // namespace
}  // namespace stemseg

using namespace stemseg;

static int32_t sgd_step_impl(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                             float weight_decay, float grad_scale, int32_t nesterov, const float* hyper, void* stream_);

extern "C" int32_t stemseg_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr,
                                    float momentum, float weight_decay, float grad_scale, int32_t nesterov,
                                    void* stream_) {
    return sgd_step_impl(param, grad, momentum_buf, n, lr, momentum, weight_decay, grad_scale, nesterov, nullptr, stream_);
}

extern "C" int32_t stemseg_sgd_step_dev(float* param, const float* grad, float* momentum_buf, int64_t n,
                                        const float* hyper, int32_t nesterov, void* stream_) {
    SS_REQUIRE(hyper != nullptr, "sgd_step_dev: null hyper-parameter array");
    return sgd_step_impl(param, grad, momentum_buf, n, 0.f, 0.f, 0.f, 1.f, nesterov, hyper, stream_);
}

static int32_t sgd_step_impl(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                             float weight_decay, float grad_scale, int32_t nesterov, const float* hyper, void* stream_) {
    SS_REQUIRE(param && grad && momentum_buf && n >= 0, "sgd_step: bad arguments");
    SS_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                 reinterpret_cast<uintptr_t>(momentum_buf)) & 15) == 0,
               "sgd_step: pointers must be 16-byte aligned");
    if (n == 0) return STEMSEG_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = 8ll * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    sgd_nesterov_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(param, grad, momentum_buf, n, lr, momentum,
                                                                           weight_decay, grad_scale, nesterov, hyper);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}
