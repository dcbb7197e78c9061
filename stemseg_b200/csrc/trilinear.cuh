// Trilinear (align_corners=False) upsampling taps and torch.linspace coordinates shared by the decoder kernels.
#pragma once

#include <cstddef>

namespace stemseg {

// ---------------------------------------------------------------------------------------------------------------
// trilinear (align_corners=False) source taps for integer scale 1 or 2 along one axis (common.py:69-78)
// ---------------------------------------------------------------------------------------------------------------
struct Tap {
    int i0, i1;
    float w0, w1;
};
__device__ __forceinline__ Tap axis_tap(int dst, int scale, int src_size) {
    Tap tp;
    if (scale == 1) {
        tp.i0 = tp.i1 = dst; tp.w0 = 1.f; tp.w1 = 0.f;
        return tp;
    }
    float s = (static_cast<float>(dst) + 0.5f) * 0.5f - 0.5f;       // area_pixel_compute_source_index
    if (s < 0.f) s = 0.f;
    tp.i0 = static_cast<int>(s);
    tp.i1 = tp.i0 + (tp.i0 < src_size - 1 ? 1 : 0);
    tp.w1 = s - static_cast<float>(tp.i0);
    tp.w0 = 1.f - tp.w1;
    return tp;
}

struct Tri {
    size_t off[8];
    float wgt[8];
};
__device__ __forceinline__ Tri make_tri(int nn, int to, int ho, int wo, int st, int tl, int hl, int wl, int c) {
    const Tap a = axis_tap(to, st, tl), b = axis_tap(ho, 2, hl), d = axis_tap(wo, 2, wl);
    Tri r;
    const int ti[2] = {a.i0, a.i1}, hi[2] = {b.i0, b.i1}, wi[2] = {d.i0, d.i1};
    const float tw[2] = {a.w0, a.w1}, hw_[2] = {b.w0, b.w1}, ww[2] = {d.w0, d.w1};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int x = i & 1, y = (i >> 1) & 1, z = i >> 2;
        r.off[i] = (((static_cast<size_t>(nn) * tl + ti[z]) * hl + hi[y]) * wl + wi[x]) * c;
        r.wgt[i] = tw[z] * hw_[y] * ww[x];
    }
    return r;
}


__device__ __forceinline__ float linspace_value(float end_abs, int steps, int i) {
    // torch.linspace(-a, a, steps) fp32: start + i*step for the first half, end - (steps-1-i)*step for the second
    if (steps == 1) return -end_abs;
    const float step = (end_abs - (-end_abs)) / static_cast<float>(steps - 1);
    return i < steps / 2 ? (-end_abs + step * static_cast<float>(i)) : (end_abs - step * static_cast<float>(steps - 1 - i));
}


}  // namespace stemseg
