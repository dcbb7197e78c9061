"""Developer tool: op-by-op check of the backward kernels against torch autograd on the GPU.

usage: python scripts/debug_backward.py <op>     (ops: head upsample pool gn chsum dgrad dgrad1 wgrad wgrad1 all)
"""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from stemseg_b200 import _lib, decoder as D, autograd as A     # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))


def ndhwc(x):        # [n,c,t,h,w] -> contiguous [n,t,h,w,c]
    return x.permute(0, 2, 3, 4, 1).contiguous()


def ncthw(x):
    return x.permute(0, 4, 1, 2, 3)


def op_upsample():
    for ts in (1, 2):
        n, t, h, w, c = 1, 4, 12, 16, 32
        lo = torch.randn(n, c, t // ts, h // 2, w // 2, device=dev, requires_grad=True)
        g = torch.randn(n, c, t, h, w, device=dev)
        F.interpolate(lo, scale_factor=(ts, 2, 2), mode="trilinear", align_corners=False).backward(g)
        d_high = ndhwc(g)
        d_low = torch.empty(n, t // ts, h // 2, w // 2, c, device=dev)
        _lib.check(lib.stemseg_upsample_transpose(_lib.ptr(d_high), n, t, h, w, c, ts, _lib.ptr(d_low), _lib.stream_ptr()))
        torch.cuda.synchronize()
        print("upsample_transpose ts=%d" % ts, rel(ncthw(d_low), lo.grad))


def op_pool():
    for pool in (0, 1):
        for norm in (0, 1):
            n, t, h, w, c = 1, 5, 6, 8, 32
            y = torch.randn(n, c, t, h, w, device=dev, requires_grad=True)
            sc = torch.randn(c, device=dev) if norm else torch.ones(c, device=dev)
            sh = torch.randn(c, device=dev) if norm else torch.zeros(c, device=dev)
            v = F.relu(y * sc.view(1, c, 1, 1, 1) + sh.view(1, c, 1, 1, 1))
            v.retain_grad()
            o = F.avg_pool3d(v, 3, stride=(2, 1, 1), padding=1) if pool else v
            g = torch.randn_like(o)
            o.backward(g)
            # d_norm = gradient wrt the normalised value (before ReLU) = v.grad masked
            want = v.grad * (v > 0)
            ss = torch.stack([sc, sh], 1).view(1, c, 2).contiguous() if norm else None
            d_norm = torch.empty(n, t, h, w, c, device=dev)
            yy = ndhwc(y.detach())
            gg = ndhwc(g)
            _lib.check(lib.stemseg_pool_relu_backward(_lib.ptr(gg), _lib.ptr(yy), _lib.ptr(ss), n, t, h, w, c, pool,
                                                      _lib.ptr(d_norm), _lib.stream_ptr()))
            torch.cuda.synchronize()
            print("pool_relu_backward pool=%d norm=%d" % (pool, norm), rel(ncthw(d_norm), want))


def op_gn():
    n, t, h, w, c, groups = 1, 4, 6, 8, 64, 32
    y = torch.randn(n, c, t, h, w, device=dev, requires_grad=True)
    gamma = torch.randn(c, device=dev, requires_grad=True)
    beta = torch.randn(c, device=dev, requires_grad=True)
    o = F.group_norm(y, groups, gamma, beta, eps=1e-5)
    g = torch.randn_like(o)
    o.backward(g)
    yy = ndhwc(y.detach())
    yg = yy.view(n, -1, groups, c // groups)
    mean = yg.mean(dim=(1, 3))
    var = yg.var(dim=(1, 3), unbiased=False)
    mean_rstd = torch.stack([mean, (var + 1e-5).rsqrt()], -1).contiguous()
    dn = ndhwc(g).clone()
    dgb = torch.empty(n, c, 2, device=dev)
    gt = torch.empty(n, groups, 2, device=dev)
    wsb = lib.stemseg_group_norm_backward_workspace_bytes(n, t * h * w, c)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    _lib.check(lib.stemseg_group_norm_backward(_lib.ptr(dn), _lib.ptr(yy), _lib.ptr(mean_rstd), _lib.ptr(gamma.detach()),
                                               n, t * h * w, c, c // groups, _lib.ptr(dgb), _lib.ptr(gt), _lib.ptr(ws),
                                               wsb, _lib.stream_ptr()))
    torch.cuda.synchronize()
    print("gn_backward dy", rel(ncthw(dn), y.grad), "dgamma", rel(dgb[0, :, 0], gamma.grad), "dbeta",
          rel(dgb[0, :, 1], beta.grad))


def op_chsum():
    rows, c = 4 * 6 * 8 * 7, 64
    x = torch.randn(rows, c, device=dev)
    out = torch.empty(c, device=dev)
    wsb = lib.stemseg_channel_sum_workspace_bytes(rows, c)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    _lib.check(lib.stemseg_channel_sum(_lib.ptr(x), rows, c, _lib.ptr(out), _lib.ptr(ws), wsb, _lib.stream_ptr()))
    torch.cuda.synchronize()
    print("channel_sum", rel(out, x.sum(0)))


def op_head():
    for ts in (1, 2):
        n, t, h, w, c, j = 1, 4, 8, 16, 32, 7
        z = torch.randn(n, c, t, h, w, device=dev, requires_grad=True)
        yl = torch.randn(n, c, t // ts, h // 2, w // 2, device=dev, requires_grad=True)
        wt = (0.3 * torch.randn(j, c, device=dev)).requires_grad_(True)
        bs = torch.randn(j, device=dev, requires_grad=True)
        act = [1, 1, 0, 0, 0, 0, 2]
        x = z + F.interpolate(yl, scale_factor=(ts, 2, 2), mode="trilinear", align_corners=False)
        x.retain_grad()
        lin = F.conv3d(x, wt.view(j, c, 1, 1, 1), bs)
        outs = []
        for k in range(j):
            v = lin[:, k]
            outs.append((v * 0.25).tanh() if act[k] == 1 else (v.sigmoid() if act[k] == 2 else v))
        o = torch.stack(outs, 1)
        g = torch.randn_like(o)
        o.backward(g)
        dx = torch.empty(n, t, h, w, c, device=dev)
        dwt = torch.empty(j, c, device=dev)
        dbs = torch.empty(j, device=dev)
        wsb = lib.stemseg_head_backward_workspace_bytes(c)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        acts = torch.tensor(act, dtype=torch.int32, device=dev)
        zz, yy = ndhwc(z.detach()), ndhwc(yl.detach())           # keep the temporaries alive across the launch
        _lib.check(lib.stemseg_head_backward(_lib.ptr(zz), _lib.ptr(yy), n, t, h, w, c, ts,
                                             _lib.ptr(wt.detach()), _lib.ptr(bs.detach()), _lib.ptr(acts), j,
                                             _lib.ptr(g), _lib.ptr(dx), _lib.ptr(dwt), _lib.ptr(dbs),
                                             _lib.ptr(ws), wsb, _lib.stream_ptr()))
        torch.cuda.synchronize()
        print("head_backward ts=%d dx" % ts, rel(ncthw(dx), x.grad), "dW", rel(dwt, wt.grad), "db", rel(dbs, bs.grad))


def _dgrad(ksize, cin, cout, t, h, w):
    planes = 2
    x = torch.randn(1, cin, t, h, w, device=dev, requires_grad=True)
    wt = (torch.randn(cout, cin, ksize, ksize, ksize, device=dev) / (cin * ksize ** 3) ** 0.5).requires_grad_(True)
    o = F.conv3d(x, wt, None, padding=ksize // 2)
    g = torch.randn_like(o)
    o.backward(g)
    taps = ksize ** 3
    dst = torch.empty(planes, cin, taps * cout, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.stemseg_pack_conv_weight_dgrad(_lib.ptr(wt.detach().contiguous()), cout, cin, 0, cin, taps,
                                                  _lib.ptr(dst), planes, _lib.stream_ptr()))
    torch.cuda.synchronize()
    packed = D.PackedConv(dst, None, cout, cin, ksize)
    dy = ndhwc(g)
    dx = D.conv3d(A._to_planes(dy, planes), packed)
    torch.cuda.synchronize()
    print("dgrad k=%d cin=%d cout=%d %dx%dx%d" % (ksize, cin, cout, t, h, w), rel(ncthw(dx), x.grad))
    return x, wt, dy


def op_dgrad():
    _dgrad(3, 32, 64, 4, 6, 8)
    _dgrad(3, 64, 32, 2, 12, 16)


def op_dgrad1():
    _dgrad(1, 32, 64, 4, 6, 8)
    _dgrad(1, 96, 32, 2, 12, 16)


def _wgrad(ksize, cin, cout, t, h, w):
    planes = 2
    x = torch.randn(1, cin, t, h, w, device=dev)
    wt = (torch.randn(cout, cin, ksize, ksize, ksize, device=dev) / (cin * ksize ** 3) ** 0.5).requires_grad_(True)
    o = F.conv3d(x, wt, None, padding=ksize // 2)
    g = torch.randn_like(o)
    o.backward(g)
    xp = D.pack_activation(x, planes)
    dst = torch.empty(cout, cin, ksize ** 3, device=dev)
    dy = ndhwc(g)
    dyp = A._to_planes(dy, planes)
    A._wgrad(dy, dyp, xp, ksize, planes, dst, 0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    A._wgrad(dy, dyp, xp, ksize, planes, dst, 0)
    b.record()
    torch.cuda.synchronize()
    print("wgrad[%s] k=%d cin=%d cout=%d %dx%dx%d" % (A.WGRAD_MODE, ksize, cin, cout, t, h, w),
          "rel %.3g" % rel(dst.view_as(wt), wt.grad), "%.1f us" % (1e3 * a.elapsed_time(b)))


def op_wgrad():
    _wgrad(3, 32, 32, 2, 4, 8)
    _wgrad(3, 32, 64, 4, 6, 8)
    _wgrad(3, 64, 32, 8, 12, 16)


def op_wgrad1():
    _wgrad(1, 32, 32, 2, 4, 8)
    _wgrad(1, 64, 96, 4, 12, 16)


def op_wgradbig():
    _wgrad(3, 128, 128, 4, 24, 40)
    _wgrad(3, 256, 256, 4, 12, 20)
    _wgrad(1, 256, 128, 8, 24, 40)
    _wgrad(3, 256, 128, 8, 96, 160)


OPS = {k[3:]: v for k, v in list(globals().items()) if k.startswith("op_")}

if __name__ == "__main__":
    which = sys.argv[1:] or ["all"]
    if which == ["all"]:
        which = list(OPS)
    for name in which:
        OPS[name]()
