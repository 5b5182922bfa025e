"""Unit test of the BatchNorm(+ReLU)(+avg-pool) backward kernels against torch autograd on the GPU."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw,C,cstride,pooled,relu", [(20, 96, 128, 0, 1), (40, 992, 1024, 0, 1), (40, 1024, 1024, 1, 1),
                                                       (80, 128, 128, 0, 1), (16, 64, 64, 0, 0), (160, 96, 256, 0, 1)])
def test_bn_relu_backward(hw, C, cstride, pooled, relu):
    from smg_b200 import _lib, engine
    eng = engine.get_engine(0, 2, 640, "fp32", owner="bnbwd")
    S = 2
    g = torch.Generator(device="cuda").manual_seed(hw * 7 + C)
    x_full = torch.randn((S, hw, hw, cstride), generator=g, device="cuda") * 1.5 + 0.3
    gamma = torch.rand(C, generator=g, device="cuda") + 0.5
    beta = torch.randn(C, generator=g, device="cuda") * 0.2
    hd = hw // 2 if pooled else hw
    da = torch.randn((S, hd, hd, C), generator=g, device="cuda")
    xs = x_full[..., :C].double()
    stats = torch.zeros((S, cstride, 2), dtype=torch.float64, device="cuda")
    stats[:, :C, 0] = xs.sum((1, 2))
    stats[:, :C, 1] = (xs * xs).sum((1, 2))
    # reference: autograd through per-sample train-mode BN (+ReLU) (+avg-pool)
    xr = x_full[..., :C].permute(0, 3, 1, 2).clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ys = []
    for s in range(S):
        y = F.batch_norm(xr[s:s + 1], None, None, gr, br, training=True, momentum=0.0, eps=1e-5)
        if relu:
            y = F.relu(y)
        if pooled:
            y = F.avg_pool2d(y, 2, 2)
        ys.append(y)
    y = torch.cat(ys)
    y.backward(da.permute(0, 3, 1, 2))
    dx_ref = xr.grad.permute(0, 2, 3, 1)
    # kernel
    sums = torch.zeros((S, C, 2), dtype=torch.float64, device="cuda")
    dst = torch.full((S, hw, hw, cstride), 0.5, device="cuda")
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    _lib.check(eng.lib.smg_debug_bn_bwd(eng.h, da.data_ptr(), C, pooled, x_full.data_ptr(), cstride, stats.data_ptr(), cstride,
                                        gamma.data_ptr(), beta.data_ptr(), C, hw, relu, S, sums.data_ptr(), dst.data_ptr(),
                                        cstride, 1, dg.data_ptr(), db.data_ptr(), None))
    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())
    print("dgamma %.2e dbeta %.2e dx %.2e" % (rel(dg, gr.grad), rel(db, br.grad), rel(dst[..., :C] - 0.5, dx_ref)))
    assert rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4
    assert rel(dst[..., :C] - 0.5, dx_ref) < 1e-4
    if cstride > C:
        assert float((dst[..., C:] - 0.5).abs().max()) == 0
