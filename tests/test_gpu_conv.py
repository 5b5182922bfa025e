"""Unit tests of the generic convolution kernels (conv_ffma.cu fp32, conv_umma.cu tcgen05 tf32/bf16)
through smg_debug_conv, against torch fp32 convolutions (TF32 disabled) on the same GPU.

Covers every shape class of the trunk: 1x1 with K = 64..1024 (K not a multiple of 64), partial last
M tile (HW % 128 != 0), the pooled transition, N = 64/128/256, 3x3 at every block's spatial size
(patch tiling with and without column splits), channel-slice output and the statistics epilogue."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-5, "tf32": 3e-3, "bf16": 2e-2}


@pytest.fixture(scope="module")
def eng():
    from smg_b200 import engine
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return engine.get_engine(0, 18, 640, "fp32", owner="conv")


def reference(x_nhwc, cin, scale, shift, relu, pool, w):
    x = x_nhwc[..., :cin].permute(0, 3, 1, 2).double()
    a = x * scale.double()[:, :, None, None] + shift.double()[:, :, None, None]
    if relu:
        a = a.clamp_min(0)
    if pool:
        a = F.avg_pool2d(a, 2, 2)
    y = F.conv2d(a, w.double(), padding=w.shape[-1] // 2)
    return y.permute(0, 2, 3, 1).float()


CASES = [
    # (n, hin, cin, in_cstride, cout, k, pool)
    (2, 16, 64, 256, 128, 1, 0),
    (1, 40, 96, 256, 128, 1, 0),      # HW=1600: partial last tile; K=96 (not a multiple of 64)
    (2, 20, 1024, 1024, 128, 1, 0),   # HW=400, largest K
    (1, 20, 1024, 1024, 64, 1, 0),    # head shape N=64
    (1, 32, 256, 256, 128, 1, 1),     # transition: pooled
    (1, 16, 512, 512, 256, 1, 1),     # transition with two N tiles
    (1, 160, 128, 128, 32, 3, 0),     # 3x3 block-1 size (4 column tiles)
    (2, 80, 128, 128, 32, 3, 0),
    (1, 40, 128, 128, 32, 3, 0),
    (3, 20, 128, 128, 32, 3, 0),
    (1, 8, 128, 128, 32, 3, 0),
]


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_h%d_cin%d_cs%d_cout%d_k%d_pool%d" % c)
def test_conv_matches_torch(eng, precision, case):
    n, hin, cin, cstride, cout, k, pool = case
    g = torch.Generator(device="cuda").manual_seed(hash(case) % 1000)
    x = torch.randn((n, hin, hin, cstride), generator=g, device="cuda")
    scale = torch.rand((n, cin), generator=g, device="cuda") + 0.5
    shift = torch.randn((n, cin), generator=g, device="cuda") * 0.3
    w = torch.randn((cout, cin, k, k), generator=g, device="cuda") / (cin * k * k) ** 0.5
    out_cstride, out_coff = cout + 64, 32
    out, stats = eng.debug_conv(precision, x, cin, scale, shift, True, pool, w, out_cstride, out_coff)
    ref = reference(x, cin, scale, shift, True, pool, w)
    got = out[..., out_coff:out_coff + cout]
    err = float((got - ref).abs().max() / ref.abs().max())
    print("%s %s: rel-max err %.2e" % (precision, case, err))
    assert err <= TOL[precision]
    assert float(out[..., :out_coff].abs().max()) == 0 and float(out[..., out_coff + cout:].abs().max()) == 0, \
        "wrote outside the channel slice"
    # statistics epilogue: (sum, sumsq) per (sample, channel) of what was written
    s_ref = got.double().sum((1, 2))
    q_ref = (got.double() ** 2).sum((1, 2))
    s_got, q_got = stats[:, out_coff:out_coff + cout, 0], stats[:, out_coff:out_coff + cout, 1]
    assert float((s_got - s_ref).abs().max() / s_ref.abs().max()) <= 1e-4
    assert float((q_got - q_ref).abs().max() / q_ref.abs().max()) <= 1e-4


@pytest.mark.parametrize("case", [(2, 40, 96, 256, 128, 1, 0), (1, 32, 256, 256, 128, 1, 1), (2, 80, 128, 128, 32, 3, 0)],
                         ids=lambda c: "n%d_h%d_cin%d_cs%d_cout%d_k%d_pool%d" % c)
def test_fp32_mode_cuda_core_kernel_still_matches(case, monkeypatch):
    """fp32 mode runs on the tensor cores with hi/lo split tf32 operands by default (covered by test_conv_matches_torch);
    SMG_FP32_TC=0 selects the CUDA-core FFMA kernel, which stays as the independent check of the same contract."""
    from smg_b200 import engine
    monkeypatch.setenv("SMG_FP32_TC", "0")
    eng = engine.Engine(0, 4, 640, "fp32")
    n, hin, cin, cstride, cout, k, pool = case
    g = torch.Generator(device="cuda").manual_seed(hash(case) % 1000)
    x = torch.randn((n, hin, hin, cstride), generator=g, device="cuda")
    scale = torch.rand((n, cin), generator=g, device="cuda") + 0.5
    shift = torch.randn((n, cin), generator=g, device="cuda") * 0.3
    w = torch.randn((cout, cin, k, k), generator=g, device="cuda") / (cin * k * k) ** 0.5
    out, _ = eng.debug_conv("fp32", x, cin, scale, shift, True, pool, w)
    ref = reference(x, cin, scale, shift, True, pool, w)
    assert float((out - ref).abs().max() / ref.abs().max()) <= TOL["fp32"]
    del eng


@pytest.mark.parametrize("tma_mask", [0, 64, 128])
@pytest.mark.parametrize("case", [(5, 80, 224, 256, 128, 1, 0), (2, 40, 96, 256, 128, 1, 0), (4, 80, 128, 128, 32, 3, 0),
                                  (1, 160, 128, 128, 32, 3, 0), (3, 20, 128, 128, 32, 3, 0), (2, 20, 1024, 1024, 128, 1, 0)],
                         ids=lambda c: "n%d_h%d_cin%d_cs%d_cout%d_k%d_pool%d" % c)
def test_tma_kernel_variants_match_torch(case, tma_mask, monkeypatch):
    """Every tf32 kernel selectable through SMG_TMA: 0 register producers (conv_umma.cu), 64 persistent 3x3 with the weights
    in tensor memory (conv3_wt.cu), 128 persistent 1x1 with the weights as the A operand (conv1_t.cu: tensor-memory resident
    for cin <= 256, streamed above).
    The multi-sample cases make persistent CTAs cross sample boundaries (table re-computation, statistics flush)."""
    from smg_b200 import engine
    monkeypatch.setenv("SMG_TMA", str(tma_mask))
    eng = engine.Engine(0, 5, 640, "fp32")
    n, hin, cin, cstride, cout, k, pool = case
    g = torch.Generator(device="cuda").manual_seed(hash(case) % 1000 + tma_mask)
    x = torch.randn((n, hin, hin, cstride), generator=g, device="cuda")
    scale = torch.rand((n, cin), generator=g, device="cuda") + 0.5
    shift = torch.randn((n, cin), generator=g, device="cuda") * 0.3
    w = torch.randn((cout, cin, k, k), generator=g, device="cuda") / (cin * k * k) ** 0.5
    out_cstride, out_coff = cout + 64, 32
    out, stats = eng.debug_conv("tf32", x, cin, scale, shift, True, pool, w, out_cstride, out_coff)
    ref = reference(x, cin, scale, shift, True, pool, w)
    got = out[..., out_coff:out_coff + cout]
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err <= TOL["tf32"]
    assert float(out[..., :out_coff].abs().max()) == 0 and float(out[..., out_coff + cout:].abs().max()) == 0
    s_ref, q_ref = got.double().sum((1, 2)), (got.double() ** 2).sum((1, 2))
    assert float((stats[:, out_coff:out_coff + cout, 0] - s_ref).abs().max() / s_ref.abs().max()) <= 1e-4
    assert float((stats[:, out_coff:out_coff + cout, 1] - q_ref).abs().max() / q_ref.abs().max()) <= 1e-4
    del eng


@pytest.mark.parametrize("tma_mask", [0, 32])
@pytest.mark.parametrize("case", [(3, 160, 256, 256, 128, 1, 1), (2, 80, 512, 512, 256, 1, 1), (3, 40, 1024, 1024, 512, 1, 1),
                                  (2, 24, 64, 96, 128, 1, 1), (5, 6, 512, 512, 256, 1, 1)],
                         ids=lambda c: "n%d_h%d_cin%d_cs%d_cout%d_k%d_pool%d" % c)
def test_transition_kernel_variants_match_torch(case, tma_mask, monkeypatch):
    """The pooled transition (BN-ReLU -> avg-pool 2x2 -> 1x1 conv) on both tf32 kernels: SMG_TMA bit 32 = the persistent
    TMA-fed trans_t.cu (the three densenet transitions at their real sizes: 128-column tiles, 80-column tiles, two
    output-channel parts; plus a map smaller than one tile and a channel slice of a wider buffer), 0 = conv_umma.cu."""
    from smg_b200 import engine
    monkeypatch.setenv("SMG_TMA", str(tma_mask))
    eng = engine.Engine(0, 5, 640, "fp32")
    n, hin, cin, cstride, cout, k, pool = case
    g = torch.Generator(device="cuda").manual_seed(hash(case) % 1000 + tma_mask)
    x = torch.randn((n, hin, hin, cstride), generator=g, device="cuda")
    scale = torch.rand((n, cin), generator=g, device="cuda") + 0.5
    shift = torch.randn((n, cin), generator=g, device="cuda") * 0.3
    w = torch.randn((cout, cin, k, k), generator=g, device="cuda") / (cin * k * k) ** 0.5
    out_cstride, out_coff = cout + 64, 32
    out, stats = eng.debug_conv("tf32", x, cin, scale, shift, True, pool, w, out_cstride, out_coff)
    ref = reference(x, cin, scale, shift, True, pool, w)
    got = out[..., out_coff:out_coff + cout]
    err = float((got - ref).abs().max() / ref.abs().max())
    print("mask %d %s: rel-max err %.2e" % (tma_mask, case, err))
    assert err <= TOL["tf32"]
    assert float(out[..., :out_coff].abs().max()) == 0 and float(out[..., out_coff + cout:].abs().max()) == 0
    s_ref, q_ref = got.double().sum((1, 2)), (got.double() ** 2).sum((1, 2))
    assert float((stats[:, out_coff:out_coff + cout, 0] - s_ref).abs().max() / s_ref.abs().max()) <= 1e-4
    assert float((stats[:, out_coff:out_coff + cout, 1] - q_ref).abs().max() / q_ref.abs().max()) <= 1e-4
    del eng


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_conv_no_relu_identity_prologue(eng, precision):
    x = torch.randn((1, 20, 20, 128), device="cuda")
    w = torch.randn((128, 128, 1, 1), device="cuda") / 11.3
    one, zero = torch.ones((1, 128), device="cuda"), torch.zeros((1, 128), device="cuda")
    out, _ = eng.debug_conv(precision, x, 128, one, zero, False, 0, w, want_stats=False)
    ref = reference(x, 128, one, zero, False, 0, w)
    assert float((out - ref).abs().max() / ref.abs().max()) <= TOL[precision]
