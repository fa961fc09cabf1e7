"""GPU: every hand-written kernel against a plain PyTorch fp32 reference of the same op.
Inputs are rounded to fp16 first (the kernels' storage type) so the tolerance only has to cover
fp32-accumulated fp16 products and the fp16 rounding of the result."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def h(x):
    return x.half().float()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


def g(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


def rn(*s, seed=0):
    return torch.randn(*s, generator=g(seed), device="cuda")


@pytest.mark.parametrize("M,K,N", [(128, 64, 32), (256, 320, 320), (1000, 768, 1280), (65, 1280, 640), (4096, 320, 960),
                                   (8, 136, 768), (300, 2560, 1280),
                                   # long-K shapes whose wave-aware tile width is not a multiple of 32 (80 / 144 / 224)
                                   (1024, 2560, 1280), (4096, 1280, 1280), (1024, 1280, 3840)])
def test_linear_tcgen05(engine, M, K, N):
    x, w, b = h(rn(M, K, seed=1)), h(rn(N, K, seed=2) / math.sqrt(K)), rn(N, seed=3)
    y = engine.op_linear(x, w, b)
    ref = F.linear(x, w, b)
    assert rel(y, ref) < 4e-3, rel(y, ref)


@pytest.mark.parametrize("kind", ["linear", "linear_res", "geglu", "conv"])
def test_cta_pair_kernel(engine, kind):
    """gemm_pair.cuh (tcgen05 cta_group::2: 256-row tiles over a CTA pair, each CTA loading half of the B tile) is opt-in
    (option gemm_pair; slower than the 1-CTA kernel on the UNet's shapes) but must stay correct: same K order per output
    element as the default kernel, so the results are the same bits."""
    def run():
        if kind == "conv":
            x, w, b = h(rn(4, 640, 32, 32, seed=1)), h(rn(640, 640, 3, 3, seed=2) / math.sqrt(9 * 640)), rn(640, seed=3)
            return engine.op_conv2d(x, w, b), F.conv2d(x, w, b, padding=1)
        M, K, N = 4096, 1280, 1280
        x, w, b = h(rn(M, K, seed=1)), h(rn(N * (2 if kind == "geglu" else 1), K, seed=2) / math.sqrt(K)), rn(N * (2 if kind == "geglu" else 1), seed=3)
        if kind == "geglu":
            a, gate = F.linear(x, w, b).chunk(2, dim=-1)
            return engine.op_linear(x, w, b, geglu=True), a * F.gelu(gate)
        r = h(rn(M, N, seed=4)) if kind == "linear_res" else None
        return engine.op_linear(x, w, b, residual=r), F.linear(x, w, b) + (r if r is not None else 0)
    y1, ref = run()
    engine.set_option("gemm_pair", 1)
    try:
        y2, _ = run()
    finally:
        engine.set_option("gemm_pair", 0)
    assert rel(y2, ref) < 4e-3, rel(y2, ref)
    assert torch.equal(y1, y2)


@pytest.mark.parametrize("act,fn", [(1, F.silu), (2, F.gelu), (3, lambda v: v * torch.sigmoid(1.702 * v))])
def test_linear_epilogues(engine, act, fn):
    M, K, N = 512, 640, 640
    x, w, b, r = h(rn(M, K, seed=1)), h(rn(N, K, seed=2) / math.sqrt(K)), rn(N, seed=3), h(rn(M, N, seed=4))
    y = engine.op_linear(x, w, b, residual=r, act=act)
    ref = fn(F.linear(x, w, b)) + r
    assert rel(y, ref) < 4e-3, rel(y, ref)


@pytest.mark.parametrize("C", [320, 640, 1280])
def test_geglu(engine, C):
    M = 384
    x, w, b = h(rn(M, C, seed=1)), h(rn(8 * C, C, seed=2) / math.sqrt(C)), rn(8 * C, seed=3)
    y = engine.op_linear(x, w, b, geglu=True)
    a, gate = F.linear(x, w, b).chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    assert rel(y, ref) < 4e-3, rel(y, ref)


@pytest.mark.parametrize("N,C,H,O", [(2, 64, 8, 64), (2, 320, 64, 320), (1, 640, 32, 1280), (3, 128, 16, 96),
                                     (2, 1280, 8, 1280), (1, 128, 128, 128), (1, 64, 256, 64), (4, 64, 4, 64), (2, 64, 2, 32),
                                     (16, 1280, 8, 1280), (16, 640, 16, 1280)])
def test_conv3x3_implicit_gemm_tma(engine, N, C, H, O):
    x, w, b = h(rn(N, C, H, H, seed=1)), h(rn(O, C, 3, 3, seed=2) / math.sqrt(9 * C)), rn(O, seed=3)
    y = engine.op_conv2d(x, w, b)
    ref = F.conv2d(x, w, b, padding=1)
    assert rel(y, ref) < 4e-3, rel(y, ref)


@pytest.mark.parametrize("N,C,H,O,k,s,pad", [(2, 9, 16, 320, 3, 1, (1, 1, 1, 1)), (2, 320, 16, 320, 3, 2, (1, 1, 1, 1)),
                                             (1, 128, 32, 128, 3, 2, (0, 0, 1, 1)), (2, 3, 28, 64, 3, 1, (1, 1, 1, 1)),
                                             (2, 64, 14, 128, 1, 2, (0, 0, 0, 0)), (1, 3, 56, 128, 14, 14, (0, 0, 0, 0)),
                                             (2, 64, 28, 64, 3, 1, (1, 1, 1, 1)), (2, 4, 8, 512, 3, 1, (1, 1, 1, 1)),
                                             (2, 960, 8, 320, 1, 1, (0, 0, 0, 0)), (2, 320, 64, 320, 3, 2, (1, 1, 1, 1)),
                                             (1, 128, 256, 128, 3, 2, (0, 0, 1, 1)), (4, 640, 32, 640, 3, 2, (1, 1, 1, 1))])
@pytest.mark.parametrize("tma_s2", [0, 1])
def test_conv_generic_im2col(engine, N, C, H, O, k, s, pad, tma_s2):
    """tma_s2=1: stride-2 3x3 convolutions with Cin % 64 == 0 run as implicit GEMM through strided TMA boxes
    (option conv_tma_stride2) instead of an explicit im2col; every other case takes the im2col path either way."""
    x, w, b = h(rn(N, C, H, H, seed=1)), h(rn(O, C, k, k, seed=2) / math.sqrt(k * k * C)), rn(O, seed=3)
    engine.set_option("conv_tma_stride2", tma_s2)
    try:
        y = engine.op_conv2d(x, w, b, stride=s, pad=pad)
    finally:
        engine.set_option("conv_tma_stride2", 1)
    pt, pl, pb, pr = pad
    ref = F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, b, stride=s)
    assert y.shape == ref.shape
    assert rel(y, ref) < 4e-3, rel(y, ref)


@pytest.mark.parametrize("N,C,H,eps,silu", [(2, 320, 16, 1e-5, True), (2, 960, 8, 1e-5, True), (1, 128, 64, 1e-6, True),
                                            (3, 1280, 4, 1e-6, False), (2, 2560, 8, 1e-5, True), (2, 1920, 2, 1e-5, True),
                                            (1, 512, 32, 1e-6, False), (5, 640, 32, 1e-5, True)])
@pytest.mark.parametrize("fused", [1, 0])
def test_groupnorm(engine, N, C, H, eps, silu, fused):
    """fused=1 (default for small / medium maps): one launch, a cluster of 16 CTAs per sample exchanging statistics
    through DSMEM; fused=0: the two-launch whole-grid path (statistics, then apply with the finalize folded in)."""
    x = h(rn(N, C, H, H, seed=1) * 2 + 0.5)
    gam, bet = 1 + 0.1 * rn(C, seed=2), 0.1 * rn(C, seed=3)
    engine.set_option("gn_fused", fused)
    engine.set_option("gn_fused_max_elems", 1 << 40)      # exercise the cluster kernel at every test shape
    try:
        y = engine.op_groupnorm(x, gam, bet, eps, silu)
    finally:
        engine.set_option("gn_fused", 1)
        engine.set_option("gn_fused_max_elems", 2621440)
    ref = F.group_norm(x, 32, gam, bet, eps)
    ref = F.silu(ref) if silu else ref
    assert float((y - ref).abs().max()) < 6e-3


@pytest.mark.parametrize("N,C,O,H,k,stride", [(2, 320, 320, 64, 3, 1), (2, 640, 1280, 16, 3, 1), (3, 320, 640, 32, 1, 1),
                                              (2, 320, 320, 32, 3, 2), (1, 128, 320, 16, 3, 1), (2, 1280, 1280, 8, 3, 1)])
def test_groupnorm_from_epilogue_statistics(engine, N, C, O, H, k, stride):
    """conv -> GroupNorm(32) + SiLU (every ResBlock): the convolution's epilogue leaves per-(32 rows, channel) partial
    sums with its output, GroupNorm folds them and makes ONE streaming pass.  Same result as the stand-alone kernels
    (fp32 statistics of the same fp16 values); the 8x8 case takes the split-K conv, whose reduction kernel writes them."""
    x, w, b = h(rn(N, C, H, H, seed=1)), h(rn(O, C, k, k, seed=2) / math.sqrt(k * k * C)), rn(O, seed=3)
    gam, bet = 1 + 0.1 * rn(O, seed=4), 0.1 * rn(O, seed=5)
    p = (k // 2,) * 4
    y = engine.op_conv2d(x, w, b, stride=stride, pad=p, gn=(gam, bet))
    engine.set_option("gn_epi_stats", 0)
    try:
        y0 = engine.op_conv2d(x, w, b, stride=stride, pad=p, gn=(gam, bet))
    finally:
        engine.set_option("gn_epi_stats", 1)
    ref = F.silu(F.group_norm(F.conv2d(x, w, b, stride=stride, padding=k // 2), 32, gam, bet, 1e-5))
    assert float((y - ref).abs().max()) < 8e-3 and float((y0 - ref).abs().max()) < 8e-3
    assert float((y - y0).abs().max()) < 2e-3          # only the summation order of the statistics differs


@pytest.mark.parametrize("N,N2,C1,C2,H", [(2, 2, 1280, 640, 8), (4, 2, 320, 320, 16), (2, 2, 640, 320, 32), (2, 1, 1280, 1280, 8),
                                          (2, 2, 640, 320, 64)])
@pytest.mark.parametrize("fused", [1, 0])
def test_groupnorm_of_concat(engine, N, N2, C1, C2, H, fused):
    """GroupNorm(32) over torch.cat([h, skip], 1) (openaimodel.py:897-899 + ResBlock in_layers) read from the two
    tensors in place; the second source may hold N2 < N samples (sample n reads n mod N2: the tensor the two CFG halves
    share).  Groups straddle the seam (e.g. 960 channels = 640 + 320: 30 per group)."""
    a = h(rn(N, C1, H, H, seed=1) * 1.5 + 0.3)
    b = h(rn(N2, C2, H, H, seed=2) * 0.7 - 0.2)
    C = C1 + C2
    gam, bet = 1 + 0.1 * rn(C, seed=3), 0.1 * rn(C, seed=4)
    cat = torch.cat([a, b.repeat(N // N2, 1, 1, 1)], 1)
    engine.set_option("gn_fused", fused)
    try:
        y = engine.op_groupnorm(a, gam, bet, 1e-5, True, x2=b)
        y_cat = engine.op_groupnorm(cat, gam, bet, 1e-5, True)
    finally:
        engine.set_option("gn_fused", 1)
    ref = F.silu(F.group_norm(cat, 32, gam, bet, 1e-5))
    assert float((y - ref).abs().max()) < 6e-3
    assert torch.equal(y, y_cat)        # same partition and summation order as over the materialised concatenation


@pytest.mark.parametrize("M,M2,K1,K2,N", [(4096, 4096, 1280, 640, 1280), (8192, 4096, 320, 320, 320), (1024, 1024, 1280, 1280, 1280),
                                          (2048, 2048, 640, 320, 640)])
def test_linear_of_concat(engine, M, M2, K1, K2, N):
    """1x1 skip_connection conv of a ResBlock whose input is torch.cat([h, skip], 1): one K loop over two TMA descriptors,
    bitwise equal to the GEMM over the materialised concatenation (same K order)."""
    a, b = h(rn(M, K1, seed=1)), h(rn(M2, K2, seed=2))
    w, bias = h(rn(N, K1 + K2, seed=3) / math.sqrt(K1 + K2)), rn(N, seed=4)
    y = engine.op_linear(a, w, bias, x2=b)
    cat = torch.cat([a, b.repeat(M // M2, 1)], 1)
    ref = F.linear(cat, w, bias)
    assert rel(y, ref) < 4e-3, rel(y, ref)
    assert torch.equal(y, engine.op_linear(cat, w, bias))


@pytest.mark.parametrize("N,C,O,H", [(2, 1280, 1280, 8), (2, 1280, 1280, 16), (1, 640, 640, 32), (1, 512, 512, 64), (1, 256, 256, 256),
                                     (3, 128, 64, 4)])
def test_upsample_conv_folded(engine, N, C, O, H):
    """Upsample.forward (openaimodel.py:109-119; model.py:53-66): conv3x3(nearest_2x(x)) as four phase-wise 2x2
    convolutions over the low-resolution input with pre-summed weights (4/9 of the MACs, no up-sampled tensor)."""
    x, w, b = h(rn(N, C, H, H, seed=1)), h(rn(O, C, 3, 3, seed=2) / math.sqrt(9 * C)), rn(O, seed=3)
    y = engine.op_upconv(x, w, b)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, b, padding=1)
    assert y.shape == ref.shape
    assert rel(y, ref) < 4e-3, rel(y, ref)


@pytest.mark.parametrize("rows,C", [(100, 320), (4096, 640), (17, 1280), (257, 1024), (5, 768)])
@pytest.mark.parametrize("vec", [1, 0])
def test_layernorm(engine, rows, C, vec):
    """vec=1 (default): 16-byte vectorised kernel (several lanes per row); vec=0: one warp per row."""
    x = h(rn(rows, C, seed=1) * 3 + 1)
    gam, bet = 1 + 0.1 * rn(C, seed=2), 0.1 * rn(C, seed=3)
    engine.set_option("ln_vec", vec)
    try:
        y = engine.op_layernorm(x, gam, bet)
    finally:
        engine.set_option("ln_vec", 1)
    assert float((y - F.layer_norm(x, (C,), gam, bet)).abs().max()) < 6e-3


@pytest.mark.parametrize("N,L,heads,d", [(2, 256, 8, 40), (1, 1024, 8, 80), (2, 64, 8, 160), (1, 257, 16, 64),
                                         (1, 256, 1, 512), (2, 16, 8, 40), (1, 4096, 8, 40), (3, 384, 8, 40),
                                         (2, 1024, 8, 40), (2, 256, 8, 160), (3, 200, 8, 80), (2, 129, 4, 64), (1, 1, 2, 40)])
@pytest.mark.parametrize("flash", [0, 3, 4])
@pytest.mark.parametrize("gain", [1.0, 3.0])
def test_attention(engine, N, L, heads, d, flash, gain):
    """flash=4 (default): fused tcgen05 kernels (two query tiles per CTA for d=40, L % 256 == 0; one tile per CTA
    otherwise; head dims 40 / 64 / 80 / 160, any sequence length -- keys beyond L are masked in the last tile), O
    accumulated in TMEM with lazy rescaling; flash=3: one query tile per CTA everywhere; flash=0 (and d=512, the VAE's
    AttnBlock): S/P materialised.
    gain=3 makes the scores ~9x larger so that the running maximum moves by more than the lazy-rescale threshold."""
    if gain != 1.0 and (flash == 0 or d not in (40, 64, 80, 160)):
        pytest.skip("the materialised path stores the scores in fp16: only exercised at unit gain")
    C = heads * d
    qkv = h(rn(N, L, 3 * C, seed=1))
    qkv[..., :2 * C] *= gain
    qkv = h(qkv)
    engine.set_option("attn_flash", flash)
    try:
        y = engine.op_attention(qkv, heads)
    finally:
        engine.set_option("attn_flash", 4)
    q, k, v = qkv.chunk(3, dim=-1)
    sp = lambda t: t.reshape(N, L, heads, d).transpose(1, 2)
    ref = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(N, L, C)
    assert float((y - ref).abs().max()) < 8e-3, float((y - ref).abs().max())


def test_concat9_bit_exact_and_ddim_update(engine, oracle):
    gen = torch.Generator().manual_seed(7)
    B, L = 3, 16
    x, z = torch.randn(B, 4, L, L, generator=gen), torch.randn(B, 4, L, L, generator=gen)
    m = torch.rand(B, 1, L, L, generator=gen)
    out = engine.concat9(x, z, m, dup=2).cpu()
    ref = torch.cat([torch.cat([x, z, m], 1)] * 2)
    assert torch.equal(out, ref)                      # pure data movement: bit exact
    eps2 = torch.randn(2 * B, 4, L, L, generator=gen)
    sch = oracle.ddim_schedule(50)
    for idx in (49, 20, 0):
        a = [sch[k][idx] for k in ("a_t", "a_prev", "sigma", "sqrt_one_minus_a")]
        xp, p0 = engine.cfg_ddim_update(x, eps2, 3.5, *a)
        rxp, rp0, _ = oracle.cfg_ddim_update(x, eps2[:B], eps2[B:], 3.5, *a)
        assert torch.equal(xp.cpu(), rxp) and torch.equal(p0.cpu(), rp0)   # same fp32 op order: bit exact


@pytest.mark.parametrize("batch", [1, 8, 30])
def test_split_k_is_batch_independent(engine, batch):
    """The long-K 3x3 convolutions of <= 8x8 maps run as 3 K-slices of 128x256 tiles (fp32 partial tiles, fixed-order
    reduction that also applies the bias).  The slice count depends on the per-sample shape only: sample 0 of any batch
    is bitwise the batch-1 result; switching the option off gives the single-pass kernel within the usual tolerance."""
    x, w, b = h(rn(batch, 1280, 8, 8, seed=1)), h(rn(1280, 1280, 3, 3, seed=2) / math.sqrt(9 * 1280)), rn(1280, seed=3)
    y = engine.op_conv2d(x, w, b)
    ref = F.conv2d(x, w, b, padding=1)
    assert rel(y, ref) < 4e-3, rel(y, ref)
    y1 = engine.op_conv2d(x[:1], w, b)
    assert torch.equal(y[:1], y1)
    engine.set_option("gemm_splitk", 0)
    try:
        y0 = engine.op_conv2d(x, w, b)
    finally:
        engine.set_option("gemm_splitk", 1)
    assert rel(y0, ref) < 4e-3 and not torch.equal(y0, y)
    # the K slices run as clusters of CTAs sharing each weight tile by TMA multicast (gemm_mcast.cuh): a pure
    # work-distribution change, the bits are those of the 1-CTA kernel
    engine.set_option("gemm_mcast", 0)
    try:
        y_nomc = engine.op_conv2d(x, w, b)
    finally:
        engine.set_option("gemm_mcast", 1)
    assert torch.equal(y_nomc, y)


def test_long_k_weight_multicast_is_bit_identical(engine):
    """Every long-K launch (K >= 1024) that fills the GPU runs as clusters of 2 CTAs on consecutive M tiles that share
    each weight tile by TMA multicast (option gemm_mcast_big; launch_gemm in engine.cu): plain linears with bias +
    residual, GEGLU, 3x3 convolutions with a time-embedding row vector and GroupNorm statistics in the epilogue.  Work
    distribution only: the bits are the 1-CTA kernel's, and the results stay batch-independent."""
    x, w, b = h(rn(4096, 5120, seed=1)), h(rn(1280, 5120, seed=2) / math.sqrt(5120)), rn(1280, seed=3)
    r = h(rn(4096, 1280, seed=4))
    xg, wg, bg = h(rn(4096, 1280, seed=5)), h(rn(2 * 5120, 1280, seed=6) / math.sqrt(1280)), rn(2 * 5120, seed=7)
    xc, wc, bc = h(rn(8, 640, 32, 32, seed=8)), h(rn(640, 640, 3, 3, seed=9) / math.sqrt(9 * 640)), rn(640, seed=10)

    xd, wd, bd = h(rn(4, 320, 64, 64, seed=11)), h(rn(320, 320, 3, 3, seed=12) / math.sqrt(9 * 320)), rn(320, seed=13)

    def run():
        return (engine.op_linear(x, w, b, residual=r), engine.op_linear(xg, wg, bg, geglu=True),
                engine.op_conv2d(xc, wc, bc), engine.op_conv2d(xc[:2], wc, bc), engine.op_conv2d(xd, wd, bd))
    on = run()
    engine.set_option("gemm_mcast_big", 0)
    try:
        off = run()
    finally:
        engine.set_option("gemm_mcast_big", 2)
    for a, o in zip(on, off):
        assert torch.equal(a, o)
    assert rel(on[4], F.conv2d(xd, wd, bd, padding=1)) < 4e-3
    assert torch.equal(on[2][:2], on[3])                          # 64 M tiles vs 16: same bits per sample
    assert rel(on[0], F.linear(x, w, b) + r) < 4e-3
    assert rel(on[2], F.conv2d(xc, wc, bc, padding=1)) < 4e-3


def test_paste_back_bit_exact(engine, oracle):
    """rfb_paste_back vs the Pillow-written fixture (bit exact) and vs the oracle at the real sizes
    (512 -> 1024 resize, 720p frame, two frames with different coefficients)."""
    import os
    import numpy as np
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "paste_64.npz"))
    out = engine.paste_back(torch.from_numpy(g["x01"])[None], torch.from_numpy(g["orig"])[None], g["coeffs"][None],
                            up=int(g["up"]))
    assert np.array_equal(out[0].cpu().numpy(), g["pasted"])
    gen = torch.Generator().manual_seed(21)
    x01 = torch.rand(2, 3, 512, 512, generator=gen)
    orig = torch.randint(0, 256, (2, 720, 1280, 3), generator=gen, dtype=torch.uint8)
    co = np.array([[1.45, 0.05, -310.0, -0.04, 1.5, -95.0, 2e-5, -1e-5], [1.2, -0.1, -150.0, 0.08, 1.25, -60.0, -3e-5, 4e-5]])
    out = engine.paste_back(x01, orig, co, up=1024).cpu().numpy()
    for i in range(2):
        ref = oracle.paste_back(x01[i].numpy(), orig[i].numpy(), co[i], up=1024)
        assert np.array_equal(out[i], ref), i
