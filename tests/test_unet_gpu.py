"""GPU: the UNet forward and the DDIM loop through the C ABI against the oracle / the reference goldens.

Stated tolerance (fp16 operands, fp32 accumulation, fp32 statistics/latents): max-abs error of eps
<= 2e-2 * max|eps_ref| per UNet call."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _g(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, name + ".npz")).items()}


@pytest.mark.parametrize("name", ["unet_L16", "unet_L32"])
def test_unet_forward_vs_reference_golden(unet_engine, name):
    g = _g(name)
    eps = unet_engine.unet_forward(g["x"], g["t"], g["ctx"]).cpu()
    err = float((eps - g["eps"]).abs().max()) / float(g["eps"].abs().max())
    print(name, "rel err", err)
    assert err < TOL, err


def test_unet_forward_L64_vs_oracle(unet_engine, oracle, unet_sd):
    gen = torch.Generator().manual_seed(11)
    x, t = torch.randn(2, 9, 64, 64, generator=gen), torch.tensor([981, 21])
    ctx = torch.randn(2, 1, 768, generator=gen)
    with torch.no_grad():
        ref = oracle.unet_forward(oracle.Params(unet_sd, oracle.PFX_UNET), x, t, ctx)
    eps = unet_engine.unet_forward(x, t, ctx).cpu()
    err = float((eps - ref).abs().max()) / float(ref.abs().max())
    print("L64 rel err", err)
    assert err < TOL, err


def test_unet_forward_multi_token_context(unet_engine, oracle, unet_sd):
    """Context length T = 3 (the reference's stack_feat conditioning, ddpm.py:1027-1030): the general
    CrossAttention path (attention.py:179-221) instead of the single-token shortcut."""
    gen = torch.Generator().manual_seed(13)
    x, t = torch.randn(2, 9, 16, 16, generator=gen), torch.tensor([601, 1])
    ctx = torch.randn(2, 3, 768, generator=gen)
    with torch.no_grad():
        ref = oracle.unet_forward(oracle.Params(unet_sd, oracle.PFX_UNET), x, t, ctx)
    eps = unet_engine.unet_forward(x, t, ctx).cpu()
    err = float((eps - ref).abs().max()) / float(ref.abs().max())
    print("T=3 rel err", err)
    assert err < TOL, err


def test_unet_forward_L128_fused_vs_materialised_attention(unet_engine):
    """BASELINE configs[3] (1024x1024 images, latent 128x128, 16384 tokens at the first level): too large for the CPU oracle
    in a unit test, so two independent CUDA implementations of the attention are checked against each other --
    the fused tcgen05 kernels (v4 at d=40, v3 at d=80) and the materialised S / softmax / P.V path."""
    gen = torch.Generator().manual_seed(14)
    x, t = torch.randn(1, 9, 128, 128, generator=gen), torch.tensor([401])
    ctx = torch.randn(1, 1, 768, generator=gen)
    a = unet_engine.unet_forward(x, t, ctx).cpu()
    unet_engine.set_option("attn_flash", 0)
    try:
        b = unet_engine.unet_forward(x, t, ctx).cpu()
    finally:
        unet_engine.set_option("attn_flash", 4)
    assert torch.isfinite(a).all()
    err = float((a - b).abs().max()) / float(b.abs().max())
    print("L128 fused vs materialised rel diff", err)
    assert err < TOL, err


def test_batch_independence(unet_engine):
    """No cross-sample arithmetic: a batch of 4 equals the per-sample results bit for bit (multi-GPU invariant)."""
    gen = torch.Generator().manual_seed(12)
    x, t = torch.randn(4, 9, 16, 16, generator=gen), torch.tensor([981, 21, 401, 401])
    ctx = torch.randn(4, 1, 768, generator=gen)
    full = unet_engine.unet_forward(x, t, ctx)
    for i in range(0, 4, 2):
        part = unet_engine.unet_forward(x[i:i + 2], t[i:i + 2], ctx[i:i + 2])
        assert torch.equal(part, full[i:i + 2])


def test_ddim_loop_vs_reference_golden(unet_engine):
    g = _g("ddim_S5_L16")
    x0, ix, ip = unet_engine.ddim_sample(g["x_T"], g["z"], g["mask"], g["c"], g["uc"], S=5, scale=3.5, log_every_t=2)
    err = float((x0.cpu() - g["x0"]).abs().max()) / float(g["x0"].abs().max())
    print("ddim rel err", err)
    assert err < 5e-2, err
    assert ix.shape[0] == int(g["n_inter"]) - 1      # the reference list also holds x_T as its first entry
    perr = float((ip[-1].cpu() - g["pred_x0_last"]).abs().max()) / float(g["x0"].abs().max())
    assert perr < 5e-2, perr


def test_sampling_loop_cuda_graph_replays_are_bit_exact(unet_engine):
    """rfb_ddim_sample captures the whole S-step loop (every kernel of every step with its own schedule scalars) into ONE
    CUDA graph the second time it sees a (shape, schedule, scale) and replays it afterwards.  Replays must be
    the bits of the eager loop -- also for NEW inputs (inputs are staged into arena buffers the graph reads) -- and a
    different schedule / option must not hit a stale graph."""
    g = _g("ddim_S5_L16")
    eng = unet_engine
    args = (g["x_T"], g["z"], g["mask"], g["c"], g["uc"])
    eng.set_option("use_graph", 0)
    eager = [t.clone() for t in eng.ddim_sample(*args, S=5, scale=3.5, log_every_t=2)]
    x2 = g["x_T"] * 0.5 + 0.1
    eager2 = eng.ddim_sample(x2, g["z"], g["mask"], g["c"], g["uc"], S=5, scale=3.5, log_every_t=2)[0].clone()
    eager_s6 = eng.ddim_sample(*args, S=6, scale=3.0, log_every_t=2)[0].clone()
    eng.set_option("use_graph", 1)
    r0 = eng.graph_replays
    for i in range(3):                                   # 1st eager, 2nd capture + launch, 3rd replay
        out = eng.ddim_sample(*args, S=5, scale=3.5, log_every_t=2)
        for a, b in zip(out, eager):
            assert torch.equal(a, b), i
    assert eng.graph_replays - r0 == 2
    l0 = eng.launch_count
    assert torch.equal(eng.ddim_sample(x2, g["z"], g["mask"], g["c"], g["uc"], S=5, scale=3.5, log_every_t=2)[0], eager2)
    assert eng.graph_replays - r0 == 3 and eng.launch_count - l0 > 1000      # replays count their kernels
    assert torch.equal(eng.ddim_sample(*args, S=6, scale=3.0, log_every_t=2)[0], eager_s6)      # new key: eager again
    assert eng.graph_replays - r0 == 3
    eng.set_option("cfg_share", 0)                       # any option change drops the cached graphs
    try:
        assert torch.equal(eng.ddim_sample(*args, S=5, scale=3.5, log_every_t=2)[0], eager[0])
        assert eng.graph_replays - r0 == 3
    finally:
        eng.set_option("cfg_share", 1)


def test_programmatic_dependent_launch_is_bit_exact(unet_engine):
    """Option "pdl" (1, the default: eager launches; 2: inside captured graphs as well): the kernels of the loop are launched with the programmatic-stream-serialization
    attribute, so a kernel's CTAs are scheduled (and run their prologue: barrier init, TMEM allocation, descriptor
    prefetch) while the previous kernel drains, and wait in `griddepcontrol.wait` before touching memory.  A missing wait
    would be a race between consecutive kernels: the eager loop, the captured graph and repeated replays must all give
    the bits of the fully serialised launches (pdl=0), at a shape that fills the GPU (L=64, N=2*4) and at a tiny one."""
    eng = unet_engine
    g = _g("ddim_S5_L16")
    small = (g["x_T"], g["z"], g["mask"], g["c"], g["uc"])
    gen = torch.Generator(device="cpu").manual_seed(5)
    B, L = 4, 64
    big = tuple(t.cuda() for t in (torch.randn(B, 4, L, L, generator=gen), torch.randn(B, 4, L, L, generator=gen),
                                   (torch.rand(B, 1, L, L, generator=gen) > 0.5).float(),
                                   torch.randn(B, 1, 768, generator=gen), torch.randn(B, 1, 768, generator=gen)))
    for args, S in ((small, 5), (big, 4)):
        eng.set_option("pdl", 0)
        eng.set_option("use_graph", 0)
        try:
            ref = eng.ddim_sample(*args, S=S, scale=3.5, log_every_t=2)[0].clone()
            eng.set_option("pdl", 1)
            assert torch.equal(eng.ddim_sample(*args, S=S, scale=3.5, log_every_t=2)[0], ref)
            eng.set_option("pdl", 2)
            eng.set_option("use_graph", 1)
            r0 = eng.graph_replays
            for i in range(4):
                assert torch.equal(eng.ddim_sample(*args, S=S, scale=3.5, log_every_t=2)[0], ref), i
            assert eng.graph_replays - r0 == 3
        finally:
            eng.set_option("pdl", 1)
            eng.set_option("use_graph", 1)


def test_fused_output_conv_cfg_ddim_update_equals_separate_kernels(unet_engine, oracle):
    """Inside the sampler the UNet's output convolution (a [N*L*L,320]x[320,36] tap GEMM) is finished by the update
    kernel itself: 9-tap gather + bias -> CFG combine -> x_{t-1} / pred_x0 (ddim.py:346,363-374), eps never stored.
    One step must give the bits of UNetModel.forward -> rfb_cfg_ddim_update (itself bit-exact vs the oracle)."""
    g = _g("ddim_S5_L16")
    eng = unet_engine
    sch = oracle.ddim_schedule(5)
    _, ix, ip = eng.ddim_sample(g["x_T"], g["z"], g["mask"], g["c"], g["uc"], S=5, scale=3.5, log_every_t=1, n_steps_limit=1)
    x9 = eng.concat9(g["x_T"], g["z"], g["mask"], dup=2)
    t = torch.full((2,), int(sch["timesteps"][4]), dtype=torch.long)
    eps2 = eng.unet_forward(x9, t, torch.cat([g["uc"], g["c"]]))
    xp, p0 = eng.cfg_ddim_update(g["x_T"], eps2, 3.5, *[sch[k][4] for k in ("a_t", "a_prev", "sigma", "sqrt_one_minus_a")])
    assert torch.equal(ix[0], xp) and torch.equal(ip[0], p0)


def test_cfg_head_sharing_is_bit_exact(unet_engine):
    """Inside the samplers the two CFG halves share their input and timestep (ddim.py:338-344): conv_in, the first
    ResBlock and attn1 of the first SpatialTransformer are computed once per pair (option cfg_share).  The results
    must be the bits of the unshared evaluation."""
    g = _g("ddim_S5_L16")
    outs = []
    for share in (1, 0):
        unet_engine.set_option("cfg_share", share)
        try:
            outs.append(unet_engine.ddim_sample(g["x_T"], g["z"], g["mask"], g["c"], g["uc"], S=5, scale=3.5)[0].cpu())
        finally:
            unet_engine.set_option("cfg_share", 1)
    assert torch.equal(outs[0], outs[1])


def test_plms_loop_vs_reference_golden(unet_engine):
    """rfb_plms_sample against the reference PLMSSampler's output (tests/golden/plms_S6_L16.npz; 7 steps, 8 UNet calls)."""
    g = _g("plms_S6_L16")
    x0, ix, ip = unet_engine.plms_sample(g["x_T"], g["z"], g["mask"], g["c"], g["uc"], S=6, scale=3.5, log_every_t=2)
    err = float((x0.cpu() - g["x0"]).abs().max()) / float(g["x0"].abs().max())
    print("plms rel err", err)
    assert err < 5e-2, err
    assert ix.shape[0] == int(g["n_inter"]) - 1
    perr = float((ip[-1].cpu() - g["pred_x0_last"]).abs().max()) / float(g["x0"].abs().max())
    assert perr < 5e-2, perr


def test_q_sample_bit_exact(unet_engine):
    g = _g("plms_S6_L16")
    out = unet_engine.q_sample(g["z"], g["q_t"], g["q_noise"])
    assert torch.equal(out.cpu(), g["q_out"])        # two fp32 multiplies and one add: bit exact
