"""GPU: the drop-in classes of reface_b200.ldm_api behave like the reference slots they replace
(ldm/util.py:78-93 instantiate_from_config targets; ddim.py:142-198 sampler contract)."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def instantiate_from_config(config):          # the reference's factory, restated (ldm/util.py:78-93)
    module, cls = config["target"].rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)(**config.get("params", dict()))


def test_unet_slot_via_instantiate_from_config(engine, oracle, unet_sd):
    cfg = {"target": "reface_b200.ldm_api.UNetModel",
           "params": dict(image_size=32, in_channels=9, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                          num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                          transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False,
                          add_conv_in_front_of_unet=False, engine=engine)}
    m = instantiate_from_config(cfg)
    m.load_state_dict({k[len(oracle.PFX_UNET):]: v for k, v in unet_sd.items()})
    g = torch.Generator().manual_seed(3)
    x, t, ctx = torch.randn(2, 9, 16, 16, generator=g), torch.tensor([981, 1]), torch.randn(2, 1, 768, generator=g)
    with torch.no_grad():
        ref = oracle.unet_forward(oracle.Params(unet_sd, oracle.PFX_UNET), x, t, ctx)
    out = m(x.cuda(), t.cuda(), context=ctx.cuda()).cpu()
    assert float((out - ref).abs().max() / ref.abs().max()) < 2e-2
    with pytest.raises(NotImplementedError):
        instantiate_from_config({"target": "reface_b200.ldm_api.UNetModel", "params": dict(model_channels=192, engine=engine)})


def test_autoencoder_slot(engine, oracle, vae_sd):
    from reface_b200.ldm_api import AutoencoderKL
    m = AutoencoderKL(ddconfig={}, lossconfig={}, embed_dim=4, engine=engine)
    m.load_state_dict({k[len(oracle.PFX_VAE):]: v for k, v in vae_sd.items()})
    g = torch.Generator().manual_seed(4)
    x = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    noise = torch.randn(1, 4, 8, 8, generator=g)
    P = oracle.Params(vae_sd, oracle.PFX_VAE)
    with torch.no_grad():
        mean, logvar = oracle.vae_encode_moments(P, x)
    post = m.encode(x.cuda())
    assert float((post.mode().cpu() - mean).abs().max() / mean.abs().max()) < 2e-2
    z = post.sample(noise.cuda()).cpu()                         # unscaled, like DiagonalGaussianDistribution.sample()
    zr = mean + torch.exp(0.5 * logvar) * noise
    assert float((z - zr).abs().max() / zr.abs().max()) < 2e-2
    img = m.decode(zr.cuda()).cpu()
    with torch.no_grad():
        ref = oracle.vae_decode(P, oracle.SCALE_FACTOR * zr)
    assert float((img - ref).abs().max() / ref.abs().max()) < 2e-2


def test_sampler_contract(unet_engine, oracle, unet_sd):
    """Same kwargs contract as ddim.py:323-334: raises without test_model_kwargs / rest; returns (samples, dict of lists)."""
    from reface_b200.ldm_api import DDIMSampler, LatentDiffusion
    model = LatentDiffusion(engine=unet_engine)
    s = DDIMSampler(model)
    g = torch.Generator().manual_seed(5)
    B, L = 1, 16
    c, uc = torch.randn(B, 1, 768, generator=g).cuda(), torch.randn(B, 1, 768, generator=g).cuda()
    z, mask = torch.randn(B, 4, L, L, generator=g).cuda(), torch.rand(B, 1, L, L, generator=g).cuda()
    x_T = torch.randn(B, 4, L, L, generator=g).cuda()
    with pytest.raises(Exception):
        s.sample(S=5, batch_size=B, shape=[4, L, L], conditioning=c, verbose=False)
    smp, inter = s.sample(S=5, batch_size=B, shape=[4, L, L], conditioning=c, verbose=False, x_T=x_T, log_every_t=2,
                          unconditional_guidance_scale=3.5, unconditional_conditioning=uc,
                          test_model_kwargs={"inpaint_image": z, "inpaint_mask": mask})
    assert smp.shape == (B, 4, L, L) and len(inter["x_inter"]) == len(inter["pred_x0"]) == 4
    ref, rint = oracle.ddim_sample(oracle.Params(unet_sd, oracle.PFX_UNET), x_T.cpu(), z.cpu(), mask.cpu(), c.cpu(), uc.cpu(),
                                   5, 3.5, log_every_t=2)
    assert len(rint["x_inter"]) == 4
    assert float((smp.cpu() - ref).abs().max() / ref.abs().max()) < 5e-2
    # 'rest' form of the kwargs (ddim.py:331-332) and scale == 1 (no CFG duplication, ddim.py:335-336)
    smp2, _ = s.sample(S=5, batch_size=B, shape=[4, L, L], conditioning=c, verbose=False, x_T=x_T,
                       unconditional_guidance_scale=1.0, rest=torch.cat([z, mask], 1))
    ref2, _ = oracle.ddim_sample(oracle.Params(unet_sd, oracle.PFX_UNET), x_T.cpu(), z.cpu(), mask.cpu(), c.cpu(), c.cpu(),
                                 5, 1.0)
    assert float((smp2.cpu() - ref2).abs().max() / ref2.abs().max()) < 5e-2


def test_ddim_update_with_noise_bit_exact(engine, oracle):
    gen = torch.Generator().manual_seed(8)
    x, eps2, nz = torch.randn(2, 4, 8, 8, generator=gen), torch.randn(4, 4, 8, 8, generator=gen), torch.randn(2, 4, 8, 8, generator=gen)
    sch = oracle.ddim_schedule(50, eta=0.7)
    f32 = np.float32
    for idx in (49, 7):
        a = [sch[k][idx] for k in ("a_t", "a_prev", "sigma", "sqrt_one_minus_a")]
        assert a[2] > 0
        xp, p0 = engine.cfg_ddim_update(x, eps2, 3.5, *a, noise=nz)
        # (1) bit-exact against an IEEE-754 fp32 evaluation of ddim.py:346,363-374 in the reference's op order
        #     (numpy fp32: every +,-,*,/ and sqrt correctly rounded -- what torch does on CUDA)
        a_t, a_prev, sigma, s1m = (f32(v) for v in a)
        eu, ec = eps2[:2].numpy(), eps2[2:].numpy()
        e = eu + f32(3.5) * (ec - eu)
        ip0 = (x.numpy() - s1m * e) / np.sqrt(a_t)
        ixp = (np.sqrt(a_prev) * ip0 + np.sqrt((f32(1.0) - a_prev) - sigma * sigma) * e) + sigma * nz.numpy()
        assert np.array_equal(p0.cpu().numpy(), ip0) and np.array_equal(xp.cpu().numpy(), ixp)
        # (2) within 1 ulp of the torch-CPU oracle: this torch build's vectorised CPU sqrt is 1 ulp off the correctly
        #     rounded value for some schedule entries (e.g. sqrt(a_prev[7])), the CUDA kernel and numpy are not
        rxp, rp0, _ = oracle.cfg_ddim_update(x, eps2[:2], eps2[2:], 3.5, *a, noise=nz)
        #     (absolute bound: one ulp of the largest term, |x| < 4 -> 2.4e-7; a relative bound would blow up on the
        #     elements where the terms cancel)
        assert float((xp.cpu() - rxp).abs().max()) < 5e-7
        assert float((p0.cpu() - rp0).abs().max()) < 5e-7 * float(rp0.abs().max().clamp_min(1.0))
    from reface_b200.runtime import ddim_schedule
    s2 = ddim_schedule(50, 0.7)
    assert np.array_equal(s2["sigma"], sch["sigma"])


def test_empty_and_single_batches(engine, oracle, unet_sd, vae_sd, clip_sd, arc_sd, fusion_sd):
    """Edge cases of the batch dimension: an EMPTY batch flows through the whole drop-in path like it does through the
    reference's torch modules (empty tensors out, no kernel launched on zero rows); a batch of one equals row 0 of a batch
    of three bit for bit (ragged last chunks of a video, shards of unequal size)."""
    from reface_b200.ldm_api import LatentDiffusion, swap_faces
    sd = {**unet_sd, **vae_sd, **clip_sd, **arc_sd, **fusion_sd}
    model = LatentDiffusion(sd, engine=engine)
    inp = {k: v.cuda() for k, v in oracle.synthetic_inputs(3, 128, seed=9).items()}
    empty = {k: v[:0] for k, v in inp.items()}
    out = swap_faces(model, S=4, scale=3.5, **empty)
    assert out["image"].shape == (0, 3, 128, 128) and out["samples"].shape == (0, 4, 16, 16) and out["c"].shape == (0, 1, 768)
    full = swap_faces(model, S=4, scale=3.5, **inp)
    one = swap_faces(model, S=4, scale=3.5, **{k: v[:1] for k, v in inp.items()})
    assert torch.equal(one["image"], full["image"][:1]) and torch.equal(one["samples"], full["samples"][:1])
    assert torch.isfinite(full["image"]).all()
