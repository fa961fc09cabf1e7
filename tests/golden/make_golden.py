"""Pins oracle/reface_oracle.py against the REAL reference modules and writes golden vectors.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

For every component it (1) builds the reference nn.Module from /root/reference (with the import
shims of oracle/ref_shims for packages missing in this image), (2) loads the oracle's seeded
state dict with strict=True -- which proves the oracle's parameter spec equals the reference's --
(3) runs both on the same seeded inputs, asserts agreement (fp32 round-off only) and (4) stores the
REFERENCE outputs in tests/golden/<name>.npz.  tests/test_oracle.py re-checks the oracle against
these files everywhere (no /root/reference needed).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

import reface_oracle as O  # noqa: E402

torch.set_grad_enabled(False)
SEED = 0


def sub_sd(spec, prefix, seed=SEED):
    sd = O.init_state_dict(spec, seed)
    return sd, {k[len(prefix):]: v for k, v in sd.items()}


def check(name, ref, ora, tol):
    err = (ref - ora).abs().max().item()
    mag = ref.abs().max().item()
    print(f"  {name}: max|ref-oracle|={err:.3e}  max|ref|={mag:.3e}")
    assert err <= tol * max(1.0, mag), (name, err, mag)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"  wrote {path} ({os.path.getsize(path)/1024:.1f} KiB)")


def golden_schedule():
    print("schedule")
    from ldm.modules.diffusionmodules.util import make_beta_schedule, make_ddim_timesteps, \
        make_ddim_sampling_parameters, timestep_embedding
    betas = make_beta_schedule("linear", 1000, linear_start=O.LINEAR_START, linear_end=O.LINEAR_END)
    ac = torch.tensor(np.cumprod(1.0 - betas, axis=0), dtype=torch.float32)
    out = {}
    for S in (5, 30, 50):
        ts = make_ddim_timesteps("uniform", S, 1000, verbose=False)
        sig, a, ap = make_ddim_sampling_parameters(ac.cpu(), ts, 0.0, verbose=False)
        sch = O.ddim_schedule(S, 0.0)
        assert np.array_equal(ts, sch["timesteps"])
        assert np.array_equal(a.numpy(), sch["a_t"]) and np.array_equal(np.float32(ap), sch["a_prev"])
        assert np.array_equal(np.sqrt(1.0 - a).numpy(), sch["sqrt_one_minus_a"])
        out[f"ts{S}"], out[f"a{S}"], out[f"ap{S}"] = ts, a.numpy(), np.float32(ap)
    t = torch.tensor([981, 1, 500])
    te = timestep_embedding(t, 320)
    assert torch.equal(te, O.timestep_embedding(t, 320))
    save("schedule", alphas_cumprod=ac, temb=te, temb_t=t, **out)


def golden_unet():
    print("unet")
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    cfg = O.UNET_CFG
    m = UNetModel(image_size=32, in_channels=9, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                  num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                  transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False,
                  add_conv_in_front_of_unet=False).eval()
    sd, sub = sub_sd(O.unet_spec(), O.PFX_UNET)
    m.load_state_dict(sub, strict=True)
    g = torch.Generator().manual_seed(1)
    for L, N in ((16, 2), (32, 2)):
        x = torch.randn(N, 9, L, L, generator=g)
        t = torch.tensor([981, 401][:N])
        ctx = torch.randn(N, 1, 768, generator=g)
        ref = m(x, t, context=ctx)
        taps = {}
        ora = O.unet_forward(O.Params(sd, O.PFX_UNET), x, t, ctx, cfg, taps=taps)
        check(f"unet L={L}", ref, ora, 2e-5)
        save(f"unet_L{L}", x=x, t=t, ctx=ctx, eps=ref,
             **{"tapmean_" + k.replace(".", "_"): v.mean() for k, v in taps.items()},
             **{"tapstd_" + k.replace(".", "_"): v.std() for k, v in taps.items()})
    return m, sd


def golden_ddim(unet, sd):
    print("ddim (reference DDIMSampler driving the reference UNet)")
    from ldm.models.diffusion.ddim import DDIMSampler
    DDIMSampler.register_buffer = lambda self, n, a: setattr(self, n, a)   # ddim.py:104-108 hard-codes cuda
    ac = O.alphas_cumprod_f32()

    class FakeLD:   # the attributes DDIMSampler touches (ddim.py:100,113-119,207,345)
        num_timesteps = 1000
        betas = torch.tensor(O.make_beta_schedule(), dtype=torch.float32)
        alphas_cumprod = ac
        alphas_cumprod_prev = torch.tensor(np.append(1.0, ac.double().numpy()[:-1]), dtype=torch.float32)
        device = torch.device("cpu")

        def apply_model(self, x, t, c):          # ddpm.py:1519-1617 + DiffusionWrapper :2244-2246
            return unet(x, t, context=torch.cat([c], 1))

    g = torch.Generator().manual_seed(2)
    B, L, S, scale = 1, 16, 5, 3.5
    x_T = torch.randn(B, 4, L, L, generator=g)
    z = torch.randn(B, 4, L, L, generator=g)
    mask = (torch.rand(B, 1, L, L, generator=g) > 0.5).float()
    c = torch.randn(B, 1, 768, generator=g)
    uc = torch.randn(B, 1, 768, generator=g)
    smp = DDIMSampler(FakeLD())
    ref, inter = smp.sample(S=S, conditioning=c, batch_size=B, shape=[4, L, L], verbose=False,
                            unconditional_guidance_scale=scale, unconditional_conditioning=uc, eta=0.0, x_T=x_T,
                            log_every_t=2, test_model_kwargs={"inpaint_image": z, "inpaint_mask": mask})
    ora, ointer = O.ddim_sample(O.Params(sd, O.PFX_UNET), x_T, z, mask, c, uc, S, scale, log_every_t=2)
    check("ddim x0", ref, ora, 5e-5)
    assert len(inter["x_inter"]) == len(ointer["x_inter"])
    save("ddim_S5_L16", x_T=x_T, z=z, mask=mask, c=c, uc=uc, x0=ref, n_inter=len(inter["x_inter"]),
         pred_x0_last=inter["pred_x0"][-1])
    # the concat must be a pure copy (bit exact)
    assert torch.equal(O.concat9(x_T, z, mask), torch.cat([x_T, z, mask], 1))


def golden_plms(unet, sd):
    print("plms (reference PLMSSampler driving the reference UNet) + q_sample")
    from ldm.models.diffusion.plms import PLMSSampler
    PLMSSampler.register_buffer = lambda self, n, a: setattr(self, n, a)   # plms.py:18-22 hard-codes cuda
    ac = O.alphas_cumprod_f32()

    class FakeLD:   # the attributes PLMSSampler touches (plms.py:15,27-35,123,183-187)
        num_timesteps = 1000
        betas = torch.tensor(O.make_beta_schedule(), dtype=torch.float32)
        alphas_cumprod = ac
        alphas_cumprod_prev = torch.tensor(np.append(1.0, ac.double().numpy()[:-1]), dtype=torch.float32)
        device = torch.device("cpu")

        def apply_model(self, x, t, c):
            return unet(x, t, context=torch.cat([c], 1))

    g = torch.Generator().manual_seed(3)
    B, L, S, scale = 1, 16, 6, 3.5
    x_T = torch.randn(B, 4, L, L, generator=g)
    z = torch.randn(B, 4, L, L, generator=g)
    mask = (torch.rand(B, 1, L, L, generator=g) > 0.5).float()
    c = torch.randn(B, 1, 768, generator=g)
    uc = torch.randn(B, 1, 768, generator=g)
    smp = PLMSSampler(FakeLD())
    ref, inter = smp.sample(S=S, conditioning=c, batch_size=B, shape=[4, L, L], verbose=False,
                            unconditional_guidance_scale=scale, unconditional_conditioning=uc, eta=0.0, x_T=x_T,
                            log_every_t=2, test_model_kwargs={"inpaint_image": z, "inpaint_mask": mask})
    ora, ointer = O.plms_sample(O.Params(sd, O.PFX_UNET), x_T, z, mask, c, uc, S, scale, log_every_t=2)
    check("plms x0", ref, ora, 5e-5)
    assert len(inter["x_inter"]) == len(ointer["x_inter"])
    # q_sample (ddpm.py:412-415) through the reference's own buffers (register_schedule, ddpm.py:255-285)
    import ldm.models.diffusion.ddpm as ddpm

    class Sched(ddpm.DDPM):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.v_posterior, self.parameterization = 0.0, "eps"
            self.register_schedule(beta_schedule="linear", timesteps=1000, linear_start=0.00085, linear_end=0.012)

    sch = Sched()
    t = torch.tensor([999])
    nz = torch.randn(B, 4, L, L, generator=g)
    qs = sch.q_sample(x_start=z, t=t, noise=nz)
    assert torch.equal(qs, O.q_sample(z, t, nz)), "q_sample must be bit exact (two fp32 multiplies and one add)"
    save("plms_S6_L16", x_T=x_T, z=z, mask=mask, c=c, uc=uc, x0=ref, n_inter=len(inter["x_inter"]),
         pred_x0_last=inter["pred_x0"][-1], q_t=t, q_noise=nz, q_out=qs)


def golden_vae():
    print("vae")
    from ldm.models.autoencoder import AutoencoderKL
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    m = AutoencoderKL(ddconfig=dd, lossconfig={"target": "torch.nn.Identity"}, embed_dim=4).eval()
    sd, sub = sub_sd(O.vae_spec(), O.PFX_VAE)
    m.load_state_dict(sub, strict=True)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    noise = torch.randn(1, 4, 8, 8, generator=g)
    post = m.encode(x)
    ref_z = O.SCALE_FACTOR * (post.mean + post.std * noise)      # distributions.py:35-37 with explicit noise
    P = O.Params(sd, O.PFX_VAE)
    mean, logvar = O.vae_encode_moments(P, x)
    check("vae mean", post.mean, mean, 2e-5)
    check("vae logvar", post.logvar, logvar, 2e-5)
    ora_z = O.vae_encode(P, x, noise)
    check("vae z", ref_z, ora_z, 2e-5)
    zz = torch.randn(1, 4, 8, 8, generator=g) * 0.18215 * 4
    ref_img = m.decode((1.0 / O.SCALE_FACTOR) * zz)
    check("vae decode", ref_img, O.vae_decode(P, zz), 2e-5)
    save("vae_64", x=x, noise=noise, mean=post.mean, logvar=post.logvar, z=ref_z, zdec=zz, img=ref_img)


def _build_clip_embedder():
    import transformers
    from transformers import CLIPConfig, CLIPModel
    conf = CLIPConfig(vision_config=dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24,
                                         num_attention_heads=16, image_size=224, patch_size=14, hidden_act="quick_gelu"),
                      text_config=dict(hidden_size=64, intermediate_size=64, num_hidden_layers=1, num_attention_heads=1),
                      projection_dim=768)
    transformers.CLIPModel.from_pretrained = classmethod(lambda cls, *a, **k: CLIPModel(conf))
    transformers.CLIPTokenizer.from_pretrained = classmethod(lambda cls, *a, **k: None)
    import ldm.modules.encoders.modules as em
    em.CLIPModel, em.CLIPTokenizer = transformers.CLIPModel, transformers.CLIPTokenizer
    return em.FrozenCLIPEmbedder().eval()


def golden_clip():
    print("clip (+mapper2/final_ln2)")
    m = _build_clip_embedder()
    sd, sub = sub_sd(O.clip_spec(), O.PFX_CLIP)
    missing, unexpected = m.load_state_dict(sub, strict=False)
    # the embedder also constructs the unused text tower / mapper / final_ln / projection_back
    # (encoders/modules.py:215-233): they must be the ONLY keys our spec lacks.
    assert not unexpected, unexpected
    for k in missing:
        assert k.startswith(("model.text_model", "model.text_projection", "model.logit_scale", "mapper.", "final_ln.",
                             "projection_back.", "model.vision_model.embeddings.position_ids")), k
    g = torch.Generator().manual_seed(4)
    img = torch.randn(1, 3, 224, 224, generator=g)
    ref = m.encode(img)
    ora = O.clip_embed(O.Params(sd, O.PFX_CLIP), img)
    check("clip", ref, ora, 5e-5)
    save("clip_B1", img_seed=4, out=ref)
    return m, sd


def golden_arcface_and_fusion(clip_mod, clip_sd):
    print("arcface + conditioning fusion")
    from src.Face_models.encoders.model_irse import Backbone
    import ldm.models.diffusion.ddpm as ddpm
    bb = Backbone(input_size=112, num_layers=50, drop_ratio=0.6, mode="ir_se").eval()
    sd, sub = sub_sd(O.arcface_spec(), O.PFX_ARC)
    missing, unexpected = bb.load_state_dict(sub, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(5)
    ref_img = torch.randn(2, 3, 224, 224, generator=g)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "arc.pth")
        torch.save(bb.state_dict(), path)
        opts = types.SimpleNamespace(other_params=types.SimpleNamespace(arcface_path=path))
        idl = ddpm.IDLoss(opts).eval()
    ref = idl.extract_feats(ref_img)[0]
    ora = O.arcface_embed(O.Params(sd, O.PFX_ARC), ref_img)
    check("arcface", ref, ora, 5e-5)

    fsd = O.init_state_dict(O.fusion_spec(), SEED)
    lin = lambda n, i, o: _mk_linear(fsd, n, i, o)
    tar = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    lm_raw = torch.zeros(2, 136)
    fake = types.SimpleNamespace(
        training=False, update_weight=False, clip_weight=1.0, ID_weight=10.0, Landmarks_weight=0.05,
        Source_CLIP_feat=True, Target_CLIP_feat=True, use_3dmm=False, normalize=False, Landmark_cond=True,
        weight_division=True, concat_feat=False, stack_feat=False, land_mark_id_seperate_layers=False,
        sep_head_att=False, device=torch.device("cpu"),
        get_learned_conditioning=lambda x: clip_mod.encode(x), face_ID_model=idl,
        proj_out_source=lin("proj_out_source", 768, 768), proj_out_target=lin("proj_out_target", 768, 768),
        ID_proj_out=lin("ID_proj_out", 512, 768))
    # torchvision>=0.17 defaults to antialias=True on tensors; the pinned 0.14 does not
    import torchvision.transforms.functional as TF
    _resize = TF.resize
    ddpm.TF.resize = lambda img, size, *a, **k: _resize(img, list(size), antialias=False)
    lm = lin("landmark_proj_out", 136, 768)(lm_raw)                     # ddpm.py:1096
    ref_c = ddpm.LatentDiffusion.conditioning_with_feat(fake, ref_img, lm, tar=tar)
    ddpm.TF.resize = _resize
    full = dict(clip_sd); full.update(sd); full.update(fsd)
    ora_c = O.conditioning_with_feat(O.Params(full), ref_img, tar, lm_raw)
    check("conditioning", ref_c, ora_c, 5e-5)
    save("cond_B2", ref_seed=5, id_feat=ref, c=ref_c, tar=tar)


def _mk_linear(sd, name, i, o):
    l = torch.nn.Linear(i, o)
    l.weight.copy_(sd[name + ".weight"]); l.bias.copy_(sd[name + ".bias"])
    return l


def golden_parse():
    print("face parsing (reference BiSeNet + 19->12 label conversion + mask preparation)")
    _cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self       # model.py:15-16 move two constants to cuda at import time
    import torch.utils.model_zoo as mz
    mz.load_url = lambda *a, **k: {}                     # resnet.py:83 downloads ImageNet weights in __init__
    try:
        import pretrained.face_parsing.face_parsing_demo as fpd
        from pretrained.face_parsing.model import BiSeNet
    finally:
        torch.Tensor.cuda = _cuda
    net = BiSeNet(n_classes=19).eval()
    sd, sub = sub_sd(O.parse_spec(), O.PFX_PARSE)
    missing, unexpected = net.load_state_dict(sub, strict=False)
    assert not unexpected, unexpected
    for k in missing:   # the auxiliary training heads and BN step counters are the only keys the oracle does not hold
        assert k.startswith(("conv_out16.", "conv_out32.")) or k.endswith("num_batches_tracked"), k
    g = torch.Generator().manual_seed(6)
    H = 256
    img01 = torch.rand(1, 3, H, H, generator=g)
    mean = torch.tensor(O.SEG_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(O.SEG_STD).view(1, 3, 1, 1)
    im = (img01.clamp(0, 1) - mean) / std                # FaceParser.preprocess_img, face_parsing_demo.py:266-268
    ref_logits = net(im)[0]
    ora_logits = O.bisenet_logits(O.Params(sd, O.PFX_PARSE), im)
    check("bisenet logits", ref_logits, ora_logits, 2e-5)
    seg19 = torch.argmax(ref_logits, dim=1)
    o19, o12 = O.face_parse(O.Params(sd, O.PFX_PARSE), img01)
    assert torch.equal(o19, seg19)
    conv = getattr(fpd, "__ffhq_masks_to_faceParser_mask_detailed")
    seg12 = torch.from_numpy(conv(seg19[0].numpy().astype(np.uint8)).astype(np.int64))[None]
    assert torch.equal(o12, seg12)
    # mask / inpaint image exactly as ldm/data/video_swap_dataset.py:150-222 builds them (numpy isin + 1 - mask)
    img_m11 = img01 * 2 - 1
    keep = np.isin(seg12[0].numpy(), list(O.REMOVE_MASK_TAR_FFHQ))
    mask_ref = torch.from_numpy(1.0 - keep.astype(np.float32))[None, None]
    m, inp = O.inpaint_from_parsing(img_m11, seg12)
    assert torch.equal(m, mask_ref) and torch.equal(inp, img_m11 * mask_ref)
    top2 = ref_logits.topk(2, dim=1).values
    save("parse_256", img_seed=6, seg19=seg19.to(torch.uint8), seg12=seg12.to(torch.uint8),
         logits_sub=ref_logits[:, :, ::16, ::16], margin=(top2[:, 0] - top2[:, 1]).to(torch.float16))


def golden_paste():
    print("paste-back (Pillow itself: resize BILINEAR, PERSPECTIVE transform, alpha composite)")
    import PIL
    from PIL import Image
    rng = np.random.default_rng(11)
    # the restatement of Pillow's resampling / transform is bit exact at several shapes ...
    for (h, w, oh, ow) in ((37, 53, 74, 106), (64, 64, 128, 128), (50, 50, 75, 120), (96, 96, 48, 60)):
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(a).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(ref, O.pil_resize_bilinear_u8(a, oh, ow)), (h, w, oh, ow)
    # ... and so is the whole per-frame paste-back of scripts/inference_swap_video.py:702-721
    g = torch.Generator().manual_seed(12)
    x01 = torch.rand(3, 64, 64, generator=g).clamp(0, 1)
    orig = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    coeffs = np.array([0.93, 0.07, -14.0, -0.05, 1.04, -9.0, 3e-5, -2e-5])
    up = 128
    x_sample = 255. * np.transpose(x01.numpy(), (1, 2, 0))
    img = Image.fromarray(x_sample.astype(np.uint8)).resize((up, up), Image.BILINEAR)
    swapped_and_pasted = img.convert('RGBA')
    pasted_image = Image.fromarray(orig).convert('RGBA')
    swapped_and_pasted.putalpha(255)
    projected = swapped_and_pasted.transform((orig.shape[1], orig.shape[0]), Image.PERSPECTIVE, coeffs, Image.BILINEAR)
    pasted_image.alpha_composite(projected)
    ref = np.asarray(pasted_image)[..., :3]
    assert set(np.unique(np.asarray(projected)[..., 3])) <= {0, 255}
    mine = O.paste_back(x01.numpy(), orig, coeffs, up=up)
    assert np.array_equal(ref, mine)
    print(f"  paste-back bit exact vs Pillow {PIL.__version__}; inside fraction {(np.asarray(projected)[..., 3] == 255).mean():.3f}")
    save("paste_64", x01=x01, orig=orig, coeffs=coeffs, up=up, pasted=ref, pillow=np.array(PIL.__version__))


def _ref_unet():
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    m = UNetModel(image_size=32, in_channels=9, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                  num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                  transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False,
                  add_conv_in_front_of_unet=False).eval()
    sd, sub = sub_sd(O.unet_spec(), O.PFX_UNET)
    m.load_state_dict(sub, strict=True)
    return m, sd


def _ref_vae():
    from ldm.models.autoencoder import AutoencoderKL
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    m = AutoencoderKL(ddconfig=dd, lossconfig={"target": "torch.nn.Identity"}, embed_dim=4).eval()
    sd, sub = sub_sd(O.vae_spec(), O.PFX_VAE)
    m.load_state_dict(sub, strict=True)
    return m, sd


def _fake_ld(unet, split=False):
    """The attributes DDIMSampler touches on LatentDiffusion (ddim.py:100,113-119,207,345); apply_model as
    ddpm.py:1519-1617 + DiffusionWrapper.forward :2244-2246.  split=True runs the batch one sample at a time (every
    sample is independent; keeps the fp32 score matrices of the L=128 level inside this container's memory)."""
    ac = O.alphas_cumprod_f32()

    class FakeLD:
        num_timesteps = 1000
        betas = torch.tensor(O.make_beta_schedule(), dtype=torch.float32)
        alphas_cumprod = ac
        alphas_cumprod_prev = torch.tensor(np.append(1.0, ac.double().numpy()[:-1]), dtype=torch.float32)
        device = torch.device("cpu")

        def apply_model(self, x, t, c):
            cc = torch.cat([c], 1)
            if not split:
                return unet(x, t, context=cc)
            return torch.cat([unet(x[i:i + 1], t[i:i + 1], context=cc[i:i + 1]) for i in range(x.shape[0])])

    return FakeLD()


def golden_unet_L128():
    """BASELINE configs[3] resolution: ONE forward of the real reference UNetModel at L=128 (1024x1024 images), N=1
    (openaimodel.py:860-907).  Input regenerated from the seed by the tests."""
    print("unet L=128 (reference UNetModel, N=1)")
    m, sd = _ref_unet()
    g = torch.Generator().manual_seed(21)
    x = torch.randn(1, 9, 128, 128, generator=g)
    ctx = torch.randn(1, 1, 768, generator=g)
    t = torch.tensor([601])
    ref = m(x, t, context=ctx)
    ora = O.unet_forward(O.Params(sd, O.PFX_UNET), x, t, ctx)
    check("unet L=128", ref, ora, 2e-5)
    save("unet_L128", seed=21, t=t, eps=ref)


def golden_ddim_video():
    """The video settings (inference_video_swap.sh:28-29): --ddim_steps 30 => 31 timesteps (util.py:48-49), scale 3."""
    print("ddim, video settings: S=30 -> 31 steps, scale 3 (reference DDIMSampler + UNet)")
    from ldm.models.diffusion.ddim import DDIMSampler
    DDIMSampler.register_buffer = lambda self, n, a: setattr(self, n, a)
    m, sd = _ref_unet()
    g = torch.Generator().manual_seed(22)
    B, L, S, scale = 1, 16, 30, 3.0
    x_T = torch.randn(B, 4, L, L, generator=g)
    z = torch.randn(B, 4, L, L, generator=g)
    mask = (torch.rand(B, 1, L, L, generator=g) > 0.5).float()
    c = torch.randn(B, 1, 768, generator=g)
    uc = torch.randn(B, 1, 768, generator=g)
    smp = DDIMSampler(_fake_ld(m))
    ref, inter = smp.sample(S=S, conditioning=c, batch_size=B, shape=[4, L, L], verbose=False,
                            unconditional_guidance_scale=scale, unconditional_conditioning=uc, eta=0.0, x_T=x_T,
                            log_every_t=1, test_model_kwargs={"inpaint_image": z, "inpaint_mask": mask})
    assert len(smp.ddim_timesteps) == 31 and len(inter["x_inter"]) == 32
    ora, _ = O.ddim_sample(O.Params(sd, O.PFX_UNET), x_T, z, mask, c, uc, S, scale)
    check("ddim video x0", ref, ora, 5e-5)
    save("ddim_S30_L16", x_T=x_T, z=z, mask=mask, c=c, uc=uc, x0=ref, x_inter=torch.stack(inter["x_inter"][1:]),
         timesteps=np.asarray(smp.ddim_timesteps))


def _img_fixture(img):
    """A decoded image is too large for a fixture: keep every third pixel and one full-resolution centre crop."""
    H = img.shape[-1]
    return dict(img_sub=img[..., ::3, ::3].contiguous(), img_crop=img[..., H // 2 - 48:H // 2 + 48, H // 2 - 48:H // 2 + 48].contiguous())


def golden_vae_big():
    """VAE encode / decode at the BASELINE resolutions (512x512: configs[1,2,4]; 1024x1024: configs[3]) through the real
    reference AutoencoderKL (autoencoder.py:324-333, model.py:434-459,535-568).  Inputs regenerated from the seed."""
    m, sd = _ref_vae()
    P = O.Params(sd, O.PFX_VAE)
    for H, seed in ((512, 31), (1024, 32)):
        print(f"vae {H}x{H} (reference AutoencoderKL)")
        g = torch.Generator().manual_seed(seed)
        L = H // 8
        x = torch.rand(1, 3, H, H, generator=g) * 2 - 1
        noise = torch.randn(1, 4, L, L, generator=g)
        zdec = torch.randn(1, 4, L, L, generator=g) * 0.18215 * 3
        post = m.encode(x)
        ref_z = O.SCALE_FACTOR * (post.mean + post.std * noise)
        check("vae z", ref_z, O.vae_encode(P, x, noise), 2e-5)
        ref_img = m.decode((1.0 / O.SCALE_FACTOR) * zdec)
        check("vae decode", ref_img, O.vae_decode(P, zdec), 2e-5)
        save(f"vae_{H}", seed=seed, mean=post.mean, logvar=post.logvar, z=ref_z, **_img_fixture(ref_img))


def _ref_conditioner(landmarks_raw=None):
    """The reference's conditioning_with_feat (ddpm.py:872-1045) driven through a namespace that carries exactly the
    attributes it reads; real FrozenCLIPEmbedder / IDLoss / Linear layers with the oracle's seeded weights."""
    from src.Face_models.encoders.model_irse import Backbone
    import ldm.models.diffusion.ddpm as ddpm
    import torchvision.transforms.functional as TF
    clip_mod = _build_clip_embedder()
    csd, csub = sub_sd(O.clip_spec(), O.PFX_CLIP)
    clip_mod.load_state_dict(csub, strict=False)
    bb = Backbone(input_size=112, num_layers=50, drop_ratio=0.6, mode="ir_se").eval()
    asd, asub = sub_sd(O.arcface_spec(), O.PFX_ARC)
    bb.load_state_dict(asub, strict=False)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "arc.pth")
        torch.save(bb.state_dict(), path)
        idl = ddpm.IDLoss(types.SimpleNamespace(other_params=types.SimpleNamespace(arcface_path=path))).eval()
    fsd = O.init_state_dict(O.fusion_spec(), SEED)
    lin = lambda n, i, o: _mk_linear(fsd, n, i, o)
    fake = types.SimpleNamespace(
        training=False, update_weight=False, clip_weight=1.0, ID_weight=10.0, Landmarks_weight=0.05,
        Source_CLIP_feat=True, Target_CLIP_feat=True, use_3dmm=False, normalize=False, Landmark_cond=True,
        weight_division=True, concat_feat=False, stack_feat=False, land_mark_id_seperate_layers=False,
        sep_head_att=False, device=torch.device("cpu"),
        get_learned_conditioning=lambda x: clip_mod.encode(x), face_ID_model=idl,
        proj_out_source=lin("proj_out_source", 768, 768), proj_out_target=lin("proj_out_target", 768, 768),
        ID_proj_out=lin("ID_proj_out", 512, 768))
    lm_proj = lin("landmark_proj_out", 136, 768)
    _resize = TF.resize

    def run(ref_img, tar, lm_raw):
        # torchvision>=0.17 defaults to antialias=True on tensors; the pinned 0.14 does not
        ddpm.TF.resize = lambda img, size, *a, **k: _resize(img, list(size), antialias=False)
        try:
            lm = lm_proj(lm_raw)                                           # get_landmarks, ddpm.py:1096
            return ddpm.LatentDiffusion.conditioning_with_feat(fake, ref_img, lm, tar=tar), lm
        finally:
            ddpm.TF.resize = _resize

    full = dict(csd); full.update(asd); full.update(fsd)
    return run, full


def golden_cond_landmarks():
    """conditioning_with_feat with DETECTED landmarks: raw 68x2 dlib pixel coordinates (values up to the image size)
    projected by landmark_proj_out (ddpm.py:1085-1096), i.e. the call of scripts/inference_test_bench.py:447-448."""
    print("conditioning with detected (non-zero) landmarks")
    run, full = _ref_conditioner()
    g = torch.Generator().manual_seed(41)
    ref_img = torch.randn(2, 3, 224, 224, generator=g)
    tar = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    lm_raw = torch.floor(torch.rand(2, 136, generator=g) * 512)           # dlib part coordinates are integers
    ref_c, lm = run(ref_img, tar, lm_raw)
    ora_c = O.conditioning_with_feat(O.Params(full), ref_img, tar, lm_raw)
    check("conditioning (landmarks)", ref_c, ora_c, 5e-5)
    save("cond_B2_lm", seed=41, lm_raw=lm_raw, lm_proj=lm, c=ref_c)


def golden_full(H=512, S=50, scale=3.5, name="full_512_S50"):
    """The WHOLE path on BASELINE's own configuration through the real reference modules: conditioning_with_feat
    (CLIP x2 + ArcFace + fusion) -> AutoencoderKL.encode + sample -> DDIMSampler.sample (S steps, CFG) over the
    reference UNetModel -> AutoencoderKL.decode -> clamp((x+1)/2)  [scripts/inference_test_bench.py:438-495], B=1,
    inputs = oracle.synthetic_inputs(1, H, seed=42) (regenerated from the seed by the tests)."""
    import time
    print(f"full path {H}x{H}, {S} DDIM steps, CFG {scale}, B=1 through the reference modules")
    from ldm.models.diffusion.ddim import DDIMSampler
    DDIMSampler.register_buffer = lambda self, n, a: setattr(self, n, a)
    inp = O.synthetic_inputs(1, H, seed=42)
    run, full = _ref_conditioner()
    c, _ = run(inp["ref_img"], inp["tar_img"], inp["landmarks136"])
    del run
    vae, vsd = _ref_vae()
    post = vae.encode(inp["inpaint_img"])
    z = O.SCALE_FACTOR * (post.mean + post.std * inp["enc_noise"])          # ddpm.py:850-857 with explicit noise
    unet, usd = _ref_unet()
    fsd = O.init_state_dict(O.fusion_spec(), SEED)
    uc = fsd["learnable_vector"].repeat(1, 1, 1)
    L = H // 8
    smp = DDIMSampler(_fake_ld(unet, split=(L > 64)))
    t0 = time.time()
    x0, inter = smp.sample(S=S, conditioning=c, batch_size=1, shape=[4, L, L], verbose=False,
                           unconditional_guidance_scale=scale, unconditional_conditioning=uc, eta=0.0, x_T=inp["x_T"],
                           log_every_t=1, test_model_kwargs={"inpaint_image": z, "inpaint_mask": inp["mask_lat"]})
    print(f"  sampler: {time.time() - t0:.0f} s")
    img = torch.clamp((vae.decode((1.0 / O.SCALE_FACTOR) * x0) + 1.0) / 2.0, 0.0, 1.0)
    xi = torch.stack(inter["x_inter"][1:])                                  # [steps,1,4,L,L], x after every step
    save(name, seed=42, H=H, S=S, scale=scale, c=c, z_inpaint=z, samples=x0, x_inter_sub=xi[..., ::4, ::4].contiguous(),
         x_inter_absmax=xi.abs().amax(dim=(1, 2, 3, 4)), **_img_fixture(img))


if __name__ == "__main__":
    which = sys.argv[1:] or ["schedule", "unet", "vae", "clip", "parse", "paste"]
    if "schedule" in which:
        golden_schedule()
    if "unet" in which:
        m, sd = golden_unet()
        golden_ddim(m, sd)
        golden_plms(m, sd)
        del m, sd
    if "plms" in which and "unet" not in which:      # regenerate only the PLMS / q_sample fixture
        m, sd = golden_unet()
        golden_plms(m, sd)
        del m, sd
    if "vae" in which:
        golden_vae()
    if "clip" in which:
        cm, csd = golden_clip()
        golden_arcface_and_fusion(cm, csd)
    if "parse" in which:
        golden_parse()
    if "paste" in which:
        golden_paste()
    # round 2: BASELINE's own configurations (run explicitly: minutes of CPU each)
    if "unet128" in which:
        golden_unet_L128()
    if "video" in which:
        golden_ddim_video()
    if "vaebig" in which:
        golden_vae_big()
    if "condlm" in which:
        golden_cond_landmarks()
    if "full512" in which:
        golden_full(512, 50, 3.5, "full_512_S50")
    if "full1024" in which:
        s1024 = int(os.environ.get("RFB_GOLDEN_S1024", 10))
        golden_full(1024, s1024, 3.5, f"full_1024_S{s1024}")
    print("OK")
