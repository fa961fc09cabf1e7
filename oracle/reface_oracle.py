"""CPU fp32 restatement of REFace's DDIM face-swap inference path (TEST INFRASTRUCTURE ONLY).

This file is the *oracle*: a plain-PyTorch, CPU, fp32, functional restatement of the reference
algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it; the product (``reface_b200``) never does.  Every function cites the reference file:line
(relative to /root/reference) it restates.  Parameters live in a flat ``dict[str, Tensor]`` that
uses the reference's own ``state_dict`` key names, so the very same dict can be loaded into the
reference ``nn.Module`` (done in ``tests/golden/make_golden.py``, which pins this oracle against
the real reference modules) and into the CUDA engine's weight packer.

Parity status: pinned against the reference's own modules (UNetModel, AutoencoderKL, Backbone,
FrozenCLIPEmbedder incl. HF CLIPVisionModel, DDIMSampler) run in the build container; the golden
outputs live in tests/golden/*.npz.  The reference ships no tests/golden vectors of its own
(SURVEY.md section 4); the closed-form constants it implies are checked in tests/test_oracle.py.
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# configuration (models/REFace/configs/project_ffhq.yaml:1-99)
# --------------------------------------------------------------------------------------------
UNET_CFG = dict(in_channels=9, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1),
                num_res_blocks=2, channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=768)
VAE_CFG = dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, in_channels=3, out_ch=3,
               embed_dim=4)
CLIP_CFG = dict(width=1024, layers=24, heads=16, patch=14, image=224, mlp=4096, proj=768,
                mapper_layers=5)
SCALE_FACTOR = 0.18215
LINEAR_START, LINEAR_END, TIMESTEPS = 0.00085, 0.012, 1000
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
CLIP_WEIGHT, ID_WEIGHT, LANDMARK_WEIGHT = 1.0, 10.0, 0.05

PFX_UNET = "model.diffusion_model."
PFX_VAE = "first_stage_model."
PFX_CLIP = "cond_stage_model."
PFX_ARC = "face_ID_model.facenet."


# --------------------------------------------------------------------------------------------
# parameter access: one code path both *declares* the parameter spec (meta tensors) and *uses* a
# real state dict, so the spec can never drift from the forward pass.
# --------------------------------------------------------------------------------------------
class Params:
    """sd=None -> spec mode: records (shape, kind) and returns meta tensors."""

    def __init__(self, sd=None, prefix=""):
        self.sd, self.prefix = sd, prefix
        self.spec = OrderedDict()

    def __call__(self, name, shape, kind):
        full = self.prefix + name
        shape = tuple(int(s) for s in shape)
        if self.sd is None:
            if full in self.spec:
                assert self.spec[full][0] == shape
            self.spec[full] = (shape, kind)
            return torch.empty(shape, device="meta")
        t = self.sd[full]
        assert tuple(t.shape) == shape, (full, tuple(t.shape), shape)
        return t

    def sub(self, prefix):
        p = Params(self.sd, self.prefix + prefix)
        p.spec = self.spec
        return p


def _seed_for(name: str, seed: int) -> int:
    return int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:7], "little")


def init_tensor(name, shape, kind, seed=0):
    """Deterministic per-tensor init (independent of generation order).  'w': N(0, 1/fan_in);
    zero_module'd tensors of the reference get the same treatment."""
    g = torch.Generator().manual_seed(_seed_for(name, seed))
    n = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32)
    if kind == "w":
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
        return n(*shape) / math.sqrt(fan_in)
    if kind == "b":
        return 0.02 * n(*shape)
    if kind == "g":
        return 1.0 + 0.1 * n(*shape)
    if kind == "emb":
        return 0.02 * n(*shape)
    if kind == "unit":
        return n(*shape)
    if kind == "bn_mean":
        return 0.1 * n(*shape)
    if kind == "bn_var":
        return 0.5 + torch.rand(*shape, generator=g, dtype=torch.float32)
    if kind == "prelu":
        return 0.25 + 0.05 * n(*shape)
    if kind == "count":
        return torch.zeros(shape, dtype=torch.long)
    raise ValueError(kind)


def init_state_dict(spec, seed=0):
    return OrderedDict((k, init_tensor(k, s, kind, seed)) for k, (s, kind) in spec.items())


# --------------------------------------------------------------------------------------------
# schedule (ldm/modules/diffusionmodules/util.py:21-74, ldm/models/diffusion/ddpm.py:255-275,
# ldm/models/diffusion/ddim.py:110-139)
# --------------------------------------------------------------------------------------------
def make_beta_schedule(n=TIMESTEPS, linear_start=LINEAR_START, linear_end=LINEAR_END):
    # util.py:23-25  (fp64)
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n, dtype=torch.float64) ** 2).numpy()


def alphas_cumprod_f32():
    # ddpm.py:262-275: cumprod in fp64, cast to fp32
    betas = make_beta_schedule()
    return torch.tensor(np.cumprod(1.0 - betas, axis=0), dtype=torch.float32)


def make_ddim_timesteps(S, T=TIMESTEPS):
    # util.py:46-60 ('uniform'); NB: S=30 yields 31 steps (assert commented out at util.py:55)
    c = T // S
    return np.asarray(list(range(0, T, c))) + 1


def ddim_schedule(S, eta=0.0):
    """Returns dict of fp32 numpy tables indexed by ddim index (ddim.py:110-139, util.py:63-74)."""
    ac = alphas_cumprod_f32()
    ts = make_ddim_timesteps(S)
    alphas = ac[ts]                                                        # fp32 tensor
    alphas_prev = np.asarray([float(ac[0])] + ac[ts[:-1]].tolist())        # fp64 holding fp32 values
    a64 = alphas.double().numpy()
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - a64) * (1 - a64 / alphas_prev))
    return dict(timesteps=ts.astype(np.int64),
                a_t=alphas.numpy().astype(np.float32),
                a_prev=alphas_prev.astype(np.float32),
                sigma=np.asarray(sigmas, dtype=np.float32),
                sqrt_one_minus_a=torch.sqrt(1.0 - alphas).numpy().astype(np.float32))


def timestep_embedding(t, dim, max_period=10000):
    # util.py:151-171 (cos || sin)
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


# --------------------------------------------------------------------------------------------
# UNet (ldm/modules/diffusionmodules/openaimodel.py, ldm/modules/attention.py)
# --------------------------------------------------------------------------------------------
def _conv(P, name, x, cin, cout, k, stride=1, padding=0, bias=True):
    w = P(name + ".weight", (cout, cin, k, k), "w")
    b = P(name + ".bias", (cout,), "b") if bias else None
    return F.conv2d(x, w, b, stride=stride, padding=padding)


def _linear(P, name, x, cin, cout, bias=True):
    w = P(name + ".weight", (cout, cin), "w")
    b = P(name + ".bias", (cout,), "b") if bias else None
    return F.linear(x, w, b)


def _gn(P, name, x, c, eps):
    return F.group_norm(x, 32, P(name + ".weight", (c,), "g"), P(name + ".bias", (c,), "b"), eps)


def _ln(P, name, x, c, eps=1e-5):
    return F.layer_norm(x, (c,), P(name + ".weight", (c,), "g"), P(name + ".bias", (c,), "b"), eps)


def unet_resblock(P, x, emb, cin, cout):
    # openaimodel.py:255-275 (use_scale_shift_norm=False, no up/down)
    h = _conv(P, "in_layers.2", F.silu(_gn(P, "in_layers.0", x, cin, 1e-5)), cin, cout, 3, padding=1)
    e = _linear(P, "emb_layers.1", F.silu(emb), 1280, cout)
    h = h + e[:, :, None, None]
    h = _conv(P, "out_layers.3", F.silu(_gn(P, "out_layers.0", h, cout, 1e-5)), cout, cout, 3, padding=1)
    if cin != cout:
        x = _conv(P, "skip_connection", x, cin, cout, 1)
    return x + h


def cross_attention(P, x, ctx, dim, heads, ctx_dim):
    # attention.py:179-221
    d = dim // heads
    q = _linear(P, "to_q", x, dim, dim, bias=False)
    k = _linear(P, "to_k", ctx, ctx_dim, dim, bias=False)
    v = _linear(P, "to_v", ctx, ctx_dim, dim, bias=False)
    b, n, _ = q.shape
    sp = lambda t: t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)
    q, k, v = sp(q), sp(k), sp(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", attn, v)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, dim)
    return _linear(P, "to_out.0", out, dim, dim)


def spatial_transformer(P, x, ctx, c, heads, ctx_dim):
    # attention.py:278-289 + BasicTransformerBlock._forward :239-243 + GEGLU :37-45
    b, _, h, w = x.shape
    x_in = x
    x = _gn(P, "norm", x, c, 1e-6)
    x = _conv(P, "proj_in", x, c, c, 1)
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    B = P.sub("transformer_blocks.0.")
    xn = _ln(B, "norm1", x, c)
    x = cross_attention(B.sub("attn1."), xn, xn, c, heads, c) + x
    x = cross_attention(B.sub("attn2."), _ln(B, "norm2", x, c), ctx, c, heads, ctx_dim) + x
    y = _linear(B, "ff.net.0.proj", _ln(B, "norm3", x, c), c, 8 * c)
    a, gate = y.chunk(2, dim=-1)
    x = _linear(B, "ff.net.2", a * F.gelu(gate), 4 * c, c) + x
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return _conv(P, "proj_out", x, c, c, 1) + x_in


def unet_layout(cfg=UNET_CFG):
    """Block list mirroring UNetModel.__init__ (openaimodel.py:665-830).  Each entry is a list of
    ops: ('res', cin, cout) | ('attn', c) | ('down', c) | ('up', c) | ('conv_in', cin, cout)."""
    mc, mult, nrb = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    inp = [[("conv_in", cfg["in_channels"], mc)]]
    chans, ch, ds = [mc], mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            ops = [("res", ch, m * mc)]
            ch = m * mc
            if ds in cfg["attention_resolutions"]:
                ops.append(("attn", ch))
            inp.append(ops)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    mid = [("res", ch, ch), ("attn", ch), ("res", ch, ch)]
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            ops = [("res", ch + ich, mc * m)]
            ch = mc * m
            if ds in cfg["attention_resolutions"]:
                ops.append(("attn", ch))
            if level and i == nrb:
                ops.append(("up", ch))
                ds //= 2
            out.append(ops)
    return inp, mid, out


def _run_ops(P, ops, h, emb, ctx, cfg):
    for j, op in enumerate(ops):
        Q = P.sub(f"{j}.")
        if op[0] == "conv_in":
            h = F.conv2d(h, Q("weight", (op[2], op[1], 3, 3), "w"), Q("bias", (op[2],), "b"), padding=1)
        elif op[0] == "res":
            h = unet_resblock(Q, h, emb, op[1], op[2])
        elif op[0] == "attn":
            h = spatial_transformer(Q, h, ctx, op[1], cfg["num_heads"], cfg["context_dim"])
        elif op[0] == "down":  # openaimodel.py:151,158-160
            h = _conv(Q, "op", h, op[1], op[1], 3, stride=2, padding=1)
        elif op[0] == "up":    # openaimodel.py:109-119
            h = F.interpolate(h, scale_factor=2, mode="nearest") if h.device.type != "meta" else \
                h.new_empty(h.shape[0], h.shape[1], h.shape[2] * 2, h.shape[3] * 2)
            h = _conv(Q, "conv", h, op[1], op[1], 3, padding=1)
    return h


def unet_forward(P, x, t, ctx, cfg=UNET_CFG, taps=None):
    """UNetModel.forward (openaimodel.py:860-907).  x [N,9,L,L] fp32, t [N] int64, ctx [N,T,768]."""
    mc = cfg["model_channels"]
    inp, mid, out = unet_layout(cfg)
    temb = timestep_embedding(t, mc) if x.device.type != "meta" else x.new_empty(x.shape[0], mc)
    emb = _linear(P, "time_embed.2", F.silu(_linear(P, "time_embed.0", temb, mc, 4 * mc)), 4 * mc, 4 * mc)
    hs, h = [], x
    for i, ops in enumerate(inp):
        h = _run_ops(P.sub(f"input_blocks.{i}."), ops, h, emb, ctx, cfg)
        hs.append(h)
        if taps is not None:
            taps[f"input_blocks.{i}"] = h
    h = _run_ops(P.sub("middle_block."), mid, h, emb, ctx, cfg)
    if taps is not None:
        taps["middle_block"] = h
    for i, ops in enumerate(out):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_ops(P.sub(f"output_blocks.{i}."), ops, h, emb, ctx, cfg)
        if taps is not None:
            taps[f"output_blocks.{i}"] = h
    h = F.silu(_gn(P, "out.0", h, mc, 1e-5))
    return _conv(P, "out.2", h, mc, cfg["out_channels"], 3, padding=1)


# --------------------------------------------------------------------------------------------
# VAE (ldm/modules/diffusionmodules/model.py, ldm/models/autoencoder.py,
#      ldm/modules/distributions/distributions.py)
# --------------------------------------------------------------------------------------------
def _swish(x):  # model.py:33-35
    return x * torch.sigmoid(x)


def vae_resnet(P, x, cin, cout):
    # model.py:121-141 (temb is None, dropout 0)
    h = _conv(P, "conv1", _swish(_gn(P, "norm1", x, cin, 1e-6)), cin, cout, 3, padding=1)
    h = _conv(P, "conv2", _swish(_gn(P, "norm2", h, cout, 1e-6)), cout, cout, 3, padding=1)
    if cin != cout:
        x = _conv(P, "nin_shortcut", x, cin, cout, 1)
    return x + h


def vae_attn(P, x, c):
    # model.py:178-202: single head, d=c, scale c^-0.5
    h = _gn(P, "norm", x, c, 1e-6)
    q, k, v = (_conv(P, n, h, c, c, 1) for n in ("q", "k", "v"))
    b, _, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(P, "proj_out", h, c, c, 1)


def vae_encoder(P, x, cfg=VAE_CFG):
    # model.py:434-459
    ch, mult, nrb = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    h = _conv(P, "conv_in", x, cfg["in_channels"], ch, 3, padding=1)
    cin = ch
    for lvl, m in enumerate(mult):
        for j in range(nrb):
            h = vae_resnet(P.sub(f"down.{lvl}.block.{j}."), h, cin, ch * m)
            cin = ch * m
        if lvl != len(mult) - 1:
            # model.py:72-79: pad (0,1,0,1) then stride-2 conv, padding 0
            if h.device.type != "meta":
                h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)
            else:
                h = h.new_empty(h.shape[0], h.shape[1], h.shape[2] + 1, h.shape[3] + 1)
            h = _conv(P, f"down.{lvl}.downsample.conv", h, cin, cin, 3, stride=2, padding=0)
    h = vae_resnet(P.sub("mid.block_1."), h, cin, cin)
    h = vae_attn(P.sub("mid.attn_1."), h, cin)
    h = vae_resnet(P.sub("mid.block_2."), h, cin, cin)
    h = _swish(_gn(P, "norm_out", h, cin, 1e-6))
    return _conv(P, "conv_out", h, cin, 2 * cfg["z_channels"], 3, padding=1)


def vae_decoder(P, z, cfg=VAE_CFG):
    # model.py:535-568
    ch, mult, nrb = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    cin = ch * mult[-1]
    h = _conv(P, "conv_in", z, cfg["z_channels"], cin, 3, padding=1)
    h = vae_resnet(P.sub("mid.block_1."), h, cin, cin)
    h = vae_attn(P.sub("mid.attn_1."), h, cin)
    h = vae_resnet(P.sub("mid.block_2."), h, cin, cin)
    for lvl in reversed(range(len(mult))):
        cout = ch * mult[lvl]
        for j in range(nrb + 1):
            h = vae_resnet(P.sub(f"up.{lvl}.block.{j}."), h, cin, cout)
            cin = cout
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest") if h.device.type != "meta" else \
                h.new_empty(h.shape[0], h.shape[1], h.shape[2] * 2, h.shape[3] * 2)
            h = _conv(P, f"up.{lvl}.upsample.conv", h, cin, cin, 3, padding=1)
    h = _swish(_gn(P, "norm_out", h, cin, 1e-6))
    return _conv(P, "conv_out", h, cin, cfg["out_ch"], 3, padding=1)


def vae_encode_moments(P, x, cfg=VAE_CFG):
    # autoencoder.py:324-328
    h = vae_encoder(P.sub("encoder."), x, cfg)
    m = _conv(P, "quant_conv", h, 2 * cfg["z_channels"], 2 * cfg["embed_dim"], 1)
    mean, logvar = m.chunk(2, dim=1)
    return mean, torch.clamp(logvar, -30.0, 20.0)          # distributions.py:27-28


def vae_encode(P, x, noise, cfg=VAE_CFG):
    """get_first_stage_encoding(encode_first_stage(x)) (ddpm.py:850-857, distributions.py:35-37);
    the posterior noise is an explicit argument."""
    mean, logvar = vae_encode_moments(P, x, cfg)
    return SCALE_FACTOR * (mean + torch.exp(0.5 * logvar) * noise)


def vae_decode(P, z, cfg=VAE_CFG):
    # ddpm.py:1277-1337 (z/scale, first 4 channels) -> autoencoder.py:330-333
    z = (1.0 / SCALE_FACTOR) * z[:, :4]
    z = _conv(P, "post_quant_conv", z, cfg["embed_dim"], cfg["z_channels"], 1)
    return vae_decoder(P.sub("decoder."), z, cfg)


# --------------------------------------------------------------------------------------------
# CLIP ViT-L/14 vision tower (third-party: transformers CLIPVisionModel, called at
# ldm/modules/encoders/modules.py:254-256) + mapper2/final_ln2 (modules.py:259-260, xf.py)
# --------------------------------------------------------------------------------------------
def _quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


def clip_vision_pooled(P, img, cfg=CLIP_CFG):
    """HF CLIPVisionTransformer: patch conv (no bias) + CLS + pos-emb -> pre_layrnorm -> 24 pre-LN
    layers (MHA 16x64 with q scaled by d^-0.5, quick-GELU MLP) -> post_layernorm(CLS)."""
    W, nh, p = cfg["width"], cfg["heads"], cfg["patch"]
    ntok = (cfg["image"] // p) ** 2 + 1
    E = P.sub("embeddings.")
    x = F.conv2d(img, E("patch_embedding.weight", (W, 3, p, p), "w"), None, stride=p)
    b = x.shape[0]
    x = x.flatten(2).transpose(1, 2)
    cls = E("class_embedding", (W,), "emb")
    x = torch.cat([cls.expand(b, 1, W), x], dim=1) + E("position_embedding.weight", (ntok, W), "emb")[None]
    x = _ln(P, "pre_layrnorm", x, W)
    d = W // nh
    for i in range(cfg["layers"]):
        L = P.sub(f"encoder.layers.{i}.")
        h = _ln(L, "layer_norm1", x, W)
        q = _linear(L, "self_attn.q_proj", h, W, W) * (d ** -0.5)
        k = _linear(L, "self_attn.k_proj", h, W, W)
        v = _linear(L, "self_attn.v_proj", h, W, W)
        sp = lambda t: t.reshape(b, -1, nh, d).transpose(1, 2)
        a = torch.softmax(sp(q) @ sp(k).transpose(-1, -2), dim=-1) @ sp(v)
        a = a.transpose(1, 2).reshape(b, -1, W)
        x = x + _linear(L, "self_attn.out_proj", a, W, W)
        h = _ln(L, "layer_norm2", x, W)
        x = x + _linear(L, "mlp.fc2", _quick_gelu(_linear(L, "mlp.fc1", h, W, cfg["mlp"])), cfg["mlp"], W)
    return _ln(P, "post_layernorm", x[:, 0], W)


def xf_block(P, x, width):
    # xf.py:66-101 with n_ctx=1, heads=1: softmax over one key == 1 -> attention returns v
    h = _ln(P, "ln_1", x, width)
    qkv = _linear(P, "attn.c_qkv", h, width, 3 * width)
    bs, n_ctx, _ = qkv.shape
    attn_ch = width
    scale = 1 / math.sqrt(math.sqrt(attn_ch))
    q, k, v = torch.split(qkv.view(bs, n_ctx, 1, -1), attn_ch, dim=-1)
    w = torch.softmax(torch.einsum("bthc,bshc->bhts", q * scale, k * scale).float(), dim=-1)
    a = torch.einsum("bhts,bshc->bthc", w, v).reshape(bs, n_ctx, -1)
    x = x + _linear(P, "attn.c_proj", a, width, width)
    h = _ln(P, "ln_2", x, width)
    return x + _linear(P, "mlp.c_proj", F.gelu(_linear(P, "mlp.c_fc", h, width, 4 * width)), 4 * width, width)


def clip_embed(P, img, cfg=CLIP_CFG):
    """FrozenCLIPEmbedder.forward (encoders/modules.py:253-261): [B,3,224,224] -> [B,1,768]."""
    z = clip_vision_pooled(P.sub("model.vision_model."), img, cfg)
    z = F.linear(z, P("model.visual_projection.weight", (cfg["proj"], cfg["width"]), "w"))
    z = z.unsqueeze(1)
    for i in range(cfg["mapper_layers"]):
        z = xf_block(P.sub(f"mapper2.resblocks.{i}."), z, cfg["proj"])
    return _ln(P, "final_ln2", z, cfg["proj"])


# --------------------------------------------------------------------------------------------
# ArcFace IR-SE50 (src/Face_models/encoders/{model_irse,helpers}.py) + pre-processing
# (ldm/models/diffusion/ddpm.py:112-124)
# --------------------------------------------------------------------------------------------
def _bn(P, name, x, c):
    return F.batch_norm(x, P(name + ".running_mean", (c,), "bn_mean"), P(name + ".running_var", (c,), "bn_var"),
                        P(name + ".weight", (c,), "g"), P(name + ".bias", (c,), "b"), False, 0.0, 1e-5)


ARC_BLOCKS = [(64, 64, 3), (64, 128, 4), (128, 256, 14), (256, 512, 3)]  # helpers.py:29-36


def arcface_units():
    units = []
    for cin, depth, n in ARC_BLOCKS:
        units.append((cin, depth, 2))
        units += [(depth, depth, 1)] * (n - 1)
    return units


def arcface_backbone(P, x):
    # model_irse.py:44-69, helpers.py:97-119, :56-72, :15-18
    x = F.conv2d(x, P("input_layer.0.weight", (64, 3, 3, 3), "w"), None, 1, 1)
    x = F.prelu(_bn(P, "input_layer.1", x, 64), P("input_layer.2.weight", (64,), "prelu"))
    for i, (cin, depth, stride) in enumerate(arcface_units()):
        U = P.sub(f"body.{i}.")
        if cin == depth:
            sc = x[:, :, ::stride, ::stride]                                  # MaxPool2d(1, stride)
        else:
            sc = _bn(U, "shortcut_layer.1", F.conv2d(x, U("shortcut_layer.0.weight", (depth, cin, 1, 1), "w"),
                                                     None, stride), depth)
        r = _bn(U, "res_layer.0", x, cin)
        r = F.conv2d(r, U("res_layer.1.weight", (depth, cin, 3, 3), "w"), None, 1, 1)
        r = F.prelu(r, U("res_layer.2.weight", (depth,), "prelu"))
        r = F.conv2d(r, U("res_layer.3.weight", (depth, depth, 3, 3), "w"), None, stride, 1)
        r = _bn(U, "res_layer.4", r, depth)
        s = r.mean(dim=(2, 3), keepdim=True)
        s = F.relu(F.conv2d(s, U("res_layer.5.fc1.weight", (depth // 16, depth, 1, 1), "w")))
        s = torch.sigmoid(F.conv2d(s, U("res_layer.5.fc2.weight", (depth, depth // 16, 1, 1), "w")))
        x = r * s + sc
    x = _bn(P, "output_layer.0", x, 512)
    x = x.reshape(x.shape[0], -1)                                             # Flatten (NCHW order)
    x = _linear(P, "output_layer.3", x, 512 * 7 * 7, 512)
    x = F.batch_norm(x, P("output_layer.4.running_mean", (512,), "bn_mean"),
                     P("output_layer.4.running_var", (512,), "bn_var"),
                     P("output_layer.4.weight", (512,), "g"), P("output_layer.4.bias", (512,), "b"), False, 0.0, 1e-5)
    return x / torch.norm(x, 2, 1, True)


def arcface_preprocess(x):
    """IDLoss.extract_feats (ddpm.py:112-119): un-CLIP-normalise, to [-1,1], AdaptiveAvgPool 256,
    crop [35:223, 32:220], AdaptiveAvgPool 112."""
    mean = torch.tensor(CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD).view(1, 3, 1, 1)
    x = x * std + mean
    x = (x - 0.5) / 0.5
    if x.shape[2] != 256:
        x = F.adaptive_avg_pool2d(x, (256, 256))
    x = x[:, :, 35:223, 32:220]
    return F.adaptive_avg_pool2d(x, (112, 112))


def arcface_embed(P, ref_img):
    return arcface_backbone(P, arcface_preprocess(ref_img))


# --------------------------------------------------------------------------------------------
# conditioning fusion (ldm/models/diffusion/ddpm.py:872-1045, :1068-1099)
# --------------------------------------------------------------------------------------------
def resize_bilinear_noaa(x, size):
    """TF.resize on a tensor with the pinned torchvision 0.14 default (bilinear, antialias off,
    align_corners False) -- ddpm.py:912; see SURVEY 8(c)(iii)."""
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False, antialias=False)


def target_clip_input(tar):
    # ddpm.py:907-912
    mean = torch.tensor(CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD).view(1, 3, 1, 1)
    t = ((tar + 1.0) / 2.0 - mean) / std
    return resize_bilinear_noaa(t, (224, 224))


def conditioning_with_feat(P, ref_img, tar_img, landmarks136):
    """c = (clip*1 + id*10 + lm*0.05)/11.05  -> [B,1,768]  (ddpm.py:904-915, :1010-1012, :1038-1039).
    landmarks136: raw dlib 68x2 landmarks or zeros (ddpm.py:1081-1083) -- an input here."""
    C = P.sub(PFX_CLIP)
    c_src = _linear(P, "proj_out_source", clip_embed(C, ref_img), 768, 768)
    c = _linear(P, "proj_out_target", clip_embed(C, target_clip_input(tar_img)), 768, 768) + c_src
    c2 = _linear(P, "ID_proj_out", arcface_embed(P.sub(PFX_ARC), ref_img), 512, 768).unsqueeze(1)
    lm = _linear(P, "landmark_proj_out", landmarks136, 136, 768).unsqueeze(1)
    tot = CLIP_WEIGHT + ID_WEIGHT + LANDMARK_WEIGHT
    return (c * CLIP_WEIGHT + c2 * ID_WEIGHT + lm * LANDMARK_WEIGHT) / tot


# --------------------------------------------------------------------------------------------
# DDIM (ldm/models/diffusion/ddim.py:200-251, :323-375)
# --------------------------------------------------------------------------------------------
def concat9(x, z_inpaint, mask):
    return torch.cat([x, z_inpaint, mask], dim=1)                              # ddim.py:330


def cfg_ddim_update(x, e_u, e_c, scale, a_t, a_prev, sigma, sqrt_one_minus_at, noise=None):
    """ddim.py:346, :363-374, all fp32, same op order."""
    f = lambda v: torch.full((x.shape[0], 1, 1, 1), float(v), dtype=torch.float32)
    e_t = e_u + scale * (e_c - e_u)
    pred_x0 = (x - f(sqrt_one_minus_at) * e_t) / f(a_t).sqrt()
    dir_xt = (1.0 - f(a_prev) - f(sigma) ** 2).sqrt() * e_t
    nz = f(sigma) * (noise if noise is not None else torch.zeros_like(x))
    return f(a_prev).sqrt() * pred_x0 + dir_xt + nz, pred_x0, e_t


def ddim_sample(P_unet, x_T, z_inpaint, mask, c, uc, S, scale, eta=0.0, log_every_t=100, cfg=UNET_CFG,
                n_steps_limit=None):
    """DDIMSampler.sample with test_model_kwargs (ddim.py:142-251,323-375).  Returns (x0, inter)."""
    sch = ddim_schedule(S, eta)
    ts = sch["timesteps"]
    total = len(ts)
    img = x_T
    inter = {"x_inter": [img], "pred_x0": [img]}
    b = x_T.shape[0]
    for i, step in enumerate(np.flip(ts)):
        if n_steps_limit is not None and i >= n_steps_limit:
            break
        index = total - i - 1
        t = torch.full((b,), int(step), dtype=torch.long)
        x9 = concat9(img, z_inpaint, mask)
        x_in, t_in, c_in = torch.cat([x9] * 2), torch.cat([t] * 2), torch.cat([uc, c])
        e_u, e_c = unet_forward(P_unet, x_in, t_in, c_in, cfg).chunk(2)
        img, pred_x0, _ = cfg_ddim_update(img, e_u, e_c, scale, sch["a_t"][index], sch["a_prev"][index],
                                          sch["sigma"][index], sch["sqrt_one_minus_a"][index])
        if index % log_every_t == 0 or index == total - 1:
            inter["x_inter"].append(img)
            inter["pred_x0"].append(pred_x0)
    return img, inter


def plms_eps_prime(e_t, old_eps, e_t_next=None):
    """p_sample_plms, plms.py:225-240: pseudo improved Euler on the first step, then Adams-Bashforth of order 2-4."""
    if len(old_eps) == 0:
        return (e_t + e_t_next) / 2
    if len(old_eps) == 1:
        return (3 * e_t - old_eps[-1]) / 2
    if len(old_eps) == 2:
        return (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
    return (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24


def plms_sample(P_unet, x_T, z_inpaint, mask, c, uc, S, scale, log_every_t=100, cfg=UNET_CFG, n_steps_limit=None):
    """PLMSSampler.sample with test_model_kwargs (plms.py:58-172 driving p_sample_plms :174-242; eta must be 0, :25-26).
    Returns (x0, inter)."""
    sch = ddim_schedule(S, 0.0)
    ts = sch["timesteps"]
    total = len(ts)
    time_range = np.flip(ts)
    img = x_T
    inter = {"x_inter": [img], "pred_x0": [img]}
    b = x_T.shape[0]
    old_eps = []

    def model_output(x, step):                                   # get_model_output, plms.py:178-192
        t = torch.full((b,), int(step), dtype=torch.long)
        x9 = concat9(x, z_inpaint, mask)
        if uc is None or scale == 1.0:
            return unet_forward(P_unet, x9, t, c, cfg)
        e_u, e_c = unet_forward(P_unet, torch.cat([x9] * 2), torch.cat([t] * 2), torch.cat([uc, c]), cfg).chunk(2)
        return e_u + scale * (e_c - e_u)

    def x_prev_and_pred_x0(x, e, index):                         # get_x_prev_and_pred_x0, plms.py:199-217 (sigma = 0)
        xp, p0, _ = cfg_ddim_update(x, e, e, 1.0, sch["a_t"][index], sch["a_prev"][index], sch["sigma"][index],
                                    sch["sqrt_one_minus_a"][index])
        return xp, p0

    for i, step in enumerate(time_range):
        if n_steps_limit is not None and i >= n_steps_limit:
            break
        index = total - i - 1
        step_next = time_range[min(i + 1, len(time_range) - 1)]
        e_t = model_output(img, step)
        e_next = None
        if len(old_eps) == 0:
            x_prev, _ = x_prev_and_pred_x0(img, e_t, index)
            e_next = model_output(x_prev, step_next)
        e_prime = plms_eps_prime(e_t, old_eps, e_next)
        img, pred_x0 = x_prev_and_pred_x0(img, e_prime, index)
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
        if index % log_every_t == 0 or index == total - 1:
            inter["x_inter"].append(img)
            inter["pred_x0"].append(pred_x0)
    return img, inter


def q_sample(x_start, t, noise):
    """DDPM.q_sample (ddpm.py:412-415) with the fp32 buffers of register_schedule (ddpm.py:283-285): the
    --Start_from_target initialisation of scripts/inference_test_bench.py:414-435."""
    ac = np.cumprod(1.0 - make_beta_schedule(), axis=0)          # fp64, as in register_schedule
    sa = torch.tensor(np.sqrt(ac), dtype=torch.float32)[t].reshape(-1, 1, 1, 1)
    s1 = torch.tensor(np.sqrt(1.0 - ac), dtype=torch.float32)[t].reshape(-1, 1, 1, 1)
    return sa * x_start + s1 * noise


# --------------------------------------------------------------------------------------------
# face parsing: BiSeNet on a ResNet-18 context path + label / mask preparation
#   pretrained/face_parsing/model.py:19-262, resnet.py:19-85, face_parsing_demo.py:74-122,236-281,
#   ldm/data/video_swap_dataset.py:135-240 (mask from the label map, inpaint image)
# --------------------------------------------------------------------------------------------
PFX_PARSE = "face_parser.seg."
SEG_MEAN = (0.485, 0.456, 0.406)     # model.py:15
SEG_STD = (0.229, 0.224, 0.225)      # model.py:16
# __ffhq_masks_to_faceParser_mask_detailed (face_parsing_demo.py:74-122): 19 face-parsing.PyTorch classes -> 12
FFHQ19_TO_12 = (0, 6, 2, 2, 3, 3, 10, 7, 7, 11, 5, 9, 1, 1, 8, 0, 0, 4, 0)
REMOVE_MASK_TAR_FFHQ = (1, 2, 3, 5, 6, 7, 9)   # project_ffhq.yaml:209-216


def _cbr(P, name, x, cin, cout, ks=3, stride=1, padding=1):
    # ConvBNReLU, model.py:19-35
    x = _conv(P, name + ".conv", x, cin, cout, ks, stride, padding, bias=False)
    return F.relu(_bn(P, name + ".bn", x, cout))


def _basic_block(P, name, x, cin, cout, stride):
    # BasicBlock, resnet.py:19-48
    r = _conv(P, name + ".conv1", x, cin, cout, 3, stride, 1, bias=False)
    r = F.relu(_bn(P, name + ".bn1", r, cout))
    r = _conv(P, name + ".conv2", r, cout, cout, 3, 1, 1, bias=False)
    r = _bn(P, name + ".bn2", r, cout)
    sc = x
    if cin != cout or stride != 1:
        sc = _bn(P, name + ".downsample.1", _conv(P, name + ".downsample.0", x, cin, cout, 1, stride, 0, bias=False), cout)
    return F.relu(sc + r)


def resnet18_features(P, x):
    # Resnet18.forward, resnet.py:71-81
    x = F.relu(_bn(P, "bn1", _conv(P, "conv1", x, 3, 64, 7, 2, 3, bias=False), 64))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    cin = 64
    for li, (cout, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), 1):
        x = _basic_block(P, f"layer{li}.0", x, cin, cout, stride)
        x = _basic_block(P, f"layer{li}.1", x, cout, cout, 1)
        cin = cout
        feats.append(x)
    return feats[1], feats[2], feats[3]      # 1/8, 1/16, 1/32


def _arm(P, name, x, cin, cout):
    # AttentionRefinementModule, model.py:71-88
    feat = _cbr(P, name + ".conv", x, cin, cout)
    att = F.avg_pool2d(feat, feat.shape[2:])
    att = _bn(P, name + ".bn_atten", _conv(P, name + ".conv_atten", att, cout, cout, 1, bias=False), cout)
    return feat * torch.sigmoid(att)


def bisenet_logits(P, x, n_classes=19, upsample=True):
    """BiSeNet.forward (model.py:241-256), first output only (the one inference uses, face_parsing_demo.py:277):
    [B,3,H,W] normalised image -> [B,n_classes,H,W] logits."""
    H, W = x.shape[2:]
    cp = P.sub("cp.")
    f8, f16, f32 = resnet18_features(cp.sub("resnet."), x)
    avg = _cbr(cp, "conv_avg", F.avg_pool2d(f32, f32.shape[2:]), 512, 128, 1, 1, 0)
    f32s = _arm(cp, "arm32", f32, 512, 128) + avg                       # avg_up: nearest from 1x1 = broadcast
    f32u = _cbr(cp, "conv_head32", F.interpolate(f32s, f16.shape[2:], mode="nearest"), 128, 128)
    f16s = _arm(cp, "arm16", f16, 256, 128) + f32u
    f16u = _cbr(cp, "conv_head16", F.interpolate(f16s, f8.shape[2:], mode="nearest"), 128, 128)
    # FeatureFusionModule, model.py:175-212 (feat_sp = res3b1 feature f8, model.py:245-246)
    fm = P.sub("ffm.")
    feat = _cbr(fm, "convblk", torch.cat([f8, f16u], 1), 256, 256, 1, 1, 0)
    att = F.avg_pool2d(feat, feat.shape[2:])
    att = torch.sigmoid(_conv(fm, "conv2", F.relu(_conv(fm, "conv1", att, 256, 64, 1, bias=False)), 64, 256, 1, bias=False))
    fuse = feat * att + feat
    # BiSeNetOutput, model.py:43-52
    o = P.sub("conv_out.")
    out = _conv(o, "conv_out", _cbr(o, "conv", fuse, 256, 256), 256, n_classes, 1, bias=False)
    if not upsample:
        return out                                                      # [B, n_classes, H/8, W/8]
    return F.interpolate(out, (H, W), mode="bilinear", align_corners=True)


def face_parse(P, img01):
    """FaceParser.forward for a 512x512 input (face_parsing_demo.py:260-281) + the 19 -> 12 class conversion (:74-122):
    img01 [B,3,512,512] in [0,1] -> (seg19 [B,H,W] int64, seg12 [B,H,W] int64)."""
    mean = torch.tensor(SEG_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(SEG_STD).view(1, 3, 1, 1)
    im = (img01.clamp(0, 1) - mean) / std
    seg19 = torch.argmax(bisenet_logits(P, im), dim=1)
    return seg19, torch.tensor(FFHQ19_TO_12)[seg19]


def inpaint_from_parsing(img_m11, seg12, remove=REMOVE_MASK_TAR_FFHQ):
    """ldm/data/video_swap_dataset.py:150-222: mask = 1 - isin(label, remove_tar); inpaint = image * mask.
    img_m11 [B,3,H,W] in [-1,1], seg12 [B,H,W] -> (mask [B,1,H,W] in {0,1}, inpaint [B,3,H,W])."""
    m = 1.0 - torch.isin(seg12, torch.tensor(remove)).float().unsqueeze(1)
    return m, img_m11 * m


# --------------------------------------------------------------------------------------------
# paste-back: scripts/inference_swap_video.py:702-724 -- the decoded 512x512 face is converted to uint8,
# resized to 1024x1024 with PIL BILINEAR, warped back into the frame with PIL's PERSPECTIVE transform (BILINEAR filter)
# and alpha-composited.  The arithmetic lives in a third-party dependency that is not under /root/reference:
# Pillow (requirements.txt:25 pins 9.0.1, environment.yml:175 pins 9.5.0).  Below is a restatement of its published
# algorithm (src/libImaging/Resample.c: precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc /
# Vertical_8bpc; Geometry.c: perspective_transform, bilinear_filter32RGB, ImagingGenericTransform; AlphaComposite.c),
# pinned bit-for-bit against the Pillow installed in the build container (12.2.0) by tests/golden/make_golden.py.
# --------------------------------------------------------------------------------------------
PIL_PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(insize, outsize):
    """precompute_coeffs (triangle filter, support 1) + normalize_coeffs_8bpc: (bounds [out,2], kk [out,ksize]) int64."""
    scale = insize / outsize
    fs = max(scale, 1.0)
    support = 1.0 * fs
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((outsize, 2), np.int64)
    kk = np.zeros((outsize, ksize), np.int64)
    ss = 1.0 / fs
    for xx in range(outsize):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), insize) - xmin
        w = np.zeros(ksize)
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        ww = w[:xmax].sum()
        if ww != 0.0:
            w[:xmax] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + w[x] * (1 << PIL_PRECISION_BITS)) if w[x] < 0 else int(0.5 + w[x] * (1 << PIL_PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def pil_resize_bilinear_u8(img, out_h, out_w):
    """Image.resize((out_w, out_h), Image.BILINEAR) for a uint8 [H,W,C] image: horizontal then vertical pass, fixed
    point with PRECISION_BITS = 22, uint8 intermediate (inference_swap_video.py:706)."""
    H, W, C = img.shape
    P = PIL_PRECISION_BITS
    bx, kx = pil_bilinear_coeffs(W, out_w)
    tmp = np.zeros((H, out_w, C), np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_w):
        xmin, n = bx[xx]
        acc = np.full((H, C), 1 << (P - 1), np.int64)
        for x in range(n):
            acc += src[:, xmin + x, :] * kx[xx, x]
        tmp[:, xx, :] = np.clip(acc >> P, 0, 255)
    by, ky = pil_bilinear_coeffs(H, out_h)
    out = np.zeros((out_h, out_w, C), np.uint8)
    src = tmp.astype(np.int64)
    for yy in range(out_h):
        ymin, n = by[yy]
        acc = np.full((out_w, C), 1 << (P - 1), np.int64)
        for y in range(n):
            acc += src[ymin + y] * ky[yy, y]
        out[yy] = np.clip(acc >> P, 0, 255)
    return out


def pil_perspective_paste(swapped, orig, coeffs):
    """swapped(RGBA, alpha 255).transform(orig.size, PERSPECTIVE, coeffs, BILINEAR) alpha-composited over orig
    (inference_swap_video.py:715-721).  Inside the warped quad alpha is exactly 255, outside exactly 0, so the composite
    is a select.  swapped [h,w,3] uint8, orig [H,W,3] uint8, coeffs: 8 floats -> (pasted [H,W,3] uint8, inside [H,W])."""
    h, w, _ = swapped.shape
    H, W, _ = orig.shape
    a = [float(v) for v in coeffs]
    ys, xs = np.mgrid[0:H, 0:W]
    xin, yin = xs + 0.5, ys + 0.5
    den = a[6] * xin + a[7] * yin + 1
    sx = (a[0] * xin + a[1] * yin + a[2]) / den
    sy = (a[3] * xin + a[4] * yin + a[5]) / den
    inside = ~((sx < 0.0) | (sx >= w) | (sy < 0.0) | (sy >= h))
    sx, sy = sx - 0.5, sy - 0.5
    fl = lambda v: np.where(v < 0.0, np.floor(v), np.trunc(v)).astype(np.int64)
    x, y = fl(sx), fl(sy)
    dx, dy = sx - x, sy - y
    x0, x1 = np.clip(x, 0, w - 1), np.clip(x + 1, 0, w - 1)
    y0, y1 = np.clip(y, 0, h - 1), np.clip(y + 1, 0, h - 1)
    s = swapped.astype(np.float64)
    v1 = s[y0, x0] + (s[y0, x1] - s[y0, x0]) * dx[..., None]
    v2 = s[y1, x0] + (s[y1, x1] - s[y1, x0]) * dx[..., None]
    v2 = np.where(((y + 1 >= 0) & (y + 1 < h))[..., None], v2, v1)
    proj = (v1 + (v2 - v1) * dy[..., None]).astype(np.uint8)          # (UINT8) v: truncation
    return np.where(inside[..., None], proj, orig), inside


def paste_back(x_sample01, orig_u8, coeffs, up=1024):
    """inference_swap_video.py:702-721 for one frame: x_sample01 [3,h,w] fp32 in [0,1] (clamped decoder output),
    orig_u8 [H,W,3], inverse perspective coefficients -> pasted frame [H,W,3] uint8."""
    x = 255.0 * np.transpose(np.asarray(x_sample01, dtype=np.float32), (1, 2, 0))      # :702, float32 arithmetic
    img = x.astype(np.uint8)                                                         # :706, truncation
    big = pil_resize_bilinear_u8(img, up, up)
    return pil_perspective_paste(big, np.asarray(orig_u8), coeffs)[0]


# --------------------------------------------------------------------------------------------
# specs + whole pipeline (scripts/inference_test_bench.py:438-495)
# --------------------------------------------------------------------------------------------
def _meta(*shape, dtype=torch.float32):
    return torch.empty(*shape, device="meta", dtype=dtype)


def unet_spec(prefix=PFX_UNET, cfg=UNET_CFG):
    P = Params(None, prefix)
    unet_forward(P, _meta(1, cfg["in_channels"], 32, 32), _meta(1, dtype=torch.long), _meta(1, 1, cfg["context_dim"]), cfg)
    return P.spec


def vae_spec(prefix=PFX_VAE, cfg=VAE_CFG):
    P = Params(None, prefix)
    vae_encode_moments(P, _meta(1, 3, 64, 64), cfg)
    vae_decode(P, _meta(1, 4, 8, 8), cfg)
    return P.spec


def clip_spec(prefix=PFX_CLIP, cfg=CLIP_CFG):
    P = Params(None, prefix)
    clip_embed(P, _meta(1, 3, cfg["image"], cfg["image"]), cfg)
    return P.spec


def arcface_spec(prefix=PFX_ARC):
    P = Params(None, prefix)
    with torch.no_grad():
        arcface_backbone(P, _meta(1, 3, 112, 112))
    return P.spec


def parse_spec(prefix=PFX_PARSE):
    P = Params(None, prefix)
    bisenet_logits(P, _meta(1, 3, 64, 64))
    return P.spec


def fusion_spec():
    spec = OrderedDict()
    spec["learnable_vector"] = ((1, 1, 768), "unit")                      # ddpm.py:698
    for n, (i, o) in dict(proj_out_source=(768, 768), proj_out_target=(768, 768), ID_proj_out=(512, 768),
                          landmark_proj_out=(136, 768)).items():
        spec[n + ".weight"] = ((o, i), "w")
        spec[n + ".bias"] = ((o,), "b")
    return spec


def full_spec():
    spec = OrderedDict()
    for s in (unet_spec(), vae_spec(), clip_spec(), arcface_spec(), fusion_spec()):
        spec.update(s)
    return spec


@torch.no_grad()
def swap_pipeline(sd, ref_img, tar_img, inpaint_img, mask_lat, landmarks136, x_T, enc_noise, S=50, scale=3.5,
                  n_steps_limit=None):
    """scripts/inference_test_bench.py:438-495 restated on explicit tensors.
    Returns dict(c, z_inpaint, samples, image)."""
    P = Params(sd)
    b = ref_img.shape[0]
    uc = sd["learnable_vector"].repeat(b, 1, 1)                                 # :441
    c = conditioning_with_feat(P, ref_img, tar_img, landmarks136)              # :447-448
    z_inpaint = vae_encode(P.sub(PFX_VAE), inpaint_img, enc_noise)             # :462-463
    samples, _ = ddim_sample(P.sub(PFX_UNET), x_T, z_inpaint, mask_lat, c, uc, S, scale,
                             n_steps_limit=n_steps_limit)                      # :469-479
    x = vae_decode(P.sub(PFX_VAE), samples)                                     # :493
    return dict(c=c, z_inpaint=z_inpaint, samples=samples, image=torch.clamp((x + 1.0) / 2.0, 0.0, 1.0))


def synthetic_inputs(B, H, seed=42):
    """Synthetic inputs of SURVEY 8(d): target U(-1,1), centred-ellipse mask, ref N(0,1), x_T, noise."""
    g = torch.Generator().manual_seed(seed)
    L = H // 8
    tar = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, H), indexing="ij")
    mask = ((xx / 0.55) ** 2 + (yy / 0.7) ** 2 > 1.0).float()[None, None].repeat(B, 1, 1, 1)
    mask_lat = F.interpolate(mask, size=(L, L), mode="bilinear", align_corners=False)
    ref = torch.randn(B, 3, 224, 224, generator=g)
    x_T = torch.randn(B, 4, L, L, generator=g)
    enc_noise = torch.randn(B, 4, L, L, generator=g)
    return dict(ref_img=ref, tar_img=tar, inpaint_img=tar * mask, mask_lat=mask_lat,
                landmarks136=torch.zeros(B, 136), x_T=x_T, enc_noise=enc_noise)
