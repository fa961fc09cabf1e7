"""Drop-in API surface: the reference's Python signatures on top of the CUDA engine.

The reference's plugin mechanism is `instantiate_from_config` (ldm/util.py:78-93): YAML `target:` strings
name classes.  Pointing those strings (and the `from ldm.models.diffusion.ddim import DDIMSampler` import of
scripts/inference_test_bench.py:339) at this module swaps the hot path for the sm_100a kernels while the
driver script stays as it is:

    unet_config.target         ldm.modules.diffusionmodules.openaimodel.UNetModel   -> reface_b200.ldm_api.UNetModel
    first_stage_config.target  ldm.models.autoencoder.AutoencoderKL                  -> reface_b200.ldm_api.AutoencoderKL
    cond_stage_config.target   ldm.modules.encoders.modules.FrozenCLIPEmbedder       -> reface_b200.ldm_api.FrozenCLIPEmbedder
    model.target               ldm.models.diffusion.ddpm.LatentDiffusion             -> reface_b200.ldm_api.LatentDiffusion

Only tensors cross the boundary (PyTorch CUDA tensors in/out); nothing here computes on the CPU.
"""
from __future__ import annotations

import contextlib
import sys
import warnings

import numpy as np
import torch

from .runtime import Engine, ddim_schedule

_ENGINES = {}


def get_engine(device: int = 0, arena_bytes: int = 0) -> Engine:
    if device not in _ENGINES:
        _ENGINES[device] = Engine(device, arena_bytes)
    return _ENGINES[device]


class _Module(torch.nn.Module):
    """nn.Module shell around engine weights.  It holds no torch parameters: `load_state_dict` -- its own, or the one of
    a PARENT module such as the reference's LatentDiffusion (scripts/inference_test_bench.py:98-103), which reaches
    this class through nn.Module._load_from_state_dict -- registers the tensors with the engine under the reference's
    canonical key names and packs the network.  eval / cuda / to / half are the inherited nn.Module no-ops."""
    canonical = ""          # engine-side key prefix, e.g. "model.diffusion_model."

    def __init__(self, engine=None, device=0):
        super().__init__()
        self._engine, self._device_index = engine, device
        self._built = False

    @property
    def engine(self):       # created on first use: constructing the shell (instantiate_from_config) needs no GPU
        if self._engine is None:
            self._engine = get_engine(self._device_index)
        return self._engine

    @property
    def device(self):
        return self.engine.device

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        mine = {self.canonical + k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        if not mine:
            if strict:
                missing_keys.append(prefix + "*")
            return
        self.engine.load_state_dict(mine)
        try:
            self._build()
            self._built = True
        except RuntimeError as e:      # e.g. "missing parameter: ..." from the packer
            error_msgs.append(f"{type(self).__name__}: {e}")

    def _require_built(self):
        if not self._built:
            raise RuntimeError(f"{type(self).__name__}: weights have not been loaded (load_state_dict) yet")


class UNetModel(_Module):
    """ldm/modules/diffusionmodules/openaimodel.py:528-907 (constructor args :558-588, forward :860)."""
    canonical = "model.diffusion_model."

    def __init__(self, image_size=32, in_channels=9, model_channels=320, out_channels=4, num_res_blocks=2,
                 attention_resolutions=(4, 2, 1), channel_mult=(1, 2, 4, 4), num_heads=8, use_spatial_transformer=True,
                 transformer_depth=1, context_dim=768, use_checkpoint=False, legacy=False,
                 add_conv_in_front_of_unet=False, engine=None, device=0, **unused):
        super().__init__(engine, device)
        cfg = (in_channels, out_channels, model_channels, num_res_blocks, tuple(attention_resolutions),
               tuple(channel_mult), num_heads, context_dim, transformer_depth, use_spatial_transformer, legacy,
               add_conv_in_front_of_unet)
        if cfg != (9, 4, 320, 2, (4, 2, 1), (1, 2, 4, 4), 8, 768, 1, True, False, False):
            raise NotImplementedError(f"reface_b200 implements the shipped REFace UNet config only, got {cfg}")
        self.in_channels, self.out_channels, self.model_channels = in_channels, out_channels, model_channels
        self.dtype = torch.float32

    def _build(self):
        self.engine.build_unet(self.canonical)

    def forward(self, x, timesteps=None, context=None, y=None, return_features=False, **kwargs):
        assert y is None, "must specify y if and only if the model is class-conditional"   # openaimodel.py:870-872
        if return_features:
            raise NotImplementedError("return_features is a training-time option")
        self._require_built()
        return self.engine.unet_forward(x, timesteps, context)


class _Posterior:
    """DiagonalGaussianDistribution (ldm/modules/distributions/distributions.py:24-61) over engine outputs."""

    def __init__(self, engine, img):
        self.engine, self.img = engine, img
        self._moments = None
        self.deterministic = False

    def _m(self):
        if self._moments is None:
            _, mean, logvar = self.engine.vae_encode(self.img, None, return_moments=True)
            self._moments = (mean, logvar)
        return self._moments

    mean = property(lambda self: self._m()[0])
    logvar = property(lambda self: self._m()[1])
    std = property(lambda self: torch.exp(0.5 * self._m()[1]))
    var = property(lambda self: torch.exp(self._m()[1]))

    def sample(self, noise=None, scale_factor=1.0):
        """mean + std * randn (distributions.py:35-37); `scale_factor` folds LatentDiffusion.scale_factor into the same
        kernel (one fp32 multiply, as ddpm.py:857)."""
        if noise is None:
            b, _, h, w = self.img.shape
            noise = torch.randn(b, 4, h // 8, w // 8, device=self.engine.device)
        return self.engine.vae_encode(self.img, noise, scale_factor=scale_factor)

    def mode(self):
        return self.mean


_POSTERIOR_CLS = {}


def _posterior_cls():
    """Inside the reference tree the posterior must pass `isinstance(.., DiagonalGaussianDistribution)`
    (get_first_stage_encoding, ddpm.py:850-857): when the reference package is loaded in this process the returned class
    also derives from its DiagonalGaussianDistribution; stand-alone it is plain _Posterior."""
    mod = sys.modules.get("ldm.modules.distributions.distributions")
    base = getattr(mod, "DiagonalGaussianDistribution", None) if mod is not None else None
    if base not in _POSTERIOR_CLS:
        _POSTERIOR_CLS[base] = _Posterior if base is None else type("_PosteriorRef", (_Posterior, base), {})
    return _POSTERIOR_CLS[base]


class AutoencoderKL(_Module):
    """ldm/models/autoencoder.py:285-333: encode(x) -> posterior, decode(z) -> image."""
    canonical = "first_stage_model."

    def __init__(self, ddconfig=None, lossconfig=None, embed_dim=4, engine=None, device=0, **unused):
        super().__init__(engine, device)
        self.embed_dim = embed_dim

    def _build(self):
        self.engine.build_vae(self.canonical)

    def encode(self, x):
        self._require_built()
        return _posterior_cls()(self.engine, x)

    def decode(self, z):
        self._require_built()
        return self.engine.vae_decode(z, scale_factor=1.0)   # the caller has applied 1/scale_factor (ddpm.py:1284)

    def forward(self, input, sample_posterior=True):                      # autoencoder.py:335-342
        posterior = self.encode(input)
        z = posterior.sample() if sample_posterior else posterior.mode()
        return self.decode(z), posterior


class FrozenCLIPEmbedder(_Module):
    """ldm/modules/encoders/modules.py:211-264: encode(image[B,3,224,224]) -> [B,1,768]."""
    canonical = "cond_stage_model."

    def __init__(self, version="openai/clip-vit-large-patch14", engine=None, device=0, **unused):
        super().__init__(engine, device)

    def _build(self):
        self.engine.build_clip(self.canonical)

    def encode(self, image):
        self._require_built()
        return self.engine.clip_encode(image)

    def forward(self, image):
        return self.encode(image)


class LatentDiffusion:
    """The slice of ldm/models/diffusion/ddpm.py:LatentDiffusion that the inference scripts touch
    (scripts/inference_test_bench.py:408-493, ldm/models/diffusion/ddim.py:100,113-119,207,345).

    `landmark_detector`: optional callable(uint8 image [H,W,3]) -> 68x2 landmark array or None (no face).  dlib is
    outside the boundary: pass e.g. `lambda im: dlib_points(detector, predictor, im)` to reproduce
    get_landmarks (ddpm.py:1068-1099).  Without one, get_landmarks warns once and takes the reference's no-face branch
    (zeros(136) -> landmark_proj_out, ddpm.py:1081-1083) for every image."""

    def __init__(self, state_dict=None, engine=None, device=0, scale_factor=0.18215, linear_start=0.00085,
                 linear_end=0.012, timesteps=1000, clip_weight=1.0, ID_weight=10.0, Landmarks_weight=0.05,
                 landmark_detector=None, **unused):
        self.engine = engine or get_engine(device)
        self.device = self.engine.device
        self.scale_factor = float(scale_factor)
        self.num_timesteps = timesteps
        sch = ddim_schedule(50, 0.0, linear_start, linear_end, timesteps)
        ac = sch["alphas_cumprod"]
        self.betas = torch.tensor(sch["betas"], device=self.device)
        self.alphas_cumprod = torch.tensor(ac, device=self.device)
        self.alphas_cumprod_prev = torch.tensor(np.append(np.float32(1.0), ac[:-1]), device=self.device)
        self.linear_start, self.linear_end = linear_start, linear_end
        self.clip_weight, self.ID_weight, self.Landmarks_weight = clip_weight, ID_weight, Landmarks_weight
        self.stack_feat = self.land_mark_id_seperate_layers = self.sep_head_att = False
        self.Landmark_cond = True
        self.Landmark_loss_weight = 0.0
        self.landmark_detector = landmark_detector
        self.learnable_vector = None
        self._warned_no_detector = False
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # -- weights
    def load_state_dict(self, sd, strict=False, release_originals=True):
        e = self.engine
        e.load_state_dict(sd)
        e.build_unet()
        e.build_vae()
        e.build_clip()
        e.build_arcface()
        if release_originals:
            e.release_packed_originals()      # keep the packed fp16 operands only (~2.7 GB instead of ~8 GB)
        self.learnable_vector = sd["learnable_vector"].to(self.device, torch.float32)
        return [], []

    # nn.Module conveniences the inference scripts call (inference_test_bench.py:111,334; inference_swap_video.py)
    def eval(self):
        return self

    def train(self, mode=True):
        return self

    def cuda(self, *a, **k):
        return self

    def to(self, *a, **k):
        return self

    def half(self):
        return self

    def float(self):
        return self

    def requires_grad_(self, *a, **k):
        return self

    def parameters(self):
        return iter(())

    @contextlib.contextmanager
    def ema_scope(self, context=None):       # use_ema: false (project_ffhq.yaml:19) -> no-op (ddpm.py:309-322)
        yield None

    # -- hot path
    def apply_model(self, x_noisy, t, cond, return_ids=False):             # ddpm.py:1519-1617, 2244-2246
        if isinstance(cond, dict):
            cond = torch.cat(cond["c_crossattn"], 1)
        elif isinstance(cond, (list, tuple)):
            cond = torch.cat(list(cond), 1)
        return self.engine.unet_forward(x_noisy, t, cond)

    def encode_first_stage(self, x):                                       # ddpm.py:1402-1439
        return _Posterior(self.engine, x)

    def get_first_stage_encoding(self, posterior, noise=None):             # ddpm.py:850-857
        if isinstance(posterior, _Posterior):
            return posterior.sample(noise, scale_factor=self.scale_factor)
        return self.scale_factor * posterior

    def q_sample(self, x_start, t, noise=None):                            # ddpm.py:412-415
        if noise is None:
            noise = torch.randn_like(x_start)
        return self.engine.q_sample(x_start, t, noise, self.linear_start, self.linear_end, self.num_timesteps)

    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):   # ddpm.py:1277-1337
        return self.engine.vae_decode(z, scale_factor=self.scale_factor)

    def get_learned_conditioning(self, c):                                 # ddpm.py:859-870
        return self.engine.clip_encode(c)

    def get_landmarks(self, x, landmarks136=None):
        """ddpm.py:1068-1099: x [B,3,H,W] in [-1,1] -> projected landmarks [B,768].  The uint8 conversion and the
        68-point layout follow the reference; the detector itself is the caller's (`landmark_detector`), or raw points
        can be handed in directly as `landmarks136` [B,136]."""
        b = x.shape[0]
        if landmarks136 is None:
            if self.landmark_detector is None:
                if not self._warned_no_detector:
                    warnings.warn("reface_b200.LatentDiffusion.get_landmarks: no landmark_detector configured -- every "
                                  "image takes the reference's no-face branch (zeros(136), ddpm.py:1081-1083); pass "
                                  "landmark_detector=... or landmarks136=... to condition on detected landmarks",
                                  RuntimeWarning, stacklevel=2)
                    self._warned_no_detector = True
                landmarks136 = torch.zeros(b, 136)
            else:
                im = (255.0 * ((x + 1.0) / 2.0).permute(0, 2, 3, 1).cpu().numpy()).astype(np.uint8)   # ddpm.py:1077-1078
                rows = []
                for i in range(b):
                    pts = self.landmark_detector(im[i])
                    rows.append(np.zeros((1, 136)) if pts is None else np.asarray(pts).reshape(1, 136))
                landmarks136 = torch.tensor(np.concatenate(rows, 0)).float()
        return self.engine.landmark_project(landmarks136)                  # ddpm.py:1096

    def conditioning_with_feat(self, x, landmarks=None, is_train=False, tar=None, tar_mask=None, landmarks136=None):
        """ddpm.py:872-1045 for the shipped config (Source+Target CLIP, ArcFace ID, landmarks, weight_division).
        `landmarks` is what the reference passes: the PROJECTED landmarks [B,768] from get_landmarks
        (scripts/inference_test_bench.py:447-448); raw 68x2 points may be given as `landmarks136` instead.  With
        neither, the no-face branch (zeros(136)) is used."""
        e = self.engine
        b = x.shape[0]
        # one CLIP pass over [source ; resized target] (the reference encodes them separately, ddpm.py:884,913: same
        # values, every sample is independent)
        c_both = e.clip_encode(torch.cat([x, e.target_clip_input(tar)], 0))
        c_src, c_tgt = c_both[:b], c_both[b:]
        idf = e.arcface_embed(x)
        return self._fuse(c_src, c_tgt, idf, landmarks, landmarks136, b)

    def _fuse(self, c_src, c_tgt, idf, landmarks, landmarks136, b):
        if landmarks is not None and landmarks136 is not None:
            raise ValueError("pass either projected `landmarks` [B,768] or raw `landmarks136` [B,136], not both")
        if landmarks is not None:
            return self.engine.condition_fuse(c_src, c_tgt, idf, None, self.clip_weight, self.ID_weight, self.Landmarks_weight,
                                              lm_proj=landmarks.reshape(b, 768))
        if landmarks136 is None:
            landmarks136 = torch.zeros(b, 136, device=self.device)
        return self.engine.condition_fuse(c_src, c_tgt, idf, landmarks136, self.clip_weight, self.ID_weight,
                                          self.Landmarks_weight)

    def source_features(self, x):
        """CLIP embedding and ArcFace identity of the source face(s): the part of conditioning_with_feat that depends
        on the source only.  Video mode computes it once and reuses it for every frame (the reference recomputes it
        for every batch from the repeated source image, scripts/inference_swap_video.py:627-632; same values)."""
        return self.engine.clip_encode(x), self.engine.arcface_embed(x)

    def conditioning_from_source_features(self, src_feats, tar, landmarks136=None, landmarks=None):
        """conditioning_with_feat (ddpm.py:872-1045) with the source terms taken from source_features()."""
        e = self.engine
        b = tar.shape[0]
        c_src, idf = src_feats
        if c_src.shape[0] == 1 and b > 1:
            c_src, idf = c_src.repeat(b, 1, 1), idf.repeat(b, 1)
        c_tgt = e.clip_encode(e.target_clip_input(tar))
        return self._fuse(c_src, c_tgt, idf, landmarks, landmarks136, b)


class DDIMSampler:
    """ldm/models/diffusion/ddim.py:96-251: same constructor and `sample` signature; the 50-step loop runs inside
    one C-ABI call (rfb_ddim_sample)."""

    def __init__(self, model, schedule="linear", **kwargs):
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True):
        assert ddim_discretize == "uniform"
        sch = ddim_schedule(ddim_num_steps, ddim_eta, self.model.linear_start, self.model.linear_end,
                            self.ddpm_num_timesteps)
        self._sch = sch
        self.ddim_timesteps = sch["timesteps"]
        dev = self.model.device
        self.ddim_alphas = torch.tensor(sch["a_t"], device=dev)
        self.ddim_alphas_prev = sch["a_prev"]
        self.ddim_sigmas = torch.tensor(sch["sigma"], device=dev)
        self.ddim_sqrt_one_minus_alphas = torch.tensor(sch["sqrt_one_minus_a"], device=dev)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0.0, mask=None, x0=None, temperature=1.0, noise_dropout=0.0, score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.0,
               unconditional_conditioning=None, src_im=None, tar=None, **kwargs):
        if conditioning is not None and not isinstance(conditioning, dict) and conditioning.shape[0] != batch_size:
            print(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")   # ddim.py:172-173
        if "test_model_kwargs" in kwargs:
            tk = kwargs["test_model_kwargs"]
            z_inpaint, m = tk["inpaint_image"], tk["inpaint_mask"]
        elif "rest" in kwargs:
            rest = kwargs["rest"]
            z_inpaint, m = rest[:, :4], rest[:, 4:5]
        else:
            raise Exception("kwargs must contain either 'test_model_kwargs' or 'rest' key")           # ddim.py:333-334
        for name, v in dict(mask=mask, score_corrector=score_corrector, callback=callback, img_callback=img_callback).items():
            if v is not None:
                raise NotImplementedError(f"{name} is not supported by the fused DDIM loop")
        if quantize_x0 or noise_dropout > 0.0:
            raise NotImplementedError("quantize_x0 / noise_dropout are not supported")
        self.make_schedule(S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        eng = self.model.engine
        dev = self.model.device
        if x_T is None:
            x_T = torch.randn(batch_size, C, H, W, device=dev)                                        # ddim.py:210-213
        n = len(self.ddim_timesteps)
        noise = None
        if eta > 0:
            noise = torch.randn(n, batch_size, C, H, W, device=dev) * temperature                       # ddim.py:371
        x0_, ix, ip = eng.ddim_sample(x_T, z_inpaint, m, conditioning, unconditional_conditioning, S,
                                      unconditional_guidance_scale, eta=eta, log_every_t=log_every_t, noise=noise,
                                      schedule=self._sch)
        inter = {"x_inter": [x_T] + list(ix), "pred_x0": [x_T] + list(ip)}                             # ddim.py:221,247-249
        return x0_, inter


class PLMSSampler(DDIMSampler):
    """ldm/models/diffusion/plms.py:11-242: same constructor and `sample` signature as the reference's PLMSSampler;
    the loop (pseudo improved Euler + Adams-Bashforth over the last three eps) runs inside one C-ABI call
    (rfb_plms_sample)."""

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True):
        if ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")                                            # plms.py:25-26
        super().make_schedule(ddim_num_steps, ddim_discretize, ddim_eta, verbose)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0.0, mask=None, x0=None, temperature=1.0, noise_dropout=0.0, score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.0,
               unconditional_conditioning=None, **kwargs):
        if conditioning is not None and not isinstance(conditioning, dict) and conditioning.shape[0] != batch_size:
            print(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")   # plms.py:89-90
        tk = kwargs["test_model_kwargs"]                                                               # plms.py:218 (KeyError as there)
        for name, v in dict(mask=mask, score_corrector=score_corrector, callback=callback, img_callback=img_callback).items():
            if v is not None:
                raise NotImplementedError(f"{name} is not supported by the fused PLMS loop")
        if quantize_x0 or noise_dropout > 0.0:
            raise NotImplementedError("quantize_x0 / noise_dropout are not supported")
        self.make_schedule(S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        if x_T is None:
            x_T = torch.randn(batch_size, C, H, W, device=self.model.device)                           # plms.py:125-126
        x0_, ix, ip = self.model.engine.plms_sample(x_T, tk["inpaint_image"], tk["inpaint_mask"], conditioning,
                                                    unconditional_conditioning, S, unconditional_guidance_scale,
                                                    log_every_t=log_every_t, schedule=self._sch)
        return x0_, {"x_inter": [x_T] + list(ix), "pred_x0": [x_T] + list(ip)}                          # plms.py:136,168-170


def swap_faces(model: LatentDiffusion, ref_img, tar_img, inpaint_img, mask_lat, landmarks136, x_T, enc_noise, S=50,
               scale=3.5, log_every_t=100, landmarks=None):
    """The per-batch body of scripts/inference_test_bench.py:438-495 through the drop-in classes.  `landmarks` (projected,
    [B,768], as model.get_landmarks returns them) may be given instead of the raw `landmarks136`."""
    b = ref_img.shape[0]
    uc = model.learnable_vector.repeat(b, 1, 1)                                                      # :441
    if landmarks is not None:
        c = model.conditioning_with_feat(ref_img, landmarks=landmarks, tar=tar_img)                   # :447-448
    else:
        c = model.conditioning_with_feat(ref_img, tar=tar_img, landmarks136=landmarks136)
    z_inpaint = model.get_first_stage_encoding(model.encode_first_stage(inpaint_img), noise=enc_noise)   # :462-463
    sampler = DDIMSampler(model)
    samples, inter = sampler.sample(S=S, conditioning=c, batch_size=b, shape=[4, x_T.shape[2], x_T.shape[3]],
                                    verbose=False, unconditional_guidance_scale=scale, unconditional_conditioning=uc,
                                    eta=0.0, x_T=x_T, log_every_t=log_every_t,
                                    test_model_kwargs={"inpaint_image": z_inpaint, "inpaint_mask": mask_lat})   # :469-479
    x = model.decode_first_stage(samples)                                                              # :493
    return dict(c=c, z_inpaint=z_inpaint, samples=samples, intermediates=inter,
                image=torch.clamp((x + 1.0) / 2.0, 0.0, 1.0))


def swap_video(model: LatentDiffusion, ref_img, tar_img, inpaint_img, mask_lat, x_T, enc_noise, landmarks136=None, S=30,
               scale=3.0, chunk=30, rank=0, world=1):
    """Video mode (BASELINE configs[4]; scripts/inference_swap_video.py:560-724 with inference_video_swap.sh:28-29:
    30 DDIM steps -> 31 timesteps, scale 3): ONE source face [1,3,224,224], F target frames [F,...].  The frames are
    streamed in contiguous chunks (reface_b200.shard.video_segments: chunk k -> rank k % world), the source CLIP /
    ArcFace features are computed once, the target CLIP embedding per frame.  Returns {frame_index: image[3,H,W]} for
    the frames of this rank; every frame's result is bitwise the one swap_faces gives for it (no cross-sample
    arithmetic anywhere in the path)."""
    from .shard import video_segments
    F_ = tar_img.shape[0]
    feats = model.source_features(ref_img[:1])
    sampler = DDIMSampler(model)
    out = {}
    for lo, hi in video_segments(F_, chunk, rank, world):
        b = hi - lo
        uc = model.learnable_vector.repeat(b, 1, 1)
        lm = None if landmarks136 is None else landmarks136[lo:hi]
        c = model.conditioning_from_source_features(feats, tar_img[lo:hi], lm)
        z = model.get_first_stage_encoding(model.encode_first_stage(inpaint_img[lo:hi]), noise=enc_noise[lo:hi])
        smp, _ = sampler.sample(S=S, conditioning=c, batch_size=b, shape=[4, x_T.shape[2], x_T.shape[3]], verbose=False,
                                unconditional_guidance_scale=scale, unconditional_conditioning=uc, eta=0.0, x_T=x_T[lo:hi],
                                test_model_kwargs={"inpaint_image": z, "inpaint_mask": mask_lat[lo:hi]})
        img = torch.clamp((model.decode_first_stage(smp) + 1.0) / 2.0, 0.0, 1.0)
        for i in range(b):
            out[lo + i] = img[i]
    return out


def paste_back(model: LatentDiffusion, images, orig_frames_u8, inv_coeffs, up=1024):
    """scripts/inference_swap_video.py:702-721 for a batch of frames, on the device and bit exact with Pillow:
    images [B,3,h,w] in [0,1] (swap_faces / swap_video outputs), orig_frames_u8 [B,H,W,3] uint8, inv_coeffs [B,8] (the
    per-frame inverse perspective coefficients of crop_and_align_face, :71-103,713) -> pasted frames [B,H,W,3] uint8."""
    return model.engine.paste_back(images, orig_frames_u8, inv_coeffs, up=up)
