"""ctypes binding of librefaceb200.so (include/reface_b200.h) with PyTorch tensors at the boundary only.

PyTorch is used here for device memory and streams; every computation happens in the CUDA library.
There is no CPU fallback: importing works anywhere (so that the symbol table can be checked), but creating
an Engine without the built library or without an sm_100a GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librefaceb200.so")
_lib = None

_vp, _f, _i, _ll, _sz = C.c_void_p, C.c_float, C.c_int, C.c_longlong, C.c_size_t

# name -> (restype, argtypes); mirrors include/reface_b200.h one to one
SIGNATURES = {
    "rfb_init": (_i, [_i, _sz, C.POINTER(_vp)]),
    "rfb_destroy": (None, [_vp]),
    "rfb_last_error": (C.c_char_p, [_vp]),
    "rfb_set_param": (_i, [_vp, C.c_char_p, _vp, _i, C.POINTER(C.c_int64)]),
    "rfb_has_param": (_i, [_vp, C.c_char_p]),
    "rfb_release_packed_originals": (_ll, [_vp]),
    "rfb_build_unet": (_i, [_vp, C.c_char_p]),
    "rfb_build_vae": (_i, [_vp, C.c_char_p]),
    "rfb_build_clip": (_i, [_vp, C.c_char_p]),
    "rfb_build_arcface": (_i, [_vp, C.c_char_p]),
    "rfb_build_face_parser": (_i, [_vp, C.c_char_p]),
    "rfb_set_option": (_i, [_vp, C.c_char_p, _ll]),
    "rfb_launch_count": (_ll, [_vp]),
    "rfb_graph_replays": (_ll, [_vp]),
    "rfb_arena_peak": (_sz, [_vp]),
    "rfb_profile_read": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_ll)]),
    "rfb_debug_read": (_i, [_vp, C.POINTER(C.c_ulonglong), _i]),
    "rfb_unet_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "rfb_concat9": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "rfb_cfg_ddim_update": (_i, [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _vp, _vp, _vp]),
    "rfb_ddim_sample": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _i, _vp,
                             _vp, _vp, _vp]),
    "rfb_plms_sample": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _f, _i, _vp, _vp, _vp,
                             _vp]),
    "rfb_q_sample": (_i, [_vp, _vp, _vp, _vp, _i, _ll, _vp, _vp]),
    "rfb_face_parse": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "rfb_inpaint_from_parsing": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "rfb_paste_back": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "rfb_vae_encode": (_i, [_vp, _vp, _vp, _i, _i, _i, C.c_double, _vp, _vp, _vp, _vp]),
    "rfb_vae_decode": (_i, [_vp, _vp, _i, _i, _i, C.c_double, _vp, _vp]),
    "rfb_clip_encode": (_i, [_vp, _vp, _i, _vp, _vp]),
    "rfb_arcface_embed": (_i, [_vp, _vp, _i, _vp, _vp]),
    "rfb_condition_fuse": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _f, _f, _vp, _vp]),
    "rfb_landmark_project": (_i, [_vp, _vp, _i, _vp, _vp]),
    "rfb_target_clip_input": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "rfb_op_linear": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp, _i, _ll, _vp, _vp]),
    "rfb_op_upconv": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "rfb_op_conv2d": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "rfb_op_groupnorm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _i, _i, _vp, _vp]),
    "rfb_op_layernorm": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _f, _vp, _vp]),
    "rfb_op_attention": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "rfb_bench_norm": (_i, [_vp, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_double), _vp]),
}


def load_library(path: str = LIB_PATH):
    """dlopen the C-ABI library and attach the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m reface_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def ddim_schedule(S: int, eta: float = 0.0, linear_start=0.00085, linear_end=0.012, T=1000):
    """Host-side schedule tables, computed exactly like the reference:
    make_beta_schedule / register_schedule (util.py:21-25, ddpm.py:262-275: fp64 -> fp32),
    make_ddim_timesteps (util.py:46-60) and make_ddim_sampling_parameters (util.py:63-74)."""
    betas = (np.linspace(linear_start ** 0.5, linear_end ** 0.5, T, dtype=np.float64) ** 2)
    ac = np.cumprod(1.0 - betas, axis=0).astype(np.float32)
    c = T // S
    ts = np.asarray(list(range(0, T, c))) + 1
    a = ac[ts]
    a_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist(), dtype=np.float64)
    a64 = a.astype(np.float64)
    sig = eta * np.sqrt((1 - a_prev) / (1 - a64) * (1 - a64 / a_prev))
    return dict(timesteps=ts.astype(np.int64), a_t=a.astype(np.float32), a_prev=a_prev.astype(np.float32),
                sigma=np.asarray(sig, dtype=np.float32), sqrt_one_minus_a=np.sqrt(np.float32(1.0) - a).astype(np.float32),
                alphas_cumprod=ac, betas=betas.astype(np.float32))


class Engine:
    """One per GPU: owns the library context (weights, activation arena)."""

    def __init__(self, device: int = 0, arena_bytes: int = 0):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("reface_b200 needs an sm_100a GPU; none is visible (no CPU fallback)")
        self.device = torch.device("cuda", device)
        h = C.c_void_p()
        rc = self.lib.rfb_init(device, arena_bytes, C.byref(h))
        self.h = h
        if rc != 0:
            msg = self.lib.rfb_last_error(h).decode() if h else "rfb_init failed"
            raise RuntimeError(msg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.rfb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.rfb_last_error(self.h).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _in(self, t, dtype=torch.float32):
        if t is None:
            return None
        if t.device != self.device or t.dtype != dtype or not t.is_contiguous():
            t = t.to(device=self.device, dtype=dtype).contiguous()
        return t

    def _new(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd, only_prefix=None):
        """Registers every floating-point entry of a reference-style state dict (host or device tensors)."""
        n = 0
        for k, v in sd.items():
            if only_prefix is not None and not k.startswith(only_prefix):
                continue
            if not torch.is_tensor(v) or not v.is_floating_point():
                continue
            v = v.detach().to(torch.float32).contiguous()
            shape = (C.c_int64 * max(1, v.dim()))(*([int(s) for s in v.shape] or [1]))
            self._ck(self.lib.rfb_set_param(self.h, k.encode(), C.c_void_p(v.data_ptr()), max(1, v.dim()), shape))
            n += 1
        return n

    def release_packed_originals(self):
        """Frees the fp32 originals of weights that now live as packed fp16 operands; returns the bytes freed."""
        return int(self.lib.rfb_release_packed_originals(self.h))

    def has_param(self, name):
        return bool(self.lib.rfb_has_param(self.h, name.encode()))

    def build_unet(self, prefix="model.diffusion_model."):
        self._ck(self.lib.rfb_build_unet(self.h, prefix.encode()))

    def build_vae(self, prefix="first_stage_model."):
        self._ck(self.lib.rfb_build_vae(self.h, prefix.encode()))

    def build_clip(self, prefix="cond_stage_model."):
        self._ck(self.lib.rfb_build_clip(self.h, prefix.encode()))

    def build_arcface(self, prefix="face_ID_model.facenet."):
        self._ck(self.lib.rfb_build_arcface(self.h, prefix.encode()))

    def set_option(self, key, value):
        if self.lib.rfb_set_option(self.h, key.encode(), int(value)) != 0:
            raise KeyError(key)

    @property
    def launch_count(self):
        return int(self.lib.rfb_launch_count(self.h))

    @property
    def graph_replays(self):
        """How many sampling loops ran as ONE CUDA-graph launch (rfb_ddim_sample, option use_graph)."""
        return int(self.lib.rfb_graph_replays(self.h))

    def profile_read(self):
        """(ms, algorithmic_flops, n_launches) of the tensor-core launches since option 'profile' was set; the executed
        FLOPs of the same launches are left in `self.last_executed_flops`."""
        ms, fl, fe, n = C.c_double(), C.c_double(), C.c_double(), _ll()
        self._ck(self.lib.rfb_profile_read(self.h, C.byref(ms), C.byref(fl), C.byref(fe), C.byref(n)))
        self.last_executed_flops = fe.value
        return ms.value, fl.value, n.value

    @property
    def arena_peak(self):
        return int(self.lib.rfb_arena_peak(self.h))

    # ------------------------------------------------------------------ hot path
    def unet_forward(self, x9, t, context):
        x9, context = self._in(x9), self._in(context)
        t = self._in(t, torch.int64)
        N, _, L, _ = x9.shape
        T = context.shape[1]
        eps = self._new(N, 4, L, L)
        if N == 0:          # empty batch: torch modules return empty tensors, so do we
            return eps
        self._ck(self.lib.rfb_unet_forward(self.h, _ptr(x9), _ptr(t), _ptr(context), N, L, T, _ptr(eps), self._stream()))
        return eps

    def concat9(self, x, z, mask, dup=1):
        x, z, mask = self._in(x), self._in(z), self._in(mask)
        B, _, L, _ = x.shape
        out = self._new(dup * B, 9, L, L)
        self._ck(self.lib.rfb_concat9(self.h, _ptr(x), _ptr(z), _ptr(mask), B, L, dup, _ptr(out), self._stream()))
        return out

    def cfg_ddim_update(self, x, eps2, scale, a_t, a_prev, sigma, sqrt_one_minus_at, noise=None, has_uncond=True):
        x, eps2, noise = self._in(x), self._in(eps2), self._in(noise)
        x_prev, p0 = torch.empty_like(x), torch.empty_like(x)
        self._ck(self.lib.rfb_cfg_ddim_update(self.h, _ptr(x), _ptr(eps2), _ptr(noise), x.numel(), float(scale),
                                              float(a_t), float(a_prev), float(sigma), float(sqrt_one_minus_at),
                                              int(has_uncond), _ptr(x_prev), _ptr(p0), self._stream()))
        return x_prev, p0

    def ddim_sample(self, x_T, z_inpaint, mask, cond, uncond, S, scale, eta=0.0, log_every_t=100, noise=None,
                    schedule=None, n_steps_limit=None):
        """Runs the whole DDIM loop on the device.  Returns (x0, x_inter[K,B,4,L,L], pred_x0[K,...])."""
        sch = schedule or ddim_schedule(S, eta)
        x_T, z_inpaint, mask, cond = self._in(x_T), self._in(z_inpaint), self._in(mask), self._in(cond)
        uncond, noise = self._in(uncond), self._in(noise)
        B, _, L, _ = x_T.shape
        T = cond.shape[1]
        ts = np.ascontiguousarray(sch["timesteps"], dtype=np.int64)
        tabs = [np.ascontiguousarray(sch[k], dtype=np.float32) for k in ("a_t", "a_prev", "sigma", "sqrt_one_minus_a")]
        n = len(ts)
        if n_steps_limit is not None:   # run only the first k loop iterations (tests): keep the top-k indices
            k = int(n_steps_limit)
            ts, tabs = ts[n - k:], [t[n - k:] for t in tabs]
            n = k
        idx_logged = [i for i in reversed(range(n)) if log_every_t > 0 and (i % log_every_t == 0 or i == n - 1)]
        K = len(idx_logged)
        x0 = torch.empty_like(x_T)
        inter_x = self._new(max(K, 1), *x_T.shape)
        inter_p = self._new(max(K, 1), *x_T.shape)
        if B == 0:
            return x0, inter_x[:K], inter_p[:K]
        hp = lambda a: a.ctypes.data_as(C.c_void_p)
        self._ck(self.lib.rfb_ddim_sample(self.h, _ptr(x_T), _ptr(z_inpaint), _ptr(mask), _ptr(cond), _ptr(uncond), B, L, T,
                                          hp(ts), hp(tabs[0]), hp(tabs[1]), hp(tabs[2]), hp(tabs[3]), n, float(scale),
                                          _ptr(noise), int(log_every_t), _ptr(x0), _ptr(inter_x), _ptr(inter_p),
                                          self._stream()))
        return x0, inter_x[:K], inter_p[:K]

    def plms_sample(self, x_T, z_inpaint, mask, cond, uncond, S, scale, log_every_t=100, schedule=None, n_steps_limit=None):
        """Runs the whole PLMS loop (plms.py:116-242, eta = 0) on the device; same returns as ddim_sample."""
        sch = schedule or ddim_schedule(S, 0.0)
        x_T, z_inpaint, mask, cond, uncond = (self._in(v) for v in (x_T, z_inpaint, mask, cond, uncond))
        B, _, L, _ = x_T.shape
        T = cond.shape[1]
        ts = np.ascontiguousarray(sch["timesteps"], dtype=np.int64)
        tabs = [np.ascontiguousarray(sch[k], dtype=np.float32) for k in ("a_t", "a_prev", "sigma", "sqrt_one_minus_a")]
        n = len(ts)
        if n_steps_limit is not None:
            k = int(n_steps_limit)
            ts, tabs = ts[n - k:], [t[n - k:] for t in tabs]
            n = k
        K = len([i for i in range(n) if log_every_t > 0 and (i % log_every_t == 0 or i == n - 1)])
        x0 = torch.empty_like(x_T)
        inter_x, inter_p = self._new(max(K, 1), *x_T.shape), self._new(max(K, 1), *x_T.shape)
        if B == 0:
            return x0, inter_x[:K], inter_p[:K]
        hp = lambda a: a.ctypes.data_as(C.c_void_p)
        self._ck(self.lib.rfb_plms_sample(self.h, _ptr(x_T), _ptr(z_inpaint), _ptr(mask), _ptr(cond), _ptr(uncond), B, L, T,
                                          hp(ts), hp(tabs[0]), hp(tabs[1]), hp(tabs[2]), hp(tabs[3]), n, float(scale),
                                          int(log_every_t), _ptr(x0), _ptr(inter_x), _ptr(inter_p), self._stream()))
        return x0, inter_x[:K], inter_p[:K]

    def q_sample(self, x_start, t, noise, linear_start=0.00085, linear_end=0.012, T=1000):
        """DDPM.q_sample (ddpm.py:412-415) with the buffers of register_schedule (fp64 sqrt, cast to fp32)."""
        x_start, noise = self._in(x_start), self._in(noise)
        betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, T, dtype=np.float64) ** 2
        ac = np.cumprod(1.0 - betas, axis=0)
        tt = np.asarray(t.cpu() if torch.is_tensor(t) else t, dtype=np.int64).reshape(-1)
        coef = np.ascontiguousarray(np.stack([np.sqrt(ac)[tt], np.sqrt(1.0 - ac)[tt]], 1), dtype=np.float32)
        B = x_start.shape[0]
        assert len(tt) == B
        out = torch.empty_like(x_start)
        self._ck(self.lib.rfb_q_sample(self.h, _ptr(x_start), _ptr(noise), coef.ctypes.data_as(C.c_void_p), B,
                                       x_start[0].numel(), _ptr(out), self._stream()))
        return out

    def build_face_parser(self, prefix="face_parser.seg."):
        self._ck(self.lib.rfb_build_face_parser(self.h, prefix.encode()))

    def face_parse(self, img01, return_logits=False):
        """BiSeNet face parsing: img01 [B,3,H,W] in [0,1] -> (seg19, seg12) uint8 [B,H,W] (+ logits [B,19,H/8,W/8])."""
        img01 = self._in(img01)
        B, _, H, W = img01.shape
        seg19 = torch.empty(B, H, W, dtype=torch.uint8, device=self.device)
        seg12 = torch.empty(B, H, W, dtype=torch.uint8, device=self.device)
        lg = self._new(B, 19, H // 8, W // 8) if return_logits else None
        self._ck(self.lib.rfb_face_parse(self.h, _ptr(img01), B, H, W, _ptr(lg), _ptr(seg19), _ptr(seg12), self._stream()))
        return (seg19, seg12, lg) if return_logits else (seg19, seg12)

    def inpaint_from_parsing(self, img, seg12, remove=(1, 2, 3, 5, 6, 7, 9)):
        """mask = 1 - isin(seg12, remove); inpaint = img * mask (video_swap_dataset.py:150-222)."""
        img = self._in(img)
        seg12 = seg12.to(self.device, torch.uint8).contiguous()
        B, _, H, W = img.shape
        mask, inp = self._new(B, 1, H, W), torch.empty_like(img)
        rem = (C.c_int * len(remove))(*[int(r) for r in remove])
        self._ck(self.lib.rfb_inpaint_from_parsing(self.h, _ptr(img), _ptr(seg12), rem, len(remove), B, H, W, _ptr(mask),
                                                   _ptr(inp), self._stream()))
        return mask, inp

    def paste_back(self, x01, orig_u8, coeffs, up=1024):
        """inference_swap_video.py:702-721 on the device, bit exact with Pillow: x01 [B,3,h,w] in [0,1], orig_u8 [B,H,W,3]
        uint8, coeffs [B,8] inverse perspective coefficients -> pasted frames [B,H,W,3] uint8."""
        x01 = self._in(x01)
        orig_u8 = orig_u8.to(self.device, torch.uint8).contiguous()
        B, _, h, w = x01.shape
        _, H, W, _ = orig_u8.shape
        co = np.ascontiguousarray(np.asarray(coeffs, dtype=np.float64).reshape(B, 8))
        out = torch.empty_like(orig_u8)
        self._ck(self.lib.rfb_paste_back(self.h, _ptr(x01), _ptr(orig_u8), co.ctypes.data_as(C.c_void_p), B, h, w, int(up),
                                         H, W, _ptr(out), self._stream()))
        return out

    def vae_encode(self, img, noise=None, return_moments=False, scale_factor=0.18215):
        """z = scale_factor * (mean + std * noise) (ddpm.py:850-857); scale_factor=1.0 is posterior.sample() itself."""
        img, noise = self._in(img), self._in(noise)
        B, _, H, W = img.shape
        z, mean, logvar = self._new(B, 4, H // 8, W // 8), self._new(B, 4, H // 8, W // 8), self._new(B, 4, H // 8, W // 8)
        if B == 0:
            return (z, mean, logvar) if return_moments else z
        self._ck(self.lib.rfb_vae_encode(self.h, _ptr(img), _ptr(noise), B, H, W, float(scale_factor), _ptr(z), _ptr(mean),
                                         _ptr(logvar), self._stream()))
        return (z, mean, logvar) if return_moments else z

    def vae_decode(self, z, scale_factor=0.18215):
        """decode_first_stage (ddpm.py:1277-1337): decoder(1/scale_factor * z); scale_factor=1.0 is AutoencoderKL.decode."""
        z = self._in(z)
        if z.shape[1] != 4:
            z = z[:, :4].contiguous()   # ddpm.py:1334-1335
        B, _, h, w = z.shape
        img = self._new(B, 3, 8 * h, 8 * w)
        if B == 0:
            return img
        self._ck(self.lib.rfb_vae_decode(self.h, _ptr(z), B, h, w, float(scale_factor), _ptr(img), self._stream()))
        return img

    def clip_encode(self, img224):
        img224 = self._in(img224)
        B = img224.shape[0]
        out = self._new(B, 1, 768)
        if B == 0:
            return out
        self._ck(self.lib.rfb_clip_encode(self.h, _ptr(img224), B, _ptr(out), self._stream()))
        return out

    def arcface_embed(self, img224):
        img224 = self._in(img224)
        B = img224.shape[0]
        out = self._new(B, 512)
        if B == 0:
            return out
        self._ck(self.lib.rfb_arcface_embed(self.h, _ptr(img224), B, _ptr(out), self._stream()))
        return out

    def target_clip_input(self, tar):
        tar = self._in(tar)
        B, _, H, W = tar.shape
        out = self._new(B, 3, 224, 224)
        if B == 0:
            return out
        self._ck(self.lib.rfb_target_clip_input(self.h, _ptr(tar), B, H, W, _ptr(out), self._stream()))
        return out

    def condition_fuse(self, clip_src, clip_tgt, id_feat, lm136=None, w_clip=1.0, w_id=10.0, w_lm=0.05, lm_proj=None):
        """The landmark term is either raw points `lm136` [B,136] or the already projected `lm_proj` [B,768] (what
        the reference's get_landmarks returns, ddpm.py:1096)."""
        clip_src, clip_tgt = self._in(clip_src).reshape(-1, 768), self._in(clip_tgt).reshape(-1, 768)
        id_feat, lm136 = self._in(id_feat), self._in(lm136)
        lm_proj = None if lm_proj is None else self._in(lm_proj).reshape(-1, 768)
        B = clip_src.shape[0]
        out = self._new(B, 1, 768)
        if B == 0:
            return out
        self._ck(self.lib.rfb_condition_fuse(self.h, _ptr(clip_src), _ptr(clip_tgt), _ptr(id_feat), _ptr(lm136),
                                             _ptr(lm_proj), B, float(w_clip), float(w_id), float(w_lm), _ptr(out),
                                             self._stream()))
        return out

    def landmark_project(self, lm136):
        """landmark_proj_out(lm136): [B,136] raw dlib points (zeros = no face) -> [B,768] (ddpm.py:1081-1096)."""
        lm136 = self._in(lm136).reshape(-1, 136)
        out = self._new(lm136.shape[0], 768)
        if lm136.shape[0] == 0:
            return out
        self._ck(self.lib.rfb_landmark_project(self.h, _ptr(lm136), lm136.shape[0], _ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ single ops (tests)
    def op_linear(self, x, w, bias=None, residual=None, act=0, geglu=False, x2=None):
        """x2 [M2,K2]: out = [x | x2] w^T with x2 read at row (m mod M2) (two-descriptor K loop, no concatenation)."""
        x, w, bias, residual, x2 = self._in(x), self._in(w), self._in(bias), self._in(residual), self._in(x2)
        M, K = x.shape
        N = w.shape[0]
        M2, K2 = (x2.shape if x2 is not None else (0, 0))
        out = self._new(M, N // 2 if geglu else N)
        self._ck(self.lib.rfb_op_linear(self.h, _ptr(x), _ptr(w), _ptr(bias), _ptr(residual), M, K, N, int(act),
                                        int(geglu), _ptr(x2), K2, M2, _ptr(out), self._stream()))
        return out

    def op_upconv(self, x, w, bias=None):
        """conv3x3(nearest_2x(x)) through the folded four-phase kernel."""
        x, w, bias = self._in(x), self._in(w), self._in(bias)
        N, Cc, H, W = x.shape
        O = w.shape[0]
        out = self._new(N, O, 2 * H, 2 * W)
        self._ck(self.lib.rfb_op_upconv(self.h, _ptr(x), _ptr(w), _ptr(bias), N, Cc, H, W, O, _ptr(out), self._stream()))
        return out

    def op_conv2d(self, x, w, bias=None, stride=1, pad=(1, 1, 1, 1), gn=None):
        """gn=(gamma, beta): the convolution is followed by GroupNorm(32, eps 1e-5) + SiLU fed by epilogue statistics."""
        x, w, bias = self._in(x), self._in(w), self._in(bias)
        gg, gb = (self._in(gn[0]), self._in(gn[1])) if gn is not None else (None, None)
        N, Cc, H, W = x.shape
        O, _, k, _ = w.shape
        pt, pl, pb, pr = pad
        Ho, Wo = (H + pt + pb - k) // stride + 1, (W + pl + pr - k) // stride + 1
        out = self._new(N, O, Ho, Wo)
        self._ck(self.lib.rfb_op_conv2d(self.h, _ptr(x), _ptr(w), _ptr(bias), N, Cc, H, W, O, k, stride, pt, pl, pb, pr,
                                        _ptr(gg), _ptr(gb), _ptr(out), self._stream()))
        return out

    def op_groupnorm(self, x, gamma, beta, eps, silu=False, x2=None):
        """x2 [N2,C2,H,W]: GroupNorm(32) over the channel concatenation [x | x2] (sample n reads x2[n mod N2])."""
        x, gamma, beta, x2 = self._in(x), self._in(gamma), self._in(beta), self._in(x2)
        N, Cc, H, W = x.shape
        N2, C2 = (x2.shape[:2] if x2 is not None else (0, 0))
        out = self._new(N, Cc + C2, H, W)
        self._ck(self.lib.rfb_op_groupnorm(self.h, _ptr(x), _ptr(gamma), _ptr(beta), N, Cc, H, W, float(eps), int(silu),
                                           _ptr(x2), C2, N2, _ptr(out), self._stream()))
        return out

    def op_layernorm(self, x, gamma, beta, eps=1e-5):
        x, gamma, beta = self._in(x), self._in(gamma), self._in(beta)
        rows, Cc = x.shape
        out = torch.empty_like(x)
        self._ck(self.lib.rfb_op_layernorm(self.h, _ptr(x), _ptr(gamma), _ptr(beta), rows, Cc, float(eps), _ptr(out),
                                           self._stream()))
        return out

    def bench_norm(self, kind, N, Cc, H, W, iters=20):
        """ms per launch of GroupNorm+SiLU (kind 0) / LayerNorm (kind 1) on a device-resident fp16 [N,H,W,C] tensor."""
        ms = C.c_double(0.0)
        self._ck(self.lib.rfb_bench_norm(self.h, int(kind), N, Cc, H, W, iters, C.byref(ms), self._stream()))
        return ms.value

    def op_attention(self, qkv, heads, scale=None):
        qkv = self._in(qkv)
        N, L, C3 = qkv.shape
        d = C3 // 3 // heads
        out = self._new(N, L, C3 // 3)
        self._ck(self.lib.rfb_op_attention(self.h, _ptr(qkv), N, L, heads, d, float(scale if scale else d ** -0.5),
                                           _ptr(out), self._stream()))
        return out
