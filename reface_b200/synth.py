"""Synthetic (random-init) checkpoints and inputs for benchmarking, generated on the GPU.

The key names / shapes are the reference's state-dict layout (param_manifest.json, derived from
models/REFace/configs/project_ffhq.yaml).  Every tensor that the reference zero-initialises
(`zero_module`, openaimodel.py:229-231,835; attention.py:272) is drawn like any other weight, otherwise
eps == 0 and every block is an identity."""
from __future__ import annotations

import json
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def manifest():
    with open(os.path.join(_HERE, "param_manifest.json")) as f:
        return {k: (tuple(v[0]), v[1]) for k, v in json.load(f).items()}


def flat_layout(spec=None):
    spec = spec or manifest()
    off, lay = 0, {}
    for k, (shape, kind) in spec.items():
        n = 1
        for s in shape:
            n *= s
        lay[k] = (off, n, shape, kind)
        off += (n + 63) // 64 * 64
    return lay, off


def random_flat(device, seed=0):
    """One flat fp32 buffer holding the whole checkpoint (so that a single NCCL broadcast ships it)."""
    lay, total = flat_layout()
    flat = torch.empty(total, dtype=torch.float32, device=device)
    g = torch.Generator(device=device).manual_seed(seed)
    for k, (off, n, shape, kind) in lay.items():
        v = flat[off:off + n]
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            v.normal_(0.0, 1.0 / math.sqrt(max(fan_in, 1)) if len(shape) > 1 else 1.0 / math.sqrt(shape[0]), generator=g)
        elif kind in ("b", "emb"):
            v.normal_(0.0, 0.02, generator=g)
        elif kind == "g":
            v.normal_(1.0, 0.1, generator=g)
        elif kind == "unit":
            v.normal_(0.0, 1.0, generator=g)
        elif kind == "bn_mean":
            v.normal_(0.0, 0.1, generator=g)
        elif kind == "bn_var":
            v.uniform_(0.5, 1.5, generator=g)
        elif kind == "prelu":
            v.normal_(0.25, 0.05, generator=g)
        else:
            v.zero_()
    return flat


def state_dict_from_flat(flat):
    lay, _ = flat_layout()
    return {k: flat[off:off + n].view(shape) for k, (off, n, shape, kind) in lay.items()}


def synthetic_inputs(B, H, device, seed=42, pinned_host=False):
    """target U(-1,1), centred-ellipse mask (bilinear to the latent grid), ref N(0,1), x_T, noise."""
    g = torch.Generator().manual_seed(seed)
    L = H // 8
    tar = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, H), indexing="ij")
    mask = ((xx / 0.55) ** 2 + (yy / 0.7) ** 2 > 1.0).float()[None, None].repeat(B, 1, 1, 1)
    mask_lat = torch.nn.functional.interpolate(mask, size=(L, L), mode="bilinear", align_corners=False)
    d = dict(ref_img=torch.randn(B, 3, 224, 224, generator=g), tar_img=tar, inpaint_img=tar * mask, mask_lat=mask_lat,
             landmarks136=torch.zeros(B, 136), x_T=torch.randn(B, 4, L, L, generator=g),
             enc_noise=torch.randn(B, 4, L, L, generator=g))
    if pinned_host:
        return {k: v.pin_memory() for k, v in d.items()}
    return {k: v.to(device) for k, v in d.items()}
