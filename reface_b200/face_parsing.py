"""Drop-in surface for the reference's face parser (pretrained/face_parsing/face_parsing_demo.py:236-318) on the
sm_100a kernels: BiSeNet (ResNet-18 context path) -> 19-class label map -> 12-class map -> inpainting mask / image.
PyTorch is used for device memory only; there is no CPU fallback."""
from __future__ import annotations

import numpy as np
import torch

from .runtime import Engine

PREFIX = "face_parser.seg."
REMOVE_MASK_TAR_FFHQ = (1, 2, 3, 5, 6, 7, 9)      # models/REFace/configs/project_ffhq.yaml:209-216


class FaceParser:
    """face_parsing_demo.py:236-281.  `seg_ckpt` is the BiSeNet state dict (the keys of 79999_iter.pth, or already
    prefixed with "face_parser.seg.") or a path to it; `forward(img)` returns the 19-class label map."""

    def __init__(self, seg_ckpt, size=1024, device=0, engine: Engine | None = None):
        self.engine = engine or Engine(device if isinstance(device, int) else 0)
        sd = torch.load(seg_ckpt, map_location="cpu") if isinstance(seg_ckpt, str) else seg_ckpt
        sd = {(k if k.startswith(PREFIX) else PREFIX + k): v for k, v in sd.items()
              if not k.endswith("num_batches_tracked") and ".conv_out16." not in "." + k and ".conv_out32." not in "." + k}
        self.engine.load_state_dict(sd)
        self.engine.build_face_parser(PREFIX)
        self.size = size

    def preprocess_img(self, img):
        """PIL image / uint8 array [H,W,3] / float tensor [.,3,H,W] in [0,1] -> float tensor [B,3,H,W] on the device.
        Resampling to 512 (BicubicDownSample / PIL BILINEAR, face_parsing_demo.py:262-267) stays with the caller: the
        kernels take the 512-sized image; clamp and normalisation happen on the device."""
        if not torch.is_tensor(img):
            a = np.asarray(img)
            img = torch.from_numpy(a[..., :3].copy()).permute(2, 0, 1).float().div(255.0)
        if img.dim() == 3:
            img = img[None]
        return img.to(self.engine.device, torch.float32)

    @torch.no_grad()
    def forward(self, img):
        seg19, _ = self.engine.face_parse(self.preprocess_img(img))
        return seg19[0].long() if seg19.shape[0] == 1 else seg19.long()

    __call__ = forward


def faceParsing_demo(model: FaceParser, img, convert_to_seg12=True, model_name="default"):
    """face_parsing_demo.py:294-318 (the "default" BiSeNet parser): uint8 numpy label map."""
    if model_name != "default":
        raise NotImplementedError("only the BiSeNet ('default') parser is implemented")
    seg19, seg12 = model.engine.face_parse(model.preprocess_img(img))
    out = (seg12 if convert_to_seg12 else seg19)[0]
    return out.cpu().numpy().astype(np.uint8)


def prepare_inpaint(model: FaceParser, img_m11, remove=REMOVE_MASK_TAR_FFHQ):
    """Target-side preparation of ldm/data/video_swap_dataset.py:135-222 for images already in [-1,1]:
    parse -> 12-class map -> mask = 1 - isin(map, remove) -> inpaint = image * mask.  Returns (mask, inpaint, seg12)."""
    img_m11 = img_m11.to(model.engine.device, torch.float32)
    _, seg12 = model.engine.face_parse((img_m11 + 1.0) * 0.5)
    mask, inp = model.engine.inpaint_from_parsing(img_m11, seg12, remove)
    return mask, inp, seg12
