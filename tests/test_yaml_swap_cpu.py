"""CPU (build container only: needs /root/reference): the REAL reference LatentDiffusion is instantiated from the
reference's own YAML (models/REFace/configs/project_ffhq.yaml) with the three plugin slots pointed at the drop-in
classes, and the reference's `model.load_state_dict(sd, strict=False)` (scripts/inference_test_bench.py:98-103) must
reach every shell through nn.Module._load_from_state_dict and trigger its build.  The engine is replaced by a recorder:
no GPU is needed to check the plumbing ("YAML-only swap")."""
import os
import sys
import tempfile

import pytest
import torch

from conftest import ROOT

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")


class _Recorder:
    device = torch.device("cpu")

    def __init__(self):
        self.keys, self.built = set(), []

    def load_state_dict(self, sd):
        self.keys |= set(sd)

    def build_unet(self, p):
        self.built.append(("unet", p))

    def build_vae(self, p):
        self.built.append(("vae", p))

    def build_clip(self, p):
        self.built.append(("clip", p))


def test_reference_latent_diffusion_with_swapped_targets(oracle):
    for p in (os.path.join(ROOT, "oracle", "ref_shims"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.dont_write_bytecode = True
    torch.set_grad_enabled(False)
    from omegaconf import OmegaConf
    from ldm.util import instantiate_from_config
    import ldm.modules.encoders.modules as em
    from src.Face_models.encoders.model_irse import Backbone
    import reface_b200.ldm_api as api

    cfg = OmegaConf.load(os.path.join(REF, "models/REFace/configs/project_ffhq.yaml")).model
    cfg.params.unet_config.target = "reface_b200.ldm_api.UNetModel"
    cfg.params.first_stage_config.target = "reface_b200.ldm_api.AutoencoderKL"
    # the cond-stage slot is swapped by aliasing the class, not the string: ddpm.py:725-736 compares
    # cond_stage_config.target with "ldm.modules.encoders.modules.FrozenCLIPEmbedder" to create proj_out_source/target
    ref_cls = em.FrozenCLIPEmbedder
    em.FrozenCLIPEmbedder = api.FrozenCLIPEmbedder
    try:
        with tempfile.TemporaryDirectory() as td:
            arc = os.path.join(td, "arc.pth")
            torch.save(Backbone(input_size=112, num_layers=50, drop_ratio=0.6, mode="ir_se").state_dict(), arc)
            cfg.params.cond_stage_config.other_params.arcface_path = arc
            cfg.params.cond_stage_config.other_params.Additional_config.LPIPS_loss_weight = 0   # LPIPS downloads weights
            model = instantiate_from_config(cfg).eval()
    finally:
        em.FrozenCLIPEmbedder = ref_cls
    shells = (model.model.diffusion_model, model.first_stage_model, model.cond_stage_model)
    assert [type(s) for s in shells] == [api.UNetModel, api.AutoencoderKL, api.FrozenCLIPEmbedder]
    assert hasattr(model, "proj_out_source") and hasattr(model, "proj_out_target")
    rec = _Recorder()
    for s in shells:
        s._engine = rec
    sd = oracle.init_state_dict(oracle.full_spec(), 0)
    res = model.load_state_dict(sd, strict=False)
    assert rec.built == [("unet", "model.diffusion_model."), ("vae", "first_stage_model."), ("clip", "cond_stage_model.")]
    for pfx in ("model.diffusion_model.", "first_stage_model.", "cond_stage_model."):
        want = {k for k in sd if k.startswith(pfx)}
        assert want and want <= rec.keys, pfx
    assert not res.unexpected_keys, res.unexpected_keys[:5]
    # the schedule buffers are rebuilt by register_schedule, never loaded from our synthetic state dict
    assert all(not k.startswith(("model.", "first_stage_model.", "cond_stage_model.")) for k in res.missing_keys)
    # nn.Module conveniences the scripts call on the model work with the shells inside (inference_test_bench.py:111,334)
    model.to(torch.device("cpu")).eval()
    # AutoencoderKL.encode must hand back something get_first_stage_encoding accepts (ddpm.py:850-857)
    from ldm.modules.distributions.distributions import DiagonalGaussianDistribution
    post = api._posterior_cls()(rec, torch.zeros(1, 3, 8, 8))
    assert isinstance(post, DiagonalGaussianDistribution) and isinstance(post, api._Posterior)
