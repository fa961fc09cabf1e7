"""CPU: the oracle (oracle/reface_oracle.py) reproduces the golden vectors generated from the REAL
reference modules by tests/golden/make_golden.py, plus the closed-form constants of SURVEY 8(c)."""
import os

import numpy as np
import torch

from conftest import GOLDEN

torch.set_grad_enabled(False)


def _g(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, name + ".npz")).items()}


def test_closed_form_constants(oracle):
    ac = oracle.alphas_cumprod_f32()
    assert abs(float(ac[0]) - 0.99915) < 1e-6 and abs(float(ac[999]) - 0.0046600985) < 1e-9
    assert list(oracle.make_ddim_timesteps(50)[:3]) == [1, 21, 41] and oracle.make_ddim_timesteps(50)[-1] == 981
    assert len(oracle.make_ddim_timesteps(30)) == 31 and list(oracle.make_ddim_timesteps(5)) == [1, 201, 401, 601, 801]
    te = oracle.timestep_embedding(torch.tensor([981]), 320)
    assert abs(float(te[0, 0]) - 0.67995721) < 1e-6 and abs(float(te[0, 160]) - 0.73325181) < 1e-6
    g = _g("schedule")
    assert torch.equal(g["alphas_cumprod"], ac)
    assert torch.equal(g["temb"], oracle.timestep_embedding(g["temb_t"], 320))


def test_unet_matches_reference_golden(oracle, unet_sd):
    g = _g("unet_L16")
    taps = {}
    eps = oracle.unet_forward(oracle.Params(unet_sd, oracle.PFX_UNET), g["x"], g["t"], g["ctx"], taps=taps)
    assert (eps - g["eps"]).abs().max() < 1e-4
    assert abs(float(taps["middle_block"].std()) - float(g["tapstd_middle_block"])) < 1e-4
    assert float(eps.abs().max()) > 0.1      # zero_module'd layers are re-randomised: the net is not vacuous


def test_ddim_matches_reference_sampler(oracle, unet_sd):
    g = _g("ddim_S5_L16")
    x0, inter = oracle.ddim_sample(oracle.Params(unet_sd, oracle.PFX_UNET), g["x_T"], g["z"], g["mask"], g["c"], g["uc"],
                                   5, 3.5, log_every_t=2)
    assert (x0 - g["x0"]).abs().max() < 1e-3 * float(g["x0"].abs().max())
    assert len(inter["x_inter"]) == int(g["n_inter"])
    assert (inter["pred_x0"][-1] - g["pred_x0_last"]).abs().max() < 1e-3 * float(g["x0"].abs().max())


def test_plms_and_q_sample_match_reference_golden(oracle, unet_sd):
    """PLMSSampler (plms.py) and DDPM.q_sample (ddpm.py:412-415): fixtures written by tests/golden/make_golden.py from the
    reference's own classes."""
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, "plms_S6_L16.npz")).items()}
    assert torch.equal(oracle.q_sample(g["z"], g["q_t"], g["q_noise"]), g["q_out"])
    x0, inter = oracle.plms_sample(oracle.Params(unet_sd, oracle.PFX_UNET), g["x_T"], g["z"], g["mask"], g["c"], g["uc"], 6,
                                   3.5, log_every_t=2)
    assert len(inter["x_inter"]) == int(g["n_inter"])
    assert float((x0 - g["x0"]).abs().max()) <= 5e-5 * float(g["x0"].abs().max())


def test_face_parser_matches_reference_golden(oracle):
    """BiSeNet + label conversion (pretrained/face_parsing): fixture written from the reference's own modules."""
    g = _g("parse_256")
    sd = oracle.init_state_dict(oracle.parse_spec(), 0)
    img01 = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(int(g["img_seed"])))
    seg19, seg12 = oracle.face_parse(oracle.Params(sd, oracle.PFX_PARSE), img01)
    assert torch.equal(seg19.to(torch.uint8), g["seg19"]) and torch.equal(seg12.to(torch.uint8), g["seg12"])
    mean = torch.tensor(oracle.SEG_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(oracle.SEG_STD).view(1, 3, 1, 1)
    lg = oracle.bisenet_logits(oracle.Params(sd, oracle.PFX_PARSE), (img01 - mean) / std)
    assert float((lg[:, :, ::16, ::16] - g["logits_sub"]).abs().max()) < 1e-4
    m, inp = oracle.inpaint_from_parsing(img01 * 2 - 1, seg12)
    assert set(m.unique().tolist()) <= {0.0, 1.0} and torch.equal(inp, (img01 * 2 - 1) * m)


def test_paste_back_matches_pillow_golden(oracle):
    """Paste-back (inference_swap_video.py:702-721): the restatement of Pillow's resize / PERSPECTIVE transform / alpha
    composite reproduces the fixture written by Pillow itself bit for bit."""
    g = np.load(os.path.join(GOLDEN, "paste_64.npz"))
    out = oracle.paste_back(g["x01"], g["orig"], g["coeffs"], up=int(g["up"]))
    assert np.array_equal(out, g["pasted"])
    assert 0.2 < float((out != g["orig"]).any(-1).mean()) < 1.0       # part of the frame is replaced, part shows through


def test_vae_matches_reference_golden(oracle, vae_sd):
    g = _g("vae_64")
    P = oracle.Params(vae_sd, oracle.PFX_VAE)
    mean, logvar = oracle.vae_encode_moments(P, g["x"])
    assert (mean - g["mean"]).abs().max() < 1e-4 and (logvar - g["logvar"]).abs().max() < 1e-4
    assert (oracle.vae_encode(P, g["x"], g["noise"]) - g["z"]).abs().max() < 1e-4
    assert (oracle.vae_decode(P, g["zdec"]) - g["img"]).abs().max() < 1e-4


def test_clip_and_conditioning_match_reference_golden(oracle, clip_sd, arc_sd, fusion_sd):
    g = _g("clip_B1")
    img = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(int(g["img_seed"])))
    out = oracle.clip_embed(oracle.Params(clip_sd, oracle.PFX_CLIP), img)
    assert (out - g["out"]).abs().max() < 1e-4
    g = _g("cond_B2")
    ref_img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(int(g["ref_seed"])))
    idf = oracle.arcface_embed(oracle.Params(arc_sd, oracle.PFX_ARC), ref_img)
    assert (idf - g["id_feat"]).abs().max() < 1e-5
    full = dict(clip_sd); full.update(arc_sd); full.update(fusion_sd)
    c = oracle.conditioning_with_feat(oracle.Params(full), ref_img, g["tar"], torch.zeros(2, 136))
    assert (c - g["c"]).abs().max() < 1e-4


def test_concat_and_update_are_exact(oracle):
    g = torch.Generator().manual_seed(0)
    x, z, m = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g), torch.rand(2, 1, 8, 8, generator=g)
    assert torch.equal(oracle.concat9(x, z, m), torch.cat([x, z, m], 1))


def test_video_settings_and_landmark_conditioning_match_reference_golden(oracle, unet_sd, clip_sd, arc_sd, fusion_sd):
    """Round-2 fixtures from the real reference: the 31-step / scale-3 sampler run of the video settings
    (first 3 steps re-run here; the generating script compared all 31) and conditioning with detected landmarks."""
    g = _g("ddim_S30_L16")
    assert len(oracle.make_ddim_timesteps(30)) == 31 and np.array_equal(g["timesteps"].numpy(), oracle.make_ddim_timesteps(30))
    _, inter = oracle.ddim_sample(oracle.Params(unet_sd, oracle.PFX_UNET), g["x_T"], g["z"], g["mask"], g["c"], g["uc"],
                                  30, 3.0, log_every_t=1, n_steps_limit=3)
    for i in range(3):
        ref = g["x_inter"][i]
        assert float((inter["x_inter"][1 + i] - ref).abs().max()) <= 5e-5 * float(ref.abs().max())
    g = _g("cond_B2_lm")
    gen = torch.Generator().manual_seed(int(g["seed"]))
    ref_img = torch.randn(2, 3, 224, 224, generator=gen)
    tar = torch.rand(2, 3, 64, 64, generator=gen) * 2 - 1
    sd = {**clip_sd, **arc_sd, **fusion_sd}
    c = oracle.conditioning_with_feat(oracle.Params(sd), ref_img, tar, g["lm_raw"])
    assert float((c - g["c"]).abs().max()) <= 5e-5 * float(g["c"].abs().max())
