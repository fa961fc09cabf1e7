"""Per-stage device times of one B=8 swap at 512x512 (conditioning / VAE encode / DDIM loop / VAE decode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reface_b200 import synth
from reface_b200.ldm_api import LatentDiffusion, DDIMSampler
from reface_b200.runtime import Engine

dev = torch.device("cuda", 0)
B, H, S = int(os.environ.get("B", 8)), int(os.environ.get("H", 512)), int(os.environ.get("S", 50))
eng = Engine(0)
for k, v in os.environ.items():
    if k.startswith("RFB_") and k != "RFB_CPU_THREADS":
        eng.set_option(k[4:].lower(), int(v))
model = LatentDiffusion(synth.state_dict_from_flat(synth.random_flat(dev, 0)), engine=eng)
inp = synth.synthetic_inputs(B, H, dev)


def timed(fn, reps=2):
    fn(); fn(); torch.cuda.synchronize()      # two warm calls: the second one captures the sampling loop's CUDA graph
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

uc = model.learnable_vector.repeat(B, 1, 1)
t_cond, c = timed(lambda: model.conditioning_with_feat(inp["ref_img"], tar=inp["tar_img"], landmarks136=inp["landmarks136"]))
t_clip, _ = timed(lambda: eng.clip_encode(inp["ref_img"]))
t_arc, _ = timed(lambda: eng.arcface_embed(inp["ref_img"]))
t_enc, z = timed(lambda: model.get_first_stage_encoding(model.encode_first_stage(inp["inpaint_img"]), noise=inp["enc_noise"]))
smp = DDIMSampler(model)
t_loop, (x0, _) = timed(lambda: smp.sample(S=S, conditioning=c, batch_size=B, shape=[4, H // 8, H // 8], verbose=False,
                                          unconditional_guidance_scale=3.5, unconditional_conditioning=uc, eta=0.0,
                                          x_T=inp["x_T"], test_model_kwargs={"inpaint_image": z, "inpaint_mask": inp["mask_lat"]}), reps=1)
t_dec, img = timed(lambda: model.decode_first_stage(x0))
tot = t_cond + t_enc + t_loop + t_dec
print(f"B={B} H={H} S={S}: cond {t_cond:.1f} ms (clip x1 {t_clip:.1f}, arcface {t_arc:.1f}) | vae enc {t_enc:.1f} | ddim loop {t_loop:.1f} "
      f"({t_loop/S:.2f}/step) | vae dec {t_dec:.1f} | total {tot:.1f} ms -> {B/tot*1e3:.2f} faces/s; finite={bool(torch.isfinite(img).all())}")
