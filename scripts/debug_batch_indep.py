import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reface_b200 import synth
from reface_b200.ldm_api import LatentDiffusion, swap_faces
from reface_b200.runtime import Engine
from reface_b200.shard import shard_batch
dev = torch.device("cuda", 0)
eng = Engine(0)
for k, v in os.environ.items():
    if k.startswith("RFB_") and k != "RFB_CPU_THREADS":
        eng.set_option(k[4:].lower(), int(v))
model = LatentDiffusion(synth.state_dict_from_flat(synth.random_flat(dev, 0)), engine=eng)
inp = synth.synthetic_inputs(8, 512, dev, seed=42)
S = int(os.environ.get("S", 1))
full = swap_faces(model, S=S, scale=3.5, **inp)
parts = [swap_faces(model, S=S, scale=3.5, **shard_batch(inp, r, 2)) for r in range(2)]
for k in ("c", "z_inpaint", "samples", "image"):
    a = full[k]; b = torch.cat([p[k] for p in parts], 0)
    print(k, "equal" if torch.equal(a, b) else f"DIFF max {float((a-b).abs().max()):.3e} of {float(a.abs().max()):.3e}")
# UNet alone
x = torch.randn(16, 9, 64, 64, device=dev); t = torch.full((16,), 981, device=dev, dtype=torch.long); ctx = torch.randn(16, 1, 768, device=dev)
e16 = eng.unet_forward(x, t, ctx); e8 = torch.cat([eng.unet_forward(x[:8], t[:8], ctx[:8]), eng.unet_forward(x[8:], t[8:], ctx[8:])])
print("unet N=16 vs 2x8:", "equal" if torch.equal(e16, e8) else f"DIFF {float((e16-e8).abs().max()):.3e}")
z = torch.randn(8, 4, 64, 64, device=dev)
d8 = eng.vae_decode(z); d4 = torch.cat([eng.vae_decode(z[:4]), eng.vae_decode(z[4:])])
print("vae decode 8 vs 2x4:", "equal" if torch.equal(d8, d4) else f"DIFF {float((d8-d4).abs().max()):.3e}")
img = torch.rand(8, 3, 512, 512, device=dev) * 2 - 1
n = torch.randn(8, 4, 64, 64, device=dev)
q8 = eng.vae_encode(img, n); q4 = torch.cat([eng.vae_encode(img[:4], n[:4]), eng.vae_encode(img[4:], n[4:])])
print("vae encode 8 vs 2x4:", "equal" if torch.equal(q8, q4) else f"DIFF {float((q8-q4).abs().max()):.3e}")
