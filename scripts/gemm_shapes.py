"""Per-launch CUDA-event times of every tensor-core launch of one UNet forward (N=16, L=64), aggregated per shape."""
import collections, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("CHILD") != "1":
    r = subprocess.run([sys.executable, __file__], env=dict(os.environ, CHILD="1"), capture_output=True, text=True)
    out = r.stdout
    if r.returncode != 0:
        print(r.stderr[-2000:])
    agg = collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"PROF kind=(\d+) M=(\d+) N=(\d+) K=(\d+) BN=(\d+) z=(\d+) mode=(\d+) us=([\d.]+) tflops=([\d.]+)", line)
        if not m:
            if not line.startswith("PROF"):
                print(line)
            continue
        key = tuple(int(v) for v in m.groups()[:7])
        a = agg.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1; a[1] += float(m.group(8)); a[2] += float(m.group(9)) * float(m.group(8))
    tot = sum(a[1] for a in agg.values())
    print(f"{'kind':>4} {'M':>6} {'N':>5} {'K':>6} {'BN':>4} {'z':>4} {'mode':>4} {'n':>3} {'us/launch':>10} {'TFLOP/s':>8} {'total us':>9} {'share':>6}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{key[0]:4d} {key[1]:6d} {key[2]:5d} {key[3]:6d} {key[4]:4d} {key[5]:4d} {key[6]:4d} {a[0]:3d} {a[1]/a[0]:10.1f} {a[2]/a[1]:8.1f} {a[1]:9.1f} {100*a[1]/tot:5.1f}%")
    print(f"total {tot/1e3:.3f} ms")
    sys.exit(0)
import torch
from reface_b200 import synth
from reface_b200.runtime import Engine
N = int(os.environ.get("N", 16)); L = int(os.environ.get("L", 64))
dev = torch.device("cuda", 0)
flat = synth.random_flat(dev, 0)
sd = {k: v for k, v in synth.state_dict_from_flat(flat).items() if k.startswith("model.diffusion_model.")}
eng = Engine(0)
for k, v in os.environ.items():
    if k.startswith("RFB_") and k != "RFB_CPU_THREADS":
        eng.set_option(k[4:].lower(), int(v))
eng.load_state_dict(sd)
eng.build_unet()
x = torch.randn(N, 9, L, L, device=dev); t = torch.full((N,), 981, device=dev, dtype=torch.long)
ctx = torch.randn(N, 1, 768, device=dev)
for _ in range(2):
    eng.unet_forward(x, t, ctx)
torch.cuda.synchronize()
eng.set_option("profile", 2)
eng.unet_forward(x, t, ctx)
sys.stdout.flush()
eng.profile_read()
eng.set_option("profile", 0)
