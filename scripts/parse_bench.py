"""Face parser (BiSeNet, 26.8 GFLOP per 512x512 image) throughput on one B200: images/s with device-resident
inputs (CUDA events), launches per call, and the CPU oracle on the host cores next to it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch
import reface_oracle as O
from reface_b200.runtime import Engine
torch.set_grad_enabled(False)
sd = O.init_state_dict(O.parse_spec(), 0)
eng = Engine(0, arena_bytes=16 << 30)
eng.load_state_dict(sd)
eng.build_face_parser(O.PFX_PARSE)
for B in (1, 8, 32):
    img = torch.rand(B, 3, 512, 512, device="cuda")
    eng.face_parse(img); torch.cuda.synchronize()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        s19, s12 = eng.face_parse(img)
        m, inp = eng.inpaint_from_parsing(img * 2 - 1, s12)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"face_parse + inpaint prep B={B} 512x512: {ms:7.3f} ms/call  {B/ms*1e3:8.1f} images/s  "
          f"{26.77e9*B/ms/1e9:7.1f} TFLOP/s (conv FLOPs)  launches/call {(eng.launch_count-l0)//10}", flush=True)
torch.set_num_threads(min(32, os.cpu_count() or 1))
P = O.Params(sd, O.PFX_PARSE)
img = torch.rand(1, 3, 512, 512)
O.face_parse(P, img)
t0 = time.time(); O.face_parse(P, img); dt = time.time() - t0
print(f"CPU oracle (torch fp32, {torch.get_num_threads()} threads): {dt*1e3:.1f} ms/image  {1/dt:.2f} images/s")
