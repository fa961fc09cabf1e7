"""Where does the 2-CTA GEMM spend its time?  Per-CTA clock64 counters (option gemm_debug) for a few shapes/configs."""
import ctypes as C
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from reface_b200.runtime import Engine

eng = Engine(0, arena_bytes=12 << 30)


def run(M, K, N, **opts):
    base = dict(gemm_pair=1, gemm_kmerge=2, gemm_bn=0, gemm_stages=0, gemm_persistent=1, gemm_debug=1)
    base.update(opts)
    for k, v in base.items():
        eng.set_option(k, v)
    x = torch.randn(M, K, device="cuda").half().float()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half().float()
    eng.op_linear(x, w)
    eng.set_option("profile", 1)
    eng.op_linear(x, w)
    ms, fl, n = eng.profile_read()
    eng.set_option("profile", 0)
    buf = (C.c_ulonglong * (148 * 8))()
    eng._ck(eng.lib.rfb_debug_read(eng.h, buf, 148 * 8))
    d = np.array(buf[:], dtype=np.float64).reshape(148, 8)
    lead = d[0::2]
    peer = d[1::2]
    tot = lead[:, 0].mean()
    print(f"M={M} K={K} N={N} {opts}: {2.0*M*N*K/ms/1e9:7.1f} TF/s  {ms*1e3:7.1f} us | MMA thread total {tot:9.0f} cyc: wait_full {100*lead[:,1].mean()/tot:5.1f}% "
          f"wait_acc {100*lead[:,2].mean()/tot:5.1f}% | k-stages {lead[:,6].mean():.0f} tiles {lead[:,7].mean():.1f} -> {tot/max(lead[:,6].mean(),1):6.0f} cyc/stage | "
          f"producer wait_empty lead {100*lead[:,3].mean()/tot:5.1f}% peer {100*peer[:,3].mean()/tot:5.1f}% | epi wait {100*lead[:,4].mean()/max(lead[:,5].mean(),1):5.1f}% of {lead[:,5].mean():9.0f}",
          flush=True)


for shp in [(65536, 2880, 320), (16384, 5760, 640), (4096, 11520, 1280)]:
    for o in [dict(), dict(gemm_kmerge=1), dict(gemm_bn=256), dict(gemm_bn=256, gemm_kmerge=1), dict(gemm_bn=64), dict(gemm_stages=2)]:
        run(*shp, **o)
print("DEBUG DONE")
