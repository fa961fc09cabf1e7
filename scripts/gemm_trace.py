"""clock64 role breakdown of gemm_persist_kernel (option gemm_debug) at the UNet's characteristic shapes: how long the MMA
warp waits for operands (full) / for a free accumulator stage (acc), the TMA producer for a free ring slot, and an epilogue
warp for the accumulators vs. how long it spends in prefetch / drain.  Averages over the CTAs of one launch."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reface_b200.runtime import Engine
eng = Engine(0, arena_bytes=12 << 30)
for k, v in os.environ.items():
    if k.startswith("RFB_"):
        eng.set_option(k[4:].lower(), int(v))


def read():
    buf = (C.c_ulonglong * (148 * 8))()
    eng._ck(eng.lib.rfb_debug_read(eng.h, buf, 148 * 8))
    v = list(buf)
    rows = [v[i * 8:(i + 1) * 8] for i in range(148) if v[i * 8]]
    if int(os.environ.get("RFB_GEMM_PAIR", 0)):   # gemm_pair_kernel: the MMA counters live in the leader (even) CTA of a pair
        rows = [v[i * 8:(i + 1) * 8] for i in range(0, 148, 2) if v[i * 8 + 7]]
    avg = lambda j: sum(r[j] for r in rows) / max(1, len(rows))
    return dict(mma_total=avg(0), wait_full=avg(1), wait_acc=avg(2), prod_wait_empty=avg(3), epi_wait_acc=avg(4), epi_total=avg(5),
                epi_prefetch=avg(6), tiles=avg(7), ctas=len(rows))


ISSUE = int(os.environ.get("ISSUE", 0))   # multicast / pair kernels: slot 6 = cycles the producer spends issuing its TMA loads


def report(name, fn):
    fn(); torch.cuda.synchronize()
    eng.set_option("gemm_debug", 1)
    fn(); torch.cuda.synchronize()
    d = read()
    eng.set_option("gemm_debug", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    t = max(d["tiles"], 1)
    print(f"{name:44s} {us:7.1f} us/call(incl. casts) | per tile: MMA warp {d['mma_total']/t:7.0f} cyc (wait operands {100*d['wait_full']/d['mma_total']:4.1f} %, "
          f"wait acc stage {100*d['wait_acc']/d['mma_total']:4.1f} %) | producer waits for a slot {100*d['prod_wait_empty']/d['mma_total']:4.1f} % | "
          + (f"producer TMA issue {100*d['epi_prefetch']/d['mma_total']:4.1f} % | " if ISSUE else "") +
          f"epilogue warp {d['epi_total']/t:7.0f} cyc (wait acc {100*d['epi_wait_acc']/d['epi_total']:4.1f} %, prefetch {100*d['epi_prefetch']/d['epi_total']:4.1f} %, "
          f"drain {100*(d['epi_total']-d['epi_wait_acc']-d['epi_prefetch'])/d['epi_total']:4.1f} %) tiles/CTA {t:.1f}", flush=True)


g = lambda *s: torch.randn(*s, device="cuda").half().float()
LONGK = int(os.environ.get("LONGK", 0))       # only the long-K shapes
for (M, K, N, res, geglu) in [(65536, 1280, 320, True, False), (16384, 2560, 640, True, False), (4096, 5120, 1280, True, False)] if LONGK else [(65536, 320, 320, False, False), (65536, 320, 320, True, False), (65536, 320, 960, False, False),
                             (65536, 320, 2560, False, True), (65536, 1280, 320, True, False), (16384, 640, 640, True, False),
                             (4096, 1280, 1280, True, False)]:
    x, w, b = g(M, K), g(N, K) / math.sqrt(K), torch.randn(N, device="cuda")
    r = g(M, N) if res else None
    report(f"linear M={M} K={K} N={N} res={int(res)} geglu={int(geglu)}", lambda: eng.op_linear(x, w, b, residual=r, geglu=geglu))
for (n, c, h, o) in [(16, 320, 64, 320), (16, 640, 32, 640), (16, 1280, 16, 1280)]:
    x, w, b = g(n, c, h, h), g(o, c, 3, 3) / math.sqrt(9 * c), torch.randn(o, device="cuda")
    report(f"conv3x3 N={n} C={c} H={h} O={o}", lambda: eng.op_conv2d(x, w, b))
