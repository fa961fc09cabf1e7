#!/usr/bin/env python
"""bench.py -- faces/sec of the REFace DDIM face-swap path (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 3 --warmup 3            # our arm (hand-written sm_100a kernels), configs[1]
    python bench.py --impl reference --steps 2 --warmup 1    # the reference algorithm's CPU port (oracle)
    python bench.py --workload 1024 --steps 5                # configs[3] shape: 1024x1024, B=2 per GPU (16 on 8 GPUs)
    python bench.py --workload video --steps 3               # configs[4]: 30 frames per GPU per step (240 on 8 GPUs)
                                                             # through swap_video: 31 timesteps, scale 3, frames/s

One "step" = one batch of B faces through the whole path: conditioning (CLIP x2 + ArcFace + fusion) ->
VAE encode -> S-step CFG DDIM over the 9-channel UNet -> VAE decode (config[1]: 512x512, 50 steps, CFG 3.5, B=8
per GPU).  Synthetic inputs / random-init weights of the reference architecture (no checkpoints offline).
Multi-GPU: one process per GPU (torchrun), weights broadcast once over NCCL, batch sharded, no collective in the
step loop ("scaling": "weak").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_FACE = {(512, 50): 83.65e12, (256, 5): 2.98e12, (1024, 50): 481.4e12}   # SURVEY 8(d), reference-equivalent
FLOP_PER_VIDEO_FRAME = 53.2e12                                                    # 512^2, 31 timesteps, one CLIP pass
# BASELINE.json configs -> (size, ddim steps, scale, faces or frames per GPU per step, arena GB)
WORKLOADS = {"512": dict(size=512, ddim_steps=50, scale=3.5, batch=8, arena_gb=40, config="configs[1] (configs[2] at 8 GPUs)"),
             "1024": dict(size=1024, ddim_steps=50, scale=3.5, batch=2, arena_gb=48, config="configs[3] at 8 GPUs"),
             "video": dict(size=512, ddim_steps=30, scale=3.0, batch=30, arena_gb=100, config="configs[4] at 8 GPUs")}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", 1386.9), burst=d.get("bf16_tflops", 1654.2),
                    hbm=d.get("hbm_gbs", 6566.7), source="measured (MEASURED_PEAKS.json)")
    return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def traffic_from_profiles():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (the long-K implicit-GEMM conv
    640->640 at 32x32, N=16) from the NEWEST committed `ncu --set full` capture: profiles/r<round>[s<session>]_traffic.json,
    written by scripts/ncu_summary.py --traffic.  Returns (bytes_per_launch, file name)."""
    import glob
    import re

    def key(path):
        m = re.match(r"r(\d+)(?:s(\d+))?_traffic\.json", os.path.basename(path))
        return (int(m.group(1)), int(m.group(2) or 1)) if m else (-1, -1)

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), key=key)
    if not files:
        return None, None
    return json.load(open(files[-1])).get("dram_bytes_per_launch"), "profiles/" + os.path.basename(files[-1])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].startswith("Active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.rows))


def n_timesteps(S):
    return len(range(0, 1000, 1000 // S))           # make_ddim_timesteps (util.py:46-60): S=30 gives 31


def cpu_baseline(args, steps, warmup, full_batch_once=False):
    """The reference algorithm (oracle port: bit-identical to the reference's modules, tests/golden/make_golden.py;
    fp32, the host's cores): a bounded sample of the same workload -- UNet CFG calls for ONE face (N=2) at the benchmark
    resolution + one VAE encode/decode + one conditioning pass; units/s = 1 / (T * t_unet_call + t_enc + t_dec + t_cond),
    T = number of DDIM timesteps.  full_batch_once: also time ONE UNet CFG call at the arm's per-GPU batch (capped at 8)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reface_oracle as O
    torch.set_grad_enabled(False)
    ncpu = os.cpu_count() or 1
    cores = min(ncpu, int(os.environ.get("RFB_CPU_THREADS", 32)))   # torch CPU ops stop scaling (and regress) beyond ~32 threads
    torch.set_num_threads(cores)
    H, T = args.size, n_timesteps(args.ddim_steps)
    L = H // 8
    video = args.workload == "video"
    sd = O.init_state_dict(O.full_spec(), 0)
    P = O.Params(sd)
    inp = O.synthetic_inputs(1, H)
    t0 = time.time()
    if video:    # the source CLIP / ArcFace features are computed once per video: per frame only the target CLIP pass
        O.clip_embed(P.sub(O.PFX_CLIP), O.target_clip_input(inp["tar_img"]))
        c = torch.randn(1, 1, 768)
    else:
        c = O.conditioning_with_feat(P, inp["ref_img"], inp["tar_img"], inp["landmarks136"])
    t_cond = time.time() - t0
    t0 = time.time()
    z = O.vae_encode(P.sub(O.PFX_VAE), inp["inpaint_img"], inp["enc_noise"])
    t_enc = time.time() - t0
    uc = sd["learnable_vector"]
    x9 = torch.cat([O.concat9(inp["x_T"], z, inp["mask_lat"])] * 2)
    tt = torch.tensor([981, 981])
    cc = torch.cat([uc, c])
    ts = []
    for i in range(warmup + steps):
        t0 = time.time()
        O.unet_forward(P.sub(O.PFX_UNET), x9, tt, cc)
        if i >= warmup:
            ts.append(time.time() - t0)
    t_unet = sum(ts) / len(ts)
    t0 = time.time()
    O.vae_decode(P.sub(O.PFX_VAE), inp["x_T"] * 0.18215)
    t_dec = time.time() - t0
    fps = 1.0 / (T * t_unet + t_enc + t_dec + t_cond)
    unit = "frames/s" if video else "faces/s"
    note = ""
    if full_batch_once and L <= 64:
        b = min(args.batch, 8)
        xb, tb = x9[:1].repeat(2 * b, 1, 1, 1), tt[:1].repeat(2 * b)
        cb = torch.cat([uc.repeat(b, 1, 1), c.repeat(b, 1, 1)])
        t0 = time.time()
        O.unet_forward(P.sub(O.PFX_UNET), xb, tb, cb)
        t_b = time.time() - t0
        note = f"; one UNet CFG call at the arm's batch (B={b}, N={2 * b}): {t_b:.2f}s = {t_b / b:.2f}s per face"
    return dict(value=fps, unit=unit, cores=cores, kind="port",
                sample=f"{steps} timed UNet CFG calls (1 face, N=2, L={L}) + 1 VAE encode + 1 decode + 1 conditioning pass"
                       f"{' (target CLIP only: source features are per video)' if video else ''}; "
                       f"t_unet={t_unet:.2f}s t_enc={t_enc:.2f}s t_dec={t_dec:.2f}s t_cond={t_cond:.2f}s; "
                       f"{unit} = 1/({T}*t_unet+t_enc+t_dec+t_cond){note}; oracle port of the reference (the reference "
                       f"tree does not travel to the GPU box), extrapolated from one face"), t_unet * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="512", choices=sorted(WORKLOADS),
                    help="BASELINE.json config: 512 = configs[1]/[2] (default), 1024 = configs[3], video = configs[4]")
    ap.add_argument("--batch", type=int, default=None, help="faces (video: frames) per GPU per step")
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--ddim-steps", type=int, default=None)
    ap.add_argument("--scale", type=float, default=None)
    ap.add_argument("--chunk", type=int, default=30, help="video: frames per streamed chunk (inference_swap_video.py batches)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    for k in ("batch", "size", "ddim_steps", "scale"):
        if getattr(args, k) is None:
            setattr(args, k, wl[k])
    video = args.workload == "video"
    unit = "frames/s" if video else "faces/s"

    # stdout carries exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
    # version banner there on the first collective), so fd 1 is pointed at stderr for the whole run and the JSON line is
    # written to the saved original descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    T = n_timesteps(args.ddim_steps)
    is_named = all(getattr(args, k) == wl[k] for k in ("batch", "size", "ddim_steps", "scale"))
    what = (f"video swap: {args.batch} frames/GPU/step streamed in chunks of {args.chunk} through swap_video (source CLIP/ArcFace "
            f"once per video, target CLIP per frame), {args.size}x{args.size}, --ddim_steps {args.ddim_steps} = {T} timesteps, CFG {args.scale}"
            if video else f"{args.size}x{args.size}, {args.ddim_steps} DDIM steps, CFG {args.scale}, batch={args.batch}/GPU")
    config = dict(workload=f"{what} ({'BASELINE ' + wl['config'] if is_named else 'custom shape'}); full path: "
                           f"{'target CLIP' if video else 'CLIPx2+ArcFace'}+fusion, VAE encode, DDIM, VAE decode",
                  global_batch=args.batch * world, parallelism=f"dp{world}",
                  l2_policy="per-step working set (activations + 2.7 GB fp16 weights) exceeds the 126 MB L2")
    if video:
        metric = f"frames/sec @{args.size}x{args.size} video, {args.ddim_steps} DDIM steps ({T} timesteps), CFG {args.scale:g}"
    elif (args.size, args.ddim_steps, args.scale) == (512, 50, 3.5):
        metric = "faces/sec @512x512, 50 DDIM steps, CFG 3.5"
    else:
        metric = f"faces/sec @{args.size}x{args.size}, {args.ddim_steps} DDIM steps, CFG {args.scale}"

    if args.impl == "reference":
        if rank != 0:
            return
        cb, step_ms = cpu_baseline(args, max(1, args.steps), max(0, args.warmup), full_batch_once=True)
        line = dict(metric=metric, value=cb["value"], unit=unit, n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=step_ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32", data="synthetic", config=config, impl="reference", cpu_baseline=cb,
                    e2e=dict(value=cb["value"], unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        emit(line)
        return

    import torch.distributed as dist
    from reface_b200 import synth
    from reface_b200.ldm_api import LatentDiffusion, swap_faces, swap_video
    from reface_b200.runtime import Engine
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # weights: generated once on rank 0, one NCCL broadcast of the flat checkpoint over NVLink
    if rank == 0:
        flat = synth.random_flat(dev, seed=0)
    else:
        flat = torch.empty(synth.flat_layout()[1], dtype=torch.float32, device=dev)
    from reface_b200 import shard
    shard.broadcast_checkpoint(flat, 0)
    eng = Engine(local, arena_bytes=int(os.environ.get("RFB_ARENA_GB", wl["arena_gb"])) << 30)
    for k, v in os.environ.items():          # e.g. RFB_GEMM_PAIR=1 to A/B an engine option from the command line
        if k.startswith("RFB_") and k not in ("RFB_CPU_THREADS", "RFB_ARENA_GB"):
            eng.set_option(k[4:].lower(), int(v))
    model = LatentDiffusion(synth.state_dict_from_flat(flat), engine=eng)
    del flat
    torch.cuda.empty_cache()

    B, H, S = args.batch, args.size, args.ddim_steps
    dev_in = synth.synthetic_inputs(B, H, dev, seed=42 + rank)
    host_in = synth.synthetic_inputs(B, H, dev, seed=42 + rank, pinned_host=True)
    out_host = torch.empty(B, 3, H, H, dtype=torch.float32).pin_memory()

    if video:
        # configs[4]: this rank's frames (global frame f = chunk k of `--chunk` frames -> rank k % world,
        # reface_b200.shard.video_segments); ONE source face for the whole video
        def run(d):
            frames = swap_video(model, d["ref_img"][:1], d["tar_img"], d["inpaint_img"], d["mask_lat"], d["x_T"],
                                d["enc_noise"], landmarks136=d["landmarks136"], S=S, scale=args.scale, chunk=args.chunk)
            return torch.stack([frames[i] for i in range(B)])
    else:
        def run(d):
            return swap_faces(model, S=S, scale=args.scale, **d)["image"]

    def step_resident():
        return run(dev_in)

    def step_e2e():
        d = {k: v.to(dev, non_blocking=True) for k, v in host_in.items()}
        img = run(d)
        out_host.copy_(img, non_blocking=True)
        return img

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(3, args.warmup)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count
    ms = timed(step_resident, args.steps)
    launches = eng.launch_count - l0
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    faces = B * world * args.steps
    value = faces / (ms / 1e3)
    e2e_val = faces / (ms_e2e / 1e3)
    # roofline of the dominant kernel (gemm_tc_kernel: tcgen05 GEMM / implicit-GEMM conv / attention products):
    # one more step with every such launch bracketed by CUDA events on the launching stream
    eng.set_option("profile", 1)
    step_resident()
    g_ms, g_flops, g_n = eng.profile_read()
    eng.set_option("profile", 0)
    pk = peaks()
    achieved = g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
    h2d = sum(v.numel() * v.element_size() for v in host_in.values())
    d2h = out_host.numel() * out_host.element_size()
    fpf = FLOP_PER_VIDEO_FRAME if (video and (H, S) == (512, 30)) else (None if video else FLOP_PER_FACE.get((H, S)))
    traffic, traffic_src = traffic_from_profiles()
    line = dict(metric=metric, value=value, unit=unit, n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f16",
                data="synthetic", config=config,
                e2e=dict(value=e2e_val, unit=unit, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         ms_per_step=ms_e2e / args.steps),
                gpu_launches=int(launches),
                roofline=dict(bound="tensor", achieved=achieved, peak=pk["tflops"], unit="TFLOP/s",
                              frac=achieved / pk["tflops"], traffic=traffic, traffic_source=traffic_src,
                              kernel="gemm_persist_kernel / gemm_mcast_kernel (tcgen05 GEMM / implicit-GEMM conv) + attn_flash4/3_kernel (fused attention)",
                              alg_flop_per_unit=fpf,
                              launches_per_step=int(g_n), kernel_ms_per_step=g_ms, alg_tflop_per_step=g_flops / 1e12,
                              executed_tflop_per_step=eng.last_executed_flops / 1e12,
                              executed_frac=(eng.last_executed_flops / (g_ms / 1e3) / 1e12 / pk["tflops"]) if g_ms > 0 else None,
                              flops_note="achieved = ALGORITHMIC (reference-equivalent, SURVEY 8d) FLOPs of the tensor-core launches / "
                                         "their CUDA-event time; the folded upsample convs execute 4/9 of their reference MACs: "
                                         "executed_* gives the tensor cores' own rate",
                              share_of_step=g_ms / (ms / args.steps), peak_source=pk["source"] + ", sustained bf16",
                              whole_path_frac=(value / world) * fpf / (pk["tflops"] * 1e12) if fpf else None),
                clocks=sampler.summary() if rank == 0 else None, arena_peak_gb=eng.arena_peak / 2 ** 30)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args, 1, 1)[0]
            except Exception as e:  # the GPU number stands on its own
                line["cpu_baseline"] = dict(error=repr(e))
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
