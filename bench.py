#!/usr/bin/env python
"""bench.py — denoised frames/s of the MikuDance denoising loop on B200 (BASELINE.json metric).

  python bench.py --gpus 1 --steps 20 --warmup 3              # this repo's sm_100a path
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # frame-sharded
  python bench.py --impl reference ...                          # the reference algorithm on host cores

A "step" is one DDIM denoising step of the whole clip: every context window's UNet3DConditionModel
forward (CFG: 2 x frames images), the window average, classifier-free guidance and the DDIM update.
value = frames / (num_inference_steps * seconds_per_step): the clip's frames divided by the time of
its full 20-step loop, with the latents / weights / reference banks already resident in HBM.
e2e   = the same through DenoiseLoop.step with HOST latents: pinned H2D of the step's inputs and a
        D2H read of the step's result inside the timed region.
Synthetic data: random-init weights of the SD-1.5 + motion-module architecture (seeded per tensor,
motion proj_out NOT zero), seeded latents / CLIP context / reference banks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (latent side, frames, num_inference_steps, context_frames)
    "A": (32, 4, 2, 30),     # 256x256, 4 frames, 2 steps (plumbing)
    "B": (96, 16, 20, 30),   # 768x768, 16 frames, 20 steps  <- headline
    "C": (96, 32, 20, 32),   # 768x768, 32 frames, one 32-frame window
    "D": (96, 64, 20, 32),   # 768x768, 64 frames, sliding windows
    "E": (128, 16, 50, 30),  # 1024x1024, 16 frames, 50 steps
}
GUIDANCE = 3.5
SCHED_KW = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="linear", clip_sample=False,
                steps_offset=1, prediction_type="v_prediction", rescale_betas_zero_snr=True,
                timestep_spacing="trailing")
METRIC = "denoised frames/sec at 768x768x16f x20 DDIM steps"


def per_image_flops(h, f):
    """Deduplicated algorithmic FLOPs per image per UNet forward."""
    hw = [h * h, (h // 2) ** 2, (h // 4) ** 2, (h // 8) ** 2]
    t = dict(conv=0, lin=0, self=0, cross=0, temp=0)

    def resnet(ci, co, n):
        t["conv"] += 18 * ci * co * n + 18 * co * co * n + (2 * ci * co * n if ci != co else 0)

    def t3d(c, n):
        t["lin"] += 40 * c * c * n
        t["self"] += 4 * n * n * c
        t["cross"] += 4 * n * 257 * c

    def mm(c, n):
        t["lin"] += 44 * c * c * n
        t["temp"] += 8 * f * c * n

    t["conv"] += 2 * 18 * 4 * 320 * hw[0]
    for ci, co, n, a in [(320, 320, hw[0], 1), (320, 640, hw[1], 1), (640, 1280, hw[2], 1), (1280, 1280, hw[3], 0)]:
        for i in range(2):
            resnet(ci if i == 0 else co, co, n)
            if a:
                t3d(co, n)
            mm(co, n)
    t["conv"] += 18 * (320 * 320 * hw[1] + 640 * 640 * hw[2] + 1280 * 1280 * hw[3])
    resnet(1280, 1280, hw[3]); t3d(1280, hw[3]); mm(1280, hw[3]); resnet(1280, 1280, hw[3])
    for co, cis, n, a in [(1280, [2560] * 3, hw[3], 0), (1280, [2560, 2560, 1920], hw[2], 1),
                          (640, [1920, 1280, 960], hw[1], 1), (320, [960, 640, 640], hw[0], 1)]:
        for ci in cis:
            resnet(ci, co, n)
            if a:
                t3d(co, n)
            mm(co, n)
    t["conv"] += 18 * (1280 * 1280 * hw[2] + 1280 * 1280 * hw[1] + 640 * 640 * hw[0])
    return t


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops_sustained=d.get("bf16_tflops_sustained", 1391.7), tflops=d.get("bf16_tflops", 1621.9),
                    hbm_gbs=d.get("hbm_gbs", 6567.7), source="measured (MEASURED_PEAKS.json)")
    return dict(tflops_sustained=1400.0, tflops=1590.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return dict(sm_mhz=statistics.median(busy) if busy else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def build_model(cfg, dev):
    from mikudance_b200 import synth
    from mikudance_b200.unet_3d import UNet3DConditionModel
    model = UNet3DConditionModel(
        sample_size=64, in_channels=4, out_channels=4, block_out_channels=cfg["block_out_channels"],
        layers_per_block=2, cross_attention_dim=cfg["cross_attention_dim"], attention_head_dim=8,
        norm_num_groups=32, norm_eps=1e-5, use_inflated_groupnorm=True, use_motion_module=True,
        motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=True, motion_module_type="Vanilla",
        motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                                  attention_block_types=("Temporal_Self", "Temporal_Self"),
                                  temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                                  temporal_attention_dim_div=1),
        unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
    model.load_state_dict(synth.synthetic_state_dict(cfg, seed=0))
    return model.to(device=dev, dtype=torch.float16).eval()


def host_threads():
    """Threads the CPU arm may use: the affinity mask capped by the cgroup CPU quota (an oversubscribed
    OpenMP pool on a quota-limited container collapses: measured 4.5 GFLOP/s with 128 threads), <= 64."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            with open(path) as f:
                parts = f.read().split()
            if path.endswith("cpu.max"):
                if parts[0] != "max":
                    n = min(n, max(1, int(int(parts[0]) / int(parts[1]))))
            else:
                q = int(parts[0])
                if q > 0:
                    with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f2:
                        n = min(n, max(1, q // int(f2.read().split()[0])))
            break
        except (OSError, ValueError, IndexError):
            continue
    return max(1, min(n, 64))


REFERENCE_ROOT = "/root/reference"


class CpuSample:
    """The reference's CPU implementation of the path on a BOUNDED, FIXED sample of the workload: ONE frame (both
    CFG branches = 2 images) of one UNet3DConditionModel forward at the config's own latent size (never a size
    chosen at run time), fp32, on the host cores.  Runs the reference's unmodified modules through
    oracle/diffusers_standin when /root/reference is mounted (kind "reference"; this container), else the oracle
    port restating them (kind "port"; the GPU box).  One method for `cpu_baseline` and for `--impl reference`.
    A clip step costs `frames` such samples (per window), the clip `num_inference_steps x frames`."""

    def __init__(self, h):
        from mikudance_b200 import synth
        self.h = h
        self.cores = host_threads()
        torch.set_num_threads(self.cores)
        cfg = synth.SD15_CONFIG
        sd = {k: v.float() for k, v in synth.synthetic_state_dict(cfg, seed=0).items()}
        self.x, self.ctx = synth.synthetic_inputs(cfg, 2, 1, h, h, lctx=257)
        banks = synth.synthetic_banks(cfg, 2, h, h)
        self.kind = "port"
        if os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models")) and os.environ.get("MDK_CPU_PORT", "0") != "1":
            try:
                sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_standin"))
                sys.path.insert(1, REFERENCE_ROOT)
                sys.path.insert(2, os.path.join(ROOT, "oracle"))
                for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
                    del sys.modules[k]            # the repo's own src/ shims must not shadow the reference's
                import make_golden
                model = make_golden.build_reference_unet(cfg)
                model.load_state_dict(sd)
                make_golden.install_banks(model, banks, cfg, do_cfg=True)
                self._fn = lambda: model(self.x, torch.tensor(499), encoder_hidden_states=self.ctx,
                                         return_dict=False)[0]
                self.kind = "reference"
            except Exception as e:  # noqa: BLE001 — fall back to the port, say why
                print(f"bench: reference modules unavailable ({type(e).__name__}: {e}); using the oracle port",
                      file=sys.stderr)
        if self.kind == "port":
            from oracle import unet3d_oracle as O
            self._fn = lambda: O.unet3d_forward(sd, cfg, self.x, 499, self.ctx, banks=banks, cfg_guidance=True)

    def run(self):
        """seconds of one sample (2 images of one UNet forward at latent h x h)"""
        with torch.no_grad():
            t0 = time.perf_counter()
            self._fn()
            return time.perf_counter() - t0

    def describe(self, n_timed):
        what = ("the reference's own modules (src/models/unet_3d_mix.py ... through oracle/diffusers_standin)"
                if self.kind == "reference" else "fp32 oracle restating the reference modules (oracle/unet3d_oracle.py)")
        return (f"1 frame (2 CFG images) of one UNet forward at latent {self.h}x{self.h} per timed step, {n_timed} timed, "
                f"{self.cores} threads, fp32, {what}")


def time_reference_unet(cfg, dev, h, f_win, n_windows, num_steps, ms_step):
    """One forward of the native reference UNet (mikudance_b200.unet_2d_ref) on a
    window's 2 x f_win images: random-init SD-1.5-size weights, synthetic condition latents, the tiled
    [uncond, cond, ...] CLIP context.  CUDA events, 1 warm-up + 3 timed forwards."""
    from mikudance_b200 import synth
    from mikudance_b200.unet_2d_ref import UNet2DConditionModel
    m = UNet2DConditionModel(block_out_channels=cfg["block_out_channels"],
                             cross_attention_dim=cfg["cross_attention_dim"])
    m.load_state_dict(synth.synthetic_state_dict(cfg, seed=0, reference_unet=True))
    m = m.to(device=dev, dtype=torch.float16).eval()
    n_img = 2 * f_win
    x, ctx = synth.synthetic_reference_inputs(cfg, n_img, h, h, lctx=257)
    x, ctx = x.to(dev, torch.float16), ctx.to(dev, torch.float16)
    eng = m.engine()
    eng.set_timestep(0)
    eng.run(x, ctx)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.run(x, ctx)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / 3
    clip_ms = num_steps * ms_step + n_windows * ms
    del m, eng
    torch.cuda.empty_cache()
    return dict(ms_per_window=ms, images=n_img, windows=n_windows, hoisted=True,
                share_of_clip=n_windows * ms / clip_ms,
                note="runs once per context window per clip (its inputs are step-invariant); the reference "
                     "runs it once per window per step")


def sharded_parity_check(dev, pg, world):
    """N > 1 only, before the timed region: the frame-sharded loop (NCCL, CUDA graph) against the single-GPU loop on
    identical inputs — tiny UNet3D, 2 DDIM steps, a window whose length does NOT divide by the rank count.  Every
    rank holds the replicated latents; the worst rank is reported as `parity_vs_single`."""
    import torch.distributed as dist
    from mikudance_b200 import synth
    from mikudance_b200.denoise import DenoiseLoop
    from mikudance_b200.scheduler import DDIMScheduler
    from mikudance_b200.unet_3d import UNet3DConditionModel
    cfg = synth.TINY_CONFIG
    m = UNet3DConditionModel(block_out_channels=cfg["block_out_channels"],
                             cross_attention_dim=cfg["cross_attention_dim"], use_inflated_groupnorm=True,
                             use_motion_module=True, motion_module_mid_block=True, motion_module_type="Vanilla",
                             motion_module_kwargs=dict(temporal_position_encoding=True,
                                                       temporal_position_encoding_max_len=32),
                             unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
    m.load_state_dict(synth.synthetic_state_dict(cfg, seed=0))
    m = m.to(device=dev, dtype=torch.float16).eval()
    F_, h = 2 * world + 1, 16
    lat, ctx = synth.synthetic_inputs(cfg, 2, F_, h, h, lctx=9)
    lat = lat[:1].half()

    def banks_for_window(wdw):
        return synth.synthetic_banks(cfg, 2 * len(wdw), h, h, seed=300 + wdw[0])

    res = {}
    for name, g in (("single", None), ("sharded", pg)):
        loop = DenoiseLoop(m, DDIMScheduler(**SCHED_KW), guidance_scale=GUIDANCE, context_frames=30,
                           context_stride=1, context_overlap=8, process_group=g, use_cuda_graph=True)
        loop.prepare(lat.to(dev).contiguous().clone(), ctx, 2, banks_for_window)
        res[name] = loop.run().float()
        mode = dict(cfg_split=loop.branch >= 0, exchange_ranks=loop.sub_world,
                    frames_per_rank=[w["fl"] for w in loop.win])
        torch.cuda.synchronize(dev)
        loop.graph = None
        dist.barrier()
    rel = ((res["sharded"] - res["single"]).norm() / res["single"].norm()).reshape(1)
    exact = torch.tensor([1.0 if torch.equal(res["sharded"], res["single"]) else 0.0], device=dev)
    dist.all_reduce(rel, op=dist.ReduceOp.MAX)
    dist.all_reduce(exact, op=dist.ReduceOp.MIN)
    m.engine().project_context(None)
    del m
    torch.cuda.empty_cache()
    return dict(rel_l2=float(rel), bit_exact=bool(exact.item() > 0.5), frames=F_, steps=2, model="tiny UNet3D",
                ranks=world, **mode)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="B", choices=sorted(CONFIGS))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-profile", action="store_true")
    ap.add_argument("--skip-reference-unet", action="store_true",
                    help="do not time the hoisted reference UNet (writer) forward that produces the banks")
    ap.add_argument("--frames", type=int, default=0,
                    help="experiments only: override the clip's frame count (e.g. 2 = the per-rank workload of "
                         "config B on 8 GPUs); the line's config states the frames actually run")
    ap.add_argument("--ncu-step", action="store_true",
                    help="run ONE eager step inside a cudaProfilerStart/Stop range and exit "
                         "(for `ncu --profile-from-start off`; prints no bench line)")
    args = ap.parse_args()
    if args.ncu_step:
        args.no_graph = True
    h, F_, num_steps, ctx_frames = CONFIGS[args.config]
    if args.frames > 0:
        F_ = args.frames
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = dict(workload=f"config {args.config}: {h * 8}x{h * 8}, {F_} frames, {num_steps} DDIM steps, CFG 3.5, "
                             f"UNet3D SD-1.5+motion (1.31 B params) random-init, synthetic latents/context/banks",
                    frames=F_, latent=h, num_inference_steps=num_steps, context_frames=ctx_frames,
                    images_per_unet_call=2 * min(F_, ctx_frames))

    from mikudance_b200.context import get_context_scheduler
    windows = [len(w) for w in get_context_scheduler("uniform")(0, num_steps, F_, ctx_frames, 1, 8)]
    # `config` is identical in both arms (same workload); how each arm executes it is under `execution`
    config = dict(workload, windows=windows,
                  l2="working set >> L2: 2.6 GB weights + banks streamed every step")

    if args.impl == "reference":
        if rank != 0:
            return
        smp = CpuSample(h)
        for _ in range(max(0, args.warmup)):
            smp.run()
        times = [smp.run() for _ in range(max(1, args.steps))]
        t = sum(times) / len(times)          # measured seconds per timed step = per 1-frame sample
        fps = 1.0 / (num_steps * t)          # 1 frame needs num_steps such UNet evaluations
        line = dict(metric=METRIC, value=fps, unit="frames/s", n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=t * 1e3, higher_is_better=True, scaling="strong",
                    vs_baseline=None, dtype="f32", data="synthetic", impl="reference", config=config,
                    execution=dict(parallelism=f"host cores ({smp.cores} threads), no GPU", cuda_graph=False,
                                   step="one timed step = ONE frame of the clip (2 CFG images) through one UNet "
                                        "forward; a clip step is `frames` of them per window"),
                    cpu_baseline=dict(value=fps, unit="frames/s", cores=smp.cores, kind=smp.kind,
                                      sample=smp.describe(len(times)),
                                      extrapolated=dict(clip_step_s=t * sum(windows), clip_s=t * sum(windows) * num_steps,
                                                        note="clip figures = sample x frame evaluations per step "
                                                             "x steps; value itself is frames / (steps x measured "
                                                             "sample time): frames/s is linear in frames")),
                    e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------------------------------
    import torch.distributed as dist
    from mikudance_b200 import _lib, ops, synth
    from mikudance_b200.denoise import DenoiseLoop
    from mikudance_b200.scheduler import DDIMScheduler

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    parity = sharded_parity_check(dev, pg, world) if world > 1 else None
    cfg = synth.SD15_CONFIG
    model = build_model(cfg, dev)
    sched = DDIMScheduler(**SCHED_KW)
    lat0, ctx = synth.synthetic_inputs(cfg, 2, F_, h, h, lctx=257)
    lat0 = lat0[:1].to(torch.float16)
    latents = lat0.to(dev).contiguous()

    def banks_for_window(wdw):
        return synth.synthetic_banks(cfg, 2 * len(wdw), h, h)

    loop = DenoiseLoop(model, sched, guidance_scale=GUIDANCE, context_frames=ctx_frames,
                       context_stride=1, context_overlap=8, process_group=pg,
                       use_cuda_graph=not args.no_graph)
    loop.prepare(latents, ctx, num_steps, banks_for_window)
    n_launch0 = _lib.launch_count()
    loop.capture()
    launches_per_step = (_lib.launch_count() - n_launch0) // (2 if not args.no_graph else 1)
    frame_evals = sum(len(w) for w in loop.windows)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i % num_steps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / k], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    if args.ncu_step:
        loop.step(0)                      # warm (tensor maps, allocator)
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        loop.step(1)
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        print(json.dumps(dict(ncu_step=True, launches_per_step=launches_per_step)))
        return
    # ---- device-resident timing ----
    for i in range(args.warmup):
        loop.step(i % num_steps)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step = timed(loop.step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = F_ / (num_steps * ms_step / 1e3)

    # ---- end to end with host buffers ----
    lat_host = lat0.clone().pin_memory()
    out_host = torch.empty_like(lat_host).pin_memory()

    def e2e_step(i):
        loop.latents.copy_(lat_host, non_blocking=True)
        loop.step(i)
        out_host.copy_(loop.latents, non_blocking=True)

    for i in range(min(2, args.warmup)):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)
    h2d = lat_host.numel() * 2 + 8 + 16
    d2h = out_host.numel() * 2

    # ---- per-kernel profile of one eager step (CUDA events around every launch) ----
    peaks = measured_peaks()
    roofline, kernels, shapes = None, None, None
    if not args.skip_profile:
        prof = ops.Profiler()
        ops.set_profiler(prof)
        loop._set_step_scalars(0)
        loop._step_body()
        torch.cuda.synchronize(dev)
        ops.set_profiler(None)
        kernels = prof.summary()
        shapes = prof.shapes(24)
        g = kernels.get("gemm_tc")
        if g and g["ms"] > 0:
            ach = g["flops"] / (g["ms"] * 1e-3) / 1e12
            roofline = dict(kernel="gemm_tc_kernel (tcgen05 GEMM + implicit-GEMM conv, all launches of one step)",
                            bound="tensor", achieved=ach, peak=peaks["tflops_sustained"], unit="TFLOP/s",
                            frac=ach / peaks["tflops_sustained"],
                            traffic=None,   # dram bytes are not measured inside this run (ncu: profiles/)
                            algorithmic_bytes_per_launch=g["bytes"] / g["n"],
                            launches=g["n"], share_of_step=g["ms"] / sum(k["ms"] for k in kernels.values()),
                            peak_source=peaks["source"] + ", sustained bf16 GEMM")
    step_tflop = frame_evals * 2 * sum(per_image_flops(h, min(F_, ctx_frames)).values()) / 1e12 \
        if cfg is synth.SD15_CONFIG else None
    step_ach = step_tflop / (ms_step * 1e-3) * 1.0 if step_tflop else None

    # ---- the hoisted reference UNet (writer): once per context window per CLIP, not per step ----
    # (outside the step metric; reported so the whole loop's cost is visible: the reference runs it every
    # step for every window, src/pipelines/pipeline_mikudance.py:647-653)
    refunet = None
    if rank == 0 and world == 1 and not args.skip_reference_unet and not args.ncu_step:
        try:
            refunet = time_reference_unet(cfg, dev, h, min(F_, ctx_frames), len(loop.windows), num_steps, ms_step)
        except Exception as e:  # noqa: BLE001 — never lose the bench line to the optional extra
            refunet = dict(error=f"{type(e).__name__}: {e}")

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        smp = CpuSample(h)
        smp.run()                             # warm-up (thread pool, allocator)
        t = smp.run()
        cpu = dict(value=1.0 / (num_steps * t), unit="frames/s", cores=smp.cores, kind=smp.kind,
                   sample=smp.describe(1) + f" ({t:.1f} s)",
                   extrapolated=dict(clip_step_s=t * frame_evals, clip_s=t * frame_evals * num_steps))

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="frames/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling="strong",
                    vs_baseline=None, dtype="f16", data="synthetic",
                    config=config,
                    execution=dict(parallelism=(f"{world} GPU(s): CFG branches on the two halves of the ranks, frames of "
                                                f"each window sharded over {loop.sub_world} rank(s)" if loop.branch >= 0
                                                else f"frames of each window sharded over {world} GPU(s)"),
                                   frames_per_rank=[w["fl"] for w in loop.win], cuda_graph=not args.no_graph),
                    clocks=clocks,
                    e2e=dict(value=F_ / (num_steps * ms_e2e / 1e3), unit="frames/s", ms_per_step=ms_e2e,
                             h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             note="per step: the latents in (pinned H2D) and out (D2H) + the step's scalars; the CLIP "
                                  "context and the reference banks (stand-in for the hoisted reference-UNet "
                                  "output) are uploaded once per clip, like the weights"),
                    gpu_launches=launches_per_step * args.steps, launches_per_step=launches_per_step,
                    roofline=dict(kernel="whole denoising step (all kernels; deduplicated algorithmic FLOPs, BASELINE.md)",
                                  bound="tensor", achieved=step_ach, peak=peaks["tflops_sustained"] * world, unit="TFLOP/s",
                                  frac=(step_ach / (peaks["tflops_sustained"] * world)) if step_ach else None, traffic=None,
                                  peak_per_gpu=peaks["tflops_sustained"],
                                  algorithmic_tflop_per_step=step_tflop,
                                  peak_source=peaks["source"] + ", sustained bf16 GEMM"),
                    roofline_gemm=roofline,
                    parity_vs_single=parity,
                    reference_unet=refunet, kernels=kernels, top_shapes=shapes, cpu_baseline=cpu)
        print(json.dumps(line))
    if world > 1:
        # drop the captured graph (it holds NCCL work) before tearing the communicator down, and never
        # let teardown hang the benchmark: the result line is already out
        sys.stdout.flush()
        loop.graph = None
        torch.cuda.synchronize(dev)
        try:
            dist.barrier()
        except Exception:  # noqa: BLE001
            pass
        os._exit(0)


if __name__ == "__main__":
    main()
