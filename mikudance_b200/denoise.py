"""DenoiseLoop — the hot path itself: the per-step body of MikuDanceVideoPipeline.__call__
(src/pipelines/pipeline_mikudance.py:573-686 of the reference) as one CUDA graph per clip:

    zero accumulators -> for each context window: gather latents (NHWC) -> UNet3D forward ->
    scatter-add prediction -> [all-reduce over frame shards] -> window average + CFG + DDIM update

Everything step-dependent (timestep, DDIM coefficients) is read from device memory, so the same
captured graph is replayed for every step; per step the host only writes 1 + 4 scalars.

Frame sharding: rank r of G runs the UNet on frames [r*fl, (r+1)*fl) of every window (both CFG
branches); the motion modules exchange frames<->pixels (or all-gather K/V) over NCCL (engine._motion);
the window accumulators are summed across ranks once per step (2.4 MB at 768x768x16f) and the DDIM
update is replicated.  With an even G and guidance on (MDK_CFG_SPLIT=0 switches it off) the two CFG branches go
to the two halves of the ranks instead (sharding.plan_ranks): the motion-module exchange then spans G/2 ranks —
none at G = 2.  Windows need not divide by the rank count (sharding.frame_split: 30 frames = 8 + 8 + 7 + 7).
Banks keep only the cond half (the uncond images never read them): 0.83 GB per 16-frame window at 768x768.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from . import ops
from .context import get_context_scheduler
from .sharding import plan_ranks, shard_window, slice_bank

F16 = torch.float16


class DenoiseLoop:
    def __init__(self, unet, scheduler, guidance_scale: float = 3.5, context_schedule: str = "uniform",
                 context_frames: int = 30, context_stride: int = 1, context_overlap: int = 8,
                 process_group=None, use_cuda_graph: bool = True):
        self.unet = unet
        self.eng = unet.engine()
        self.dev = self.eng.dev
        self.scheduler = scheduler
        self.guidance_scale = float(guidance_scale)
        self.do_cfg = guidance_scale > 1.0                      # pipeline_mikudance.py:397
        self.ctx_sched = get_context_scheduler(context_schedule)
        self.context_frames, self.context_stride, self.context_overlap = (
            context_frames, context_stride, context_overlap)
        self.pg = process_group
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
            self.rank = dist.get_rank(process_group)
        else:
            self.world, self.rank = 1, 0
        import os
        plan = plan_ranks(self.rank, self.world, self.do_cfg, os.environ.get("MDK_CFG_SPLIT", "1") == "1")
        self.branch, self.sub_rank, self.sub_world = plan["branch"], plan["sub_rank"], plan["sub_world"]
        if self.branch >= 0:
            import torch.distributed as dist
            ranks = dist.get_process_group_ranks(process_group)
            half = self.sub_world
            # every rank creates both branch groups (new_group is collective), then keeps its own
            groups = [dist.new_group(ranks[b * half:(b + 1) * half]) for b in (0, 1)] if half > 1 else [None, None]
            self.eng.set_process_group(groups[self.branch], self.sub_rank, half)
        else:
            self.eng.set_process_group(process_group, self.rank, self.world)
        self.use_graph = use_cuda_graph
        self.graph = None

    # ------------------------------------------------------------------------------------------
    def prepare(self, latents: torch.Tensor, ctx: torch.Tensor, num_inference_steps: int,
                banks_for_window: Optional[Callable[[List[int]], Optional[Dict[str, torch.Tensor]]]] = None):
        """latents [1, 4, F, h, w] (device, fp16, updated in place by step());
        ctx [2, L, D] = [uncond; cond] (or [1, L, D] without guidance);
        banks_for_window(window) -> {attention path: [(nb * len(window)) , hw, C]} full-window banks
        (the stand-in for / output of the reference UNet; this rank slices its frames)."""
        dev = self.dev
        # the engine only exists on a CUDA device (UNetEngine refuses CPU models): same device required
        same_dev = latents.device.type == dev.type and (
            latents.device.index is None or dev.index is None or latents.device.index == dev.index)
        assert same_dev and latents.dtype == F16 and latents.is_contiguous(), (latents.device, dev, latents.dtype)
        self.latents = latents
        _, self.c, self.F, self.h, self.w = latents.shape
        self.nb = 2 if self.do_cfg else 1
        self.ctx = ctx.to(device=dev, dtype=F16).contiguous()
        # the context this rank's UNet calls see (CFG split: one branch); its cross-attention K / V^T once per clip
        self.ctx_run = self.ctx if self.branch < 0 else self.ctx[self.branch:self.branch + 1]
        self.eng.project_context(self.ctx_run)
        self.scheduler.set_timesteps(num_inference_steps)
        self.timesteps = [int(t) for t in self.scheduler.timesteps]
        # the reference always calls the scheduler with step=0 (pipeline_mikudance.py:603-612)
        self.windows = list(self.ctx_sched(0, num_inference_steps, self.F, self.context_frames,
                                           self.context_stride, self.context_overlap))
        self.win = []
        for wdw in self.windows:
            L = len(wdw)
            if len(set(int(i) for i in wdw)) != L:
                # the reference's indexed assignment (pipeline_mikudance.py:662-664) is last-write-wins with the
                # counter bumped once; the accumulate kernel adds without atomics and assumes distinct frames
                raise NotImplementedError(f"context window {list(wdw)} repeats a frame (context_stride > 1 with "
                                          "wrap-around): not supported")
            mine, lo = shard_window(wdw, self.sub_rank, self.sub_world)
            fl = len(mine)
            idx = torch.tensor(mine, dtype=torch.int32, device=dev)
            banks = None
            if banks_for_window is not None:
                full = banks_for_window(wdw)
                if full is not None and self.branch != 0:          # the uncond branch never reads a bank
                    # keep the cond half of (b fl) only: the rows that are read (uncond images never see a bank)
                    r0 = fl if self.do_cfg else 0
                    banks = {k: slice_bank(v, self.nb, L, self.sub_rank, self.sub_world)[r0:]
                             .to(device=dev, dtype=F16).contiguous() for k, v in full.items()}
            self.win.append(dict(frames=wdw, idx=idx, fl=fl, f_off=lo, L=L, banks=banks))
        self.acc = torch.zeros((self.nb, self.c, self.F, self.h, self.w), dtype=torch.float32, device=dev)
        self.counter = torch.zeros(self.F, dtype=torch.float32, device=dev)
        self.coef = torch.zeros(4, dtype=torch.float32, device=dev)
        # per-step scalars live in device tables; step() moves row i into the buffers the captured
        # graph reads (stream-ordered D2D copies: no host buffer is reused while a copy is in flight)
        self.t_table = torch.tensor(self.timesteps, dtype=torch.int64, device=dev)
        self.coef_table = torch.stack([self.scheduler.step_coefficients(t)[0] for t in self.timesteps]
                                      ).to(device=dev, dtype=torch.float32).contiguous()
        self.vpred = self.scheduler.config.prediction_type == "v_prediction"
        self.graph = None
        return self

    # ------------------------------------------------------------------------------------------
    def _step_body(self):
        eng = self.eng
        self.acc.zero_()
        self.counter.zero_()
        n_unc_all = self.do_cfg
        br = self.branch
        for wd in self.win:
            fl = wd["fl"]
            if br < 0:
                x_in = ops.latents_to_nhwc(self.latents, b=self.nb, frame_idx=wd["idx"], fl=fl,
                                           cpad=eng.cin_pad)
                pred = eng.run(x_in, self.nb, fl, self.h, self.w, self.ctx_run, wd["banks"],
                               n_uncond=(fl if n_unc_all else 0), f_off=wd["f_off"], f_total=wd["L"])
                ops.pred_accumulate(pred, self.acc, self.counter, frame_idx=wd["idx"], fl=fl)
            else:
                # CFG split: this rank evaluates ONE branch; its row of the accumulator gets the prediction,
                # the other row stays zero until the all-reduce; the uncond ranks own the frame counter
                x_in = ops.latents_to_nhwc(self.latents, b=1, frame_idx=wd["idx"], fl=fl, cpad=eng.cin_pad)
                pred = eng.run(x_in, 1, fl, self.h, self.w, self.ctx_run, wd["banks"],
                               n_uncond=(fl if br == 0 else 0), f_off=wd["f_off"], f_total=wd["L"])
                ops.pred_accumulate(pred, self.acc[br:br + 1], self.counter if br == 0 else None,
                                    frame_idx=wd["idx"], fl=fl)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.acc, group=self.pg)
            dist.all_reduce(self.counter, group=self.pg)
        ops.cfg_ddim_step(self.acc, self.counter, self.latents, self.coef, self.guidance_scale,
                          self.vpred)

    def _set_step_scalars(self, i: int):
        self.eng.t_dev.copy_(self.t_table[i:i + 1], non_blocking=True)
        self.coef.copy_(self.coef_table[i], non_blocking=True)

    def capture(self):
        """Warm up once eagerly (kernel attributes, allocator, NCCL), restore the latents, capture."""
        saved = self.latents.clone()
        self._set_step_scalars(0)
        self._step_body()
        torch.cuda.synchronize(self.dev)
        self.latents.copy_(saved)
        if self.use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_body()
            self.graph = g
            self.latents.copy_(saved)
        torch.cuda.synchronize(self.dev)

    def step(self, i: int):
        """One DDIM step (all windows); latents are updated in place."""
        self._set_step_scalars(i)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()

    def run(self, callback=None, callback_steps: int = 1):
        if self.graph is None and self.use_graph:
            self.capture()
        for i, t in enumerate(self.timesteps):
            self.step(i)
            if callback is not None and i % callback_steps == 0:
                callback(i, t, self.latents)
        return self.latents
