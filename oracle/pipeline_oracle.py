"""ORACLE (test infrastructure).  The denoising step loop of MikuDanceVideoPipeline.__call__
(src/pipelines/pipeline_mikudance.py:573-686; identical body in pipeline_stage2_vdo.py:574-697)
restated in fp32 on top of oracle.unet3d_oracle and oracle.ddim_oracle.  The reference UNet call
(:647-653) is replaced by its output: per-window feature banks supplied by the caller."""
from typing import Callable, Dict, Optional

import torch

from .context_oracle import uniform
from .ddim_oracle import DDIMOracle
from .unet3d_oracle import unet3d_forward


def denoise_loop(sd, cfg, latents: torch.Tensor, ctx: torch.Tensor, num_steps: int,
                 guidance_scale: float, banks_for_window: Callable[[list], Optional[Dict]],
                 context_frames: int = 30, context_stride: int = 1, context_overlap: int = 8,
                 scheduler: Optional[DDIMOracle] = None, trace: Optional[list] = None) -> torch.Tensor:
    """latents [1, 4, F, h, w]; ctx [2, L, D] = [uncond(zeros); cond] (:418-423)."""
    sch = scheduler or DDIMOracle()
    sch.set_timesteps(num_steps)
    latents = latents.float() * sch.init_noise_sigma                          # :206
    do_cfg = guidance_scale > 1.0                                              # :397
    F_ = latents.shape[2]
    for t in sch.timesteps:                                                    # :576
        noise_pred = torch.zeros((2 if do_cfg else 1,) + tuple(latents.shape[1:]))   # :577-584
        counter = torch.zeros(1, 1, F_, 1, 1)                                  # :585-589
        windows = uniform(0, num_steps, F_, context_frames, context_stride, context_overlap)  # :603-612
        for c in windows:                                                      # :624 (batch size 1)
            lat_in = latents[:, :, c].repeat(2 if do_cfg else 1, 1, 1, 1, 1)   # :626-630
            pred = unet3d_forward(sd, cfg, lat_in, int(t), ctx if do_cfg else ctx[1:],
                                  banks=banks_for_window(c), cfg_guidance=do_cfg)      # :655-660
            noise_pred[:, :, c] = noise_pred[:, :, c] + pred                   # :662-664
            counter[:, :, c] = counter[:, :, c] + 1
        if do_cfg:                                                             # :670-674
            u, cnd = (noise_pred / counter).chunk(2)
            noise = u + guidance_scale * (cnd - u)
        else:
            noise = noise_pred / counter
        latents = sch.step(noise, int(t), latents)                             # :676-678
        if trace is not None:
            trace.append(latents.clone())
    return latents
