"""MikuDanceVideoPipeline / Pose2VideoPipeline — the reference's pipeline call surface
(src/pipelines/pipeline_mikudance.py:36-704, src/pipelines/pipeline_stage2_vdo.py:36-712) around
the B200 denoising loop.

Only the step loop (pipeline_mikudance.py:573-686) is this repository's hot path; it runs in
`DenoiseLoop` (one CUDA graph per clip, sm_100a kernels).  Everything around it — CLIP image
embedding (:405-423), VAE encodes of the condition images (:455-549), the reference UNet that writes
the feature banks (:647-653) and the VAE decode (:115-150) — is SURVEY.md §8f "next": here those
stages call the user-supplied PyTorch modules (`vae`, `image_encoder`, `reference_unet`) exactly like
the reference does.  Two deliberate differences, neither changes results:
  * the reference UNet is evaluated once per context window instead of once per step per window:
    its inputs are step-invariant (timestep zeros_like(t), constant latents and embeddings), so is its
    output (the reference's stage-1 pipeline already hoists it,
    pipeline_stage1_img.py:348-359);
  * window accumulation / CFG / DDIM run in fp32 inside one fused kernel instead of fp16 eager ops.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Union

import numpy as np
import torch

from .denoise import DenoiseLoop
from .reference_control import ReferenceAttentionControl


@dataclass
class MikuDanceVideoPipelineOutput:
    videos: Union[torch.Tensor, np.ndarray]


@dataclass
class Pose2VideoPipelineOutput:
    videos: Union[torch.Tensor, np.ndarray]


def _preprocess_image(image, height: int, width: int, normalize: bool, multiple: int = 8) -> torch.Tensor:
    """VaeImageProcessor.preprocess for a PIL image: RGB, Lanczos resize to (width, height) rounded down
    to a multiple of the VAE scale factor, [0,1] float CHW, optionally mapped to [-1,1]."""
    from PIL import Image
    if isinstance(image, torch.Tensor):
        t = image if image.dim() == 4 else image.unsqueeze(0)
        return (2.0 * t - 1.0) if normalize else t
    image = image.convert("RGB")
    height, width = height - height % multiple, width - width % multiple
    image = image.resize((width, height), resample=Image.LANCZOS)
    arr = np.asarray(image).astype(np.float32) / 255.0
    t = torch.from_numpy(arr).permute(2, 0, 1).unsqueeze(0)
    return (2.0 * t - 1.0) if normalize else t


class MikuDanceVideoPipeline:
    _default_context_frames = 30
    _output_cls = MikuDanceVideoPipelineOutput

    def __init__(self, vae, image_encoder, reference_unet, denoising_unet, scheduler,
                 image_proj_model=None, tokenizer=None, text_encoder=None, video_decoder=False):
        self.vae, self.image_encoder = vae, image_encoder
        self.reference_unet, self.denoising_unet = reference_unet, denoising_unet
        self.scheduler = scheduler
        self.image_proj_model, self.tokenizer, self.text_encoder = image_proj_model, tokenizer, text_encoder
        self.video_decoder = video_decoder
        boc = getattr(getattr(vae, "config", None), "block_out_channels", (0, 0, 0, 0))
        self.vae_scale_factor = 2 ** (len(boc) - 1)
        self.decode_chunk_size = 16
        self.process_group = None      # set to a torch.distributed group to shard frames over GPUs
        self._device = None

    # -- plumbing -------------------------------------------------------------------------------
    def to(self, device=None, dtype=None):
        for m in (self.vae, self.image_encoder, self.reference_unet, self.denoising_unet):
            if m is not None and hasattr(m, "to"):
                m.to(device) if dtype is None else m.to(device, dtype=dtype)
        self._device = torch.device(device) if device is not None else self._device
        return self

    @property
    def device(self):
        return self._device if self._device is not None else self.denoising_unet.device

    def progress_bar(self, total):
        try:
            from tqdm.auto import tqdm
            return tqdm(total=total)
        except Exception:  # noqa: BLE001
            import contextlib
            return contextlib.nullcontext(type("P", (), {"update": lambda self, n=1: None})())

    def prepare_latents(self, batch_size, num_channels_latents, width, height, video_length, dtype,
                        device, generator, latents=None):
        shape = (batch_size, num_channels_latents, video_length, height // self.vae_scale_factor,
                 width // self.vae_scale_factor)
        if latents is None:
            gdev = generator.device if generator is not None else torch.device("cpu")
            # diffusers randn_tensor: draw on the generator's device (CPU for scripts/inference_video.py:104)
            latents = torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def _vae_latent(self, x: torch.Tensor) -> torch.Tensor:
        vdt = next(self.vae.parameters()).dtype if hasattr(self.vae, "parameters") else x.dtype
        x = x.to(dtype=vdt, device=self.device)
        return self.vae.encode(x).latent_dist.mean * 0.18215

    def decode_latents(self, latents: torch.Tensor) -> np.ndarray:
        """per-frame VAE decode (pipeline_mikudance.py:115-130)."""
        b, c, f, h, w = latents.shape
        lat = (latents / 0.18215).permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        frames = []
        for i in range(lat.shape[0]):
            fr = self.vae.decode(lat[i:i + 1]).sample
            frames.append((fr / 2 + 0.5).clamp(0, 1))
        video = torch.cat(frames).reshape(b, f, -1, frames[0].shape[-2], frames[0].shape[-1])
        return video.permute(0, 2, 1, 3, 4).cpu().float().numpy()

    # -- the call -------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, ref_image, ref_skel_image, tgt_pose_images, tgt_face_images, tgt_hand_images,
                 scene_motion_npy, width, height, video_length, num_inference_steps, guidance_scale,
                 num_images_per_prompt=1, eta: float = 0.0, generator=None, output_type: Optional[str] = "tensor",
                 return_dict: bool = True, callback: Optional[Callable] = None, callback_steps: Optional[int] = 1,
                 context_schedule="uniform", context_frames=None, context_stride=1, context_overlap=8,
                 context_batch_size=1, interpolation_factor=1, latents=None, **kwargs):
        if eta != 0.0:
            raise NotImplementedError("eta > 0 is not used by the reference's DDIM configuration")
        if self.video_decoder:
            # the reference decodes with AutoencoderKLTemporalDecoder.decode(num_frames=) in chunks of 16
            # (pipeline_mikudance.py:132-150, 692-695); that model is outside the hot path and not provided:
            # never fall back silently to the per-frame decoder
            raise NotImplementedError("video_decoder=True (temporal VAE decoder) is not implemented; "
                                      "construct the pipeline with video_decoder=False")
        if num_images_per_prompt != 1:
            raise NotImplementedError("num_images_per_prompt != 1: the denoising loop works on a latent batch of 1 "
                                      "(as scripts/inference_video.py calls it)")
        if context_batch_size != 1:
            raise NotImplementedError("context_batch_size != 1 (the reference always uses 1)")
        if interpolation_factor >= 2:
            raise NotImplementedError("latent interpolation (interpolation_factor >= 2) is outside the hot path")
        context_frames = context_frames or self._default_context_frames
        device = self.device
        do_cfg = guidance_scale > 1.0
        dtype = self.denoising_unet.dtype

        # CLIP image embedding of the reference image (:405-423)
        from transformers import CLIPImageProcessor
        clip_image = CLIPImageProcessor().preprocess(ref_image.resize((224, 224)), return_tensors="pt").pixel_values
        enc_dtype = next(self.image_encoder.parameters()).dtype
        emb = self.image_encoder(clip_image.to(device, dtype=enc_dtype)).last_hidden_state
        emb = self.image_encoder.vision_model.post_layernorm(emb)
        image_prompt_embeds = self.image_encoder.visual_projection(emb)          # [1, 257, 768]
        if do_cfg:
            image_prompt_embeds = torch.cat([torch.zeros_like(image_prompt_embeds), image_prompt_embeds], 0)

        writer = ReferenceAttentionControl(self.reference_unet, do_classifier_free_guidance=do_cfg,
                                           mode="write", batch_size=1, fusion_blocks="full")
        reader = ReferenceAttentionControl(self.denoising_unet, do_classifier_free_guidance=do_cfg,
                                           mode="read", batch_size=1, fusion_blocks="full")

        latents = self.prepare_latents(num_images_per_prompt, self.denoising_unet.in_channels, width, height,
                                       video_length, image_prompt_embeds.dtype, device, generator, latents)

        # condition latents (:455-569): ref image, ref skeleton, per-frame pose / face / hand, scene motion
        vsf = self.vae_scale_factor
        F_ = video_length

        def rep(x):     # (1, d, h, w) -> (F, d, h, w)
            return x.repeat(F_, 1, 1, 1)

        ref_lat = rep(self._vae_latent(_preprocess_image(ref_image, height, width, True, vsf)))
        skel_lat = rep(self._vae_latent(_preprocess_image(ref_skel_image, height, width, False, vsf)))

        def per_frame(images):
            return torch.cat([self._vae_latent(_preprocess_image(im, height, width, False, vsf))
                              for im in images], 0)

        pose_lat, face_lat, hand_lat = per_frame(tgt_pose_images), per_frame(tgt_face_images), per_frame(tgt_hand_images)
        motion = torch.from_numpy(scene_motion_npy).to(dtype=ref_lat.dtype, device=device)   # [F, 2, h, w]
        ref_latents = torch.cat([ref_lat, skel_lat, pose_lat, face_lat, hand_lat, motion], dim=1)  # [F, 22, h, w]

        # the loop: reference UNet once per window (hoisted), then the CUDA-graph denoising loop
        t0 = torch.zeros((), dtype=torch.int64, device=device)
        nb = 2 if do_cfg else 1

        def banks_for_window(window: List[int]):
            f = len(window)
            ref_in = ref_latents[window].repeat(nb, 1, 1, 1)                       # (b f) d h w, b-major
            emb_in = image_prompt_embeds.repeat((f, 1, 1))                         # tiled [u,c,u,c..] (:645)
            self.reference_unet(ref_in, torch.zeros_like(t0), encoder_hidden_states=emb_in, return_dict=False)
            reader.update(writer)
            banks = {}
            names = {id(m): n for n, m in self.denoising_unet.named_modules()}
            for blk in self.denoising_unet.spatial_blocks():
                if len(blk.bank) == 1:
                    banks[names[id(blk)].rsplit(".transformer_blocks", 1)[0]] = blk.bank[0].clone()
            reader.clear()
            writer.clear()
            return banks

        loop = DenoiseLoop(self.denoising_unet, self.scheduler, guidance_scale=guidance_scale,
                           context_schedule=context_schedule, context_frames=context_frames,
                           context_stride=context_stride, context_overlap=context_overlap,
                           process_group=self.process_group)
        lat16 = latents.to(device=device, dtype=torch.float16).contiguous()
        loop.prepare(lat16, image_prompt_embeds, num_inference_steps, banks_for_window)
        with self.progress_bar(total=num_inference_steps) as bar:
            def cb(i, t, lat):
                bar.update()
                if callback is not None and i % callback_steps == 0:
                    callback(i, t, lat)
            latents = loop.run(callback=cb, callback_steps=1).to(dtype)

        images = self.decode_latents(latents)
        if output_type == "tensor":
            images = torch.from_numpy(images)
        if not return_dict:
            return images
        return self._output_cls(videos=images)


class Pose2VideoPipeline(MikuDanceVideoPipeline):
    """src/pipelines/pipeline_stage2_vdo.py: same body, `context_frames` default 32, no video decoder."""
    _default_context_frames = 32
    _output_cls = Pose2VideoPipelineOutput

    def __init__(self, vae, image_encoder, reference_unet, denoising_unet, scheduler,
                 image_proj_model=None, tokenizer=None, text_encoder=None):
        super().__init__(vae, image_encoder, reference_unet, denoising_unet, scheduler, image_proj_model,
                         tokenizer, text_encoder, video_decoder=False)
