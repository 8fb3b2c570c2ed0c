"""Opt-in stand-in for the three `diffusers` names scripts/inference_video.py:10 imports, for environments without
diffusers (it is absent from the build image): put `<repo>/compat` on sys.path AFTER making sure the real package is
not installed — this module must never shadow a real diffusers.

    AutoencoderKL  -> mikudance_b200.vae.AutoencoderKL          (sm_100a kernels; same keys / encode / decode)
    DDIMScheduler  -> mikudance_b200.scheduler.DDIMScheduler    (the duck type the pipelines call)
    AutoencoderKLTemporalDecoder -> not provided (only used with --video_decoder): raises on construction
"""
from mikudance_b200.scheduler import DDIMScheduler  # noqa: F401
from mikudance_b200.vae import AutoencoderKL  # noqa: F401

__version__ = "0.24.0-mikudance-b200-standin"


class AutoencoderKLTemporalDecoder:
    def __init__(self, *a, **k):
        raise NotImplementedError("AutoencoderKLTemporalDecoder (the --video_decoder path) is not provided")

    @classmethod
    def from_pretrained(cls, *a, **k):
        raise NotImplementedError("AutoencoderKLTemporalDecoder (the --video_decoder path) is not provided")
