"""mikudance_b200 — B200 (sm_100a) native implementation of the MikuDance video-diffusion
denoising loop (UNet3DConditionModel forward + CFG + DDIM) behind the reference's Python call
surface.  All arithmetic runs in hand-written CUDA kernels loaded through the C ABI in
include/mdk.h; there is no PyTorch/CPU fallback on the product path."""

__version__ = "0.1.0"
