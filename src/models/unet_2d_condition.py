"""`src.models.unet_2d_condition` of the reference: scripts/inference_video.py:81-85 loads the base SD-1.5
UNet with it only to pass it to `unet_2d_mix.UNet2DConditionModel.from_unet` -> a weights carrier."""
from mikudance_b200.unet_2d_ref import UNet2DWeights as UNet2DConditionModel  # noqa: F401
