"""`src.models.unet_2d_mix` of the reference (the reference UNet / writer) -> mikudance_b200.unet_2d_ref."""
from mikudance_b200.unet_2d_ref import UNet2DConditionModel, UNet2DConditionOutput  # noqa: F401
