"""`src.models.unet_3d_mix` of the reference -> mikudance_b200.unet_3d."""
from mikudance_b200.unet_3d import UNet3DConditionModel, UNet3DConditionOutput  # noqa: F401
