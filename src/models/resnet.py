"""`src.models.resnet` of the reference -> parameter containers of mikudance_b200.unet_3d."""
from mikudance_b200.unet_3d import InflatedConv3d, InflatedGroupNorm, ResnetBlock3D  # noqa: F401
