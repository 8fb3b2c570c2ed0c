"""`src.models.attention` of the reference -> mikudance_b200.unet_3d."""
from mikudance_b200.unet_3d import TemporalBasicTransformerBlock  # noqa: F401
from mikudance_b200.unet_2d_ref import BasicTransformerBlock  # noqa: F401,E402
