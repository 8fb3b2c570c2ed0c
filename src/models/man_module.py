"""`src.models.man_module` of the reference -> parameter container of mikudance_b200.unet_2d_ref."""
from mikudance_b200.unet_2d_ref import MANModule  # noqa: F401
