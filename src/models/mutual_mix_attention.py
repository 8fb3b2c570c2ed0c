"""`src.models.mutual_mix_attention` of the reference -> mikudance_b200.reference_control."""
from mikudance_b200.reference_control import ReferenceAttentionControl, torch_dfs  # noqa: F401
