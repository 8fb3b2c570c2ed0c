"""FeedForward / GEGLU as in diffusers 0.24.0 models/attention.py (restated)."""
import torch.nn.functional as F
from torch import nn

from .attention_processor import Attention  # noqa: F401  (re-exported like the real module)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states, scale=1.0):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu",
                 final_dropout=False):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        if activation_fn != "geglu":
            raise NotImplementedError("stand-in: only geglu is on the hot path")
        self.net = nn.ModuleList([GEGLU(dim, inner_dim), nn.Dropout(dropout),
                                  nn.Linear(inner_dim, dim_out)])

    def forward(self, hidden_states, scale=1.0):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class AdaLayerNorm(nn.Module):  # imported by the reference, never constructed on this path
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("stand-in: not on the hot path")
