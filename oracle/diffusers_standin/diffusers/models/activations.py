"""get_activation as in diffusers 0.24.0 models/activations.py (restated)."""
from torch import nn

ACTIVATION_FUNCTIONS = {"swish": nn.SiLU, "silu": nn.SiLU, "mish": nn.Mish, "gelu": nn.GELU, "relu": nn.ReLU}


def get_activation(act_fn: str) -> nn.Module:
    act_fn = act_fn.lower()
    if act_fn not in ACTIVATION_FUNCTIONS:
        raise ValueError(f"Unsupported activation function: {act_fn}")
    return ACTIVATION_FUNCTIONS[act_fn]()
