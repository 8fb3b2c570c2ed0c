"""LoRACompatibleConv / LoRACompatibleLinear (diffusers 0.24.0 models/lora.py) without a LoRA layer:
plain Conv2d / Linear whose forward accepts and ignores the `scale` argument."""
from torch import nn


class LoRACompatibleConv(nn.Conv2d):
    def forward(self, hidden_states, scale: float = 1.0):
        return super().forward(hidden_states)


class LoRACompatibleLinear(nn.Linear):
    def forward(self, hidden_states, scale: float = 1.0):
        return super().forward(hidden_states)
