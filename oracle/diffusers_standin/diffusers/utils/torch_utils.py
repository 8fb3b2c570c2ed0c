import torch


def apply_freeu(resolution_idx, hidden_states, res_hidden_states, **freeu_kwargs):
    raise NotImplementedError("stand-in: FreeU is never enabled by the reference")


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.torch_utils.randn_tensor: draw on the generator's device, then move."""
    gdev = generator.device if generator is not None else (device or torch.device("cpu"))
    return torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device or gdev)
