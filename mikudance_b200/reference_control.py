"""ReferenceAttentionControl — same constructor / update / clear contract as
src/models/mutual_mix_attention.py:19-378 of the reference, without monkey-patching: the reader's
"read" behaviour (K/V = LN(x) + bank for the cond half, plain self-attention for the CFG uncond half)
is what UNetEngine._spatial executes; this object only records the mode on the model and moves the
feature banks between writer and reader blocks in the reference's pairing order.
"""
from __future__ import annotations

import torch


def torch_dfs(model: torch.nn.Module):
    result = [model]
    for child in model.children():
        result += torch_dfs(child)
    return result


def _is_transformer_block(m) -> bool:
    # class-name based so that reference-style writer UNets (BasicTransformerBlock) pair as well
    return type(m).__name__ in ("BasicTransformerBlock", "TemporalBasicTransformerBlock") and \
        hasattr(m, "norm1")


class ReferenceAttentionControl:
    def __init__(self, unet, mode="write", do_classifier_free_guidance=False,
                 attention_auto_machine_weight=float("inf"), gn_auto_machine_weight=1.0,
                 style_fidelity=1.0, reference_attn=True, reference_adain=False,
                 fusion_blocks="midup", batch_size=1) -> None:
        assert mode in ["read", "write"]
        assert fusion_blocks in ["midup", "full"]
        self.unet = unet
        self.mode = mode
        self.reference_attn = reference_attn
        self.reference_adain = reference_adain
        self.fusion_blocks = fusion_blocks
        self.do_classifier_free_guidance = do_classifier_free_guidance
        if reference_attn:
            mods = self._blocks(unet)
            for i, module in enumerate(mods):
                if not hasattr(module, "bank") or module.bank is None:
                    module.bank = []
                module.bank = []                                   # :314
                module.attn_weight = float(i) / float(len(mods))   # :315
            if hasattr(unet, "_ref_control"):
                # `blocks`: the hooked transformer blocks — in write mode the reference UNet appends
                # norm1(hidden_states) to exactly these blocks' banks (mutual_mix_attention.py:139-140)
                unet._ref_control = dict(mode=mode, fusion_blocks=fusion_blocks, blocks=mods,
                                         do_classifier_free_guidance=do_classifier_free_guidance)
                if fusion_blocks != "full" and mode == "read":
                    # "midup": only mid/up blocks read banks; the others keep empty banks, which the
                    # engine already treats as plain self-attention (mutual_mix_attention.py:169-172)
                    pass

    def _blocks(self, unet):
        if self.fusion_blocks == "midup":
            mods = [m for m in (torch_dfs(unet.mid_block) + torch_dfs(unet.up_blocks))
                    if _is_transformer_block(m)]
        else:
            mods = [m for m in torch_dfs(unet) if _is_transformer_block(m)]
        return sorted(mods, key=lambda x: -x.norm1.normalized_shape[0])   # stable, :300-302

    def update(self, writer, dtype=torch.float16):
        """reader.bank <- writer.bank, cast to fp16 (mutual_mix_attention.py:317-354)."""
        if self.reference_attn:
            readers = self._blocks(self.unet)
            writers = writer._blocks(writer.unet)
            for r, w in zip(readers, writers):
                r.bank = [v.clone().to(dtype) for v in w.bank]

    def clear(self):
        if self.reference_attn:
            for r in self._blocks(self.unet):
                r.bank.clear()
