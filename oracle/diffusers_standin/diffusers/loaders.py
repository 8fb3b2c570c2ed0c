class UNet2DConditionLoadersMixin:
    """LoRA / attention-processor loading mixin of diffusers 0.24.0: no behaviour on the inference path."""
