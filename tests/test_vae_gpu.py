"""GPU (-m gpu): the native VAE against the fp32 restatement of diffusers' AutoencoderKL
(oracle/vae_oracle.py, pinned to the LDM encoder / decoder implementations inside the installed transformers by
tests/test_vae_oracle_pin.py).

STATUS: validated on a B200 in round 2 (profiles/r02_first_call.log); collected by the default -m gpu run.  The host orchestration is covered on CPU by tests/test_vae_oracle.py."""
import os

import pytest

pytestmark = pytest.mark.gpu

import gpu_diag as D  # noqa: E402


def test_vae_tiny_encode_decode_and_own_kernels():
    assert D.check_vae()


def test_vae_sd_size():
    assert D.check_vae_sd()
