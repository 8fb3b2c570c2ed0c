"""GPU (-m gpu), needs >= 2 GPUs on the box (skipped otherwise): the frame-sharded denoising loop over NCCL against
the single-GPU loop on identical inputs — tests/multigpu_check.py under torch.distributed.run on every visible
GPU count in {2, 4, 8}: CFG split + all-to-all exchange (default), plain frame sharding (MDK_CFG_SPLIT=0),
the K/V all-gather mode, even and uneven windows, CUDA graph == eager.  `gpurun --gpus N -- python -m pytest
tests/test_multigpu_nccl.py -m gpu` runs it; logs of passing runs at N = 2 / 4 / 8 are committed under profiles/."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("cfg_split", ["1", "0"])
def test_sharded_loop_equals_single_gpu_over_nccl(world, cfg_split):
    if NGPU < world:
        pytest.skip(f"needs {world} GPUs, {NGPU} visible")
    env = dict(os.environ, MDK_CFG_SPLIT=cfg_split)
    port = 29600 + world + (0 if cfg_split == "1" else 20)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", f"--master-port={port}",
                          os.path.join(ROOT, "tests", "multigpu_check.py")],
                         capture_output=True, text=True, env=env, timeout=600)
    tail = "\n".join(out.stdout.strip().splitlines()[-12:])
    assert out.returncode == 0 and "MULTIGPU PASS" in out.stdout, tail + "\n" + out.stderr[-1500:]
