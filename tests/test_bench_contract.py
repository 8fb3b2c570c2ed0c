"""CPU: bench.py's algorithmic-FLOP model against the figures of SURVEY.md §8d / BASELINE.md, its CLI contract
and the reference arm's JSON line (a bounded CPU sample; no GPU needed)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_per_image_flops_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    t96 = sum(bench.per_image_flops(96, 16).values()) / 1e12
    assert abs(t96 - 2.8524) < 1e-3                       # TFLOP per image per UNet forward, h = 96, f = 16
    assert abs(32 * t96 - 91.28) < 0.02                   # per step at config B (32 images)
    assert abs(sum(bench.per_image_flops(96, 32).values()) / 1e12 - 2.8559) < 1e-3
    assert abs(sum(bench.per_image_flops(128, 16).values()) / 1e12 - 5.9284) < 1e-3   # config E
    assert abs(sum(bench.per_image_flops(32, 4).values()) / 1e12 - 0.2554) < 1e-3     # config A
    parts = bench.per_image_flops(96, 16)
    tot = sum(parts.values())
    assert 0.32 < parts["conv"] / tot < 0.34 and 0.43 < parts["lin"] / tot < 0.45     # conv 33 %, linear 44 %
    assert set(bench.CONFIGS) == {"A", "B", "C", "D", "E"} and bench.CONFIGS["B"] == (96, 16, 20, 30)


@pytest.mark.timeout(600)
def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "A",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=580)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    # the reference's own modules where /root/reference is mounted (this container), the oracle port elsewhere
    want_kind = "reference" if os.path.isdir("/root/reference/src/models") else "port"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == want_kind and line["cpu_baseline"]["cores"] >= 1
    # what ran is what is reported: one frame (2 images) at the config's own latent size per timed step, and
    # ms_per_step is its measured time (the clip extrapolation lives under cpu_baseline.extrapolated only)
    assert line["config"]["latent"] == 32 and "latent 32x32" in line["cpu_baseline"]["sample"]
    assert abs(line["value"] - 1.0 / (line["config"]["num_inference_steps"] * line["ms_per_step"] / 1e3)) < 1e-9
    assert line["cpu_baseline"]["extrapolated"]["clip_s"] > 0
    assert line["e2e"] == dict(value=line["value"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert line["config"]["workload"].startswith("config A") and line["n_gpus"] == 1
    # under torchrun only rank 0 works: the other ranks exit 0 without output
    out2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                          capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=120)
    assert out2.returncode == 0 and out2.stdout.strip() == ""
