"""CPU: the reference arm of bench.py keeps the JSON contract (tiny configuration so that it runs in seconds), and the
distinct-offset formulation of the continuous position bias used by the CUDA path equals the reference's all-pairs MLP."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", "tiny", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "volumes/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["steps"] == 2
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    """under torchrun only rank 0 runs and prints the reference arm; the other ranks exit 0 without work"""
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", "tiny", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, cwd=str(ROOT), env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_cpb_distinct_offset_table_equals_all_pairs_mlp():
    """attention.py:257-276 evaluates the MLP on all (h*w)^2 relative positions; the CUDA path evaluates the
    (2h-1)(2w-1) distinct offsets and gathers with `pair_index`. Same numbers (CPU, fp32, plain torch)."""
    from ctpa_clip_b200.ct_clip.attention import ContinuousPositionBias, pair_index
    torch.manual_seed(3)
    h, w = 4, 6
    cpb = ContinuousPositionBias(dim=32, heads=4)
    pos = torch.stack(torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")).reshape(2, -1).t()
    rel = (pos[:, None, :] - pos[None, :, :]).float()
    x = torch.sign(rel) * torch.log(rel.abs() + 1)
    for layer in cpb.net:
        x = layer(x)
    ref = x.permute(2, 0, 1)                                                   # (heads, n, n)
    t = cpb.rel_offsets(h, w, "cpu")                                           # (R, 2) signed-log offsets, R = (2h-1)(2w-1)
    for layer in cpb.net:
        t = layer(t)
    tab = t.t()                                                                # (heads, R)
    assert tab.shape == (4, (2 * h - 1) * (2 * w - 1))
    assert torch.allclose(tab[:, pair_index(h, w, "cpu")], ref, atol=1e-6)
