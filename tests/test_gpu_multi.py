"""GPU (-m gpu), needs >= 2 devices (skipped on a 1-GPU box; run with `gpurun --gpus 2|4`): the data-parallel path on REAL
devices, one process per GPU, against the single-process reference on the concatenated batch — SURVEY §8(e)'s oracle.
Every rank checks the loss, EVERY parameter gradient after reduction and the EMA-updated codebook against
tests/golden/ctclip_mid4.pt (written by the UNMODIFIED reference, tools/make_golden.py mid4)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
WORKER = Path(__file__).resolve().parent / "_dp_worker.py"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, mode, env=None):
    port = _free_port()
    e = dict(os.environ)
    e.update(env or {})
    procs = [subprocess.Popen([sys.executable, str(WORKER), str(r), str(world), str(port), mode], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True, env=e) for r in range(world)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "DP_OK" in out, f"rank {r} failed:\n{out[-4000:]}"
    return outs


def _need(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (one per rank: spinning ranks never share a device)")


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_data_parallel_step_matches_single_process_reference(world, exchange):
    _need(world)
    outs = _run(world, "train", {"CTCLIP_LATENT_EXCHANGE": exchange})
    print("\n".join(o.strip().splitlines()[-1] for o in outs))


@pytest.mark.parametrize("world", [2, 4])
def test_peer_memory_loss_kernel_real_devices(world):
    _need(world)
    _run(world, "kernel")
