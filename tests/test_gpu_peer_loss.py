"""GPU (-m gpu): the fused latent exchange + global-batch InfoNCE over peer memory (`ctclip_clip_loss_allgather`,
SURVEY §8(e)) against the single-device `ctclip_clip_loss` on the concatenated batch — loss and latent gradients bit-exact
(same fp32 summation order; the temperature share is an atomic sum, compared to 1e-5), several steps in a row so that both buffer parities and the step-valued flags are exercised."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
WORKER = Path(__file__).resolve().parent / "_peer_loss_worker.py"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,b,d,steps", [(2, 8, 512, 5), (4, 3, 64, 5), (2, 2, 0, 0)])
def test_peer_memory_loss_matches_single_device(world, b, d, steps):
    """steps > 0: kernel level, 5 consecutive steps bit-exact; steps == 0: through CTCLIP.forward(return_loss=True) on the
    tiny configuration, data-parallel loss == single-process loss on the concatenated batch (+ backward runs)."""
    assert torch.cuda.is_available()
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, str(WORKER), str(r), str(world), str(port), str(b), str(d), str(steps)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=420)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "PEER_LOSS_OK" in out, f"rank {r} failed:\n{out[-3000:]}"
