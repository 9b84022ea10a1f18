"""Symmetric (peer-mapped) latent buffers for the fused all-gather + logits kernel (csrc/symm.cu).

One `LatentExchange` per (local batch, latent dim, world): allocates this rank's buffer through the library
(cudaMalloc — CUDA IPC cannot export caching-allocator blocks), exchanges the 64-byte IPC handles with
`torch.distributed.all_gather_object` and maps every peer's buffer. `torch.distributed` is only the rendezvous; the
latents themselves travel through peer stores issued by the kernel.
"""
from __future__ import annotations

import ctypes as C
import os
import socket
import sys

import torch
import torch.distributed as dist

from . import _lib

_EXCHANGES: dict = {}


class LatentExchange:
    def __init__(self, b_local: int, d: int, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.b_local, self.d = b_local, d
        self.step = 0
        self.own = C.c_void_p(0)
        self.imported: list = []
        self.table = None
        lib = _lib.lib()
        lib.ctclip_symm_latent_bytes.restype = C.c_size_t
        nbytes = lib.ctclip_symm_latent_bytes(b_local, d, self.world)
        if nbytes == 0:
            raise _lib.CtclipError(f"latent exchange: unsupported shape b={b_local} d={d} world={self.world}")
        err = None
        handle = (C.c_ubyte * 64)()
        if lib.ctclip_symm_alloc(C.c_size_t(nbytes), C.byref(self.own)) != 0 or lib.ctclip_symm_export(self.own, handle) != 0:
            err = f"rank {self.rank}: {_lib.last_error()}"      # still take part in the rendezvous below
        mine = {"handle": bytes(handle), "host": socket.gethostname(), "pid": os.getpid(), "shape": (b_local, d, self.world),
                "ok": err is None}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        ptrs = (C.c_void_p * self.world)()
        for r, peer in enumerate(everyone):
            if err is not None:
                break
            if not peer["ok"]:
                err = f"rank {r} could not allocate / export its buffer"
                break
            if r == self.rank:
                ptrs[r] = self.own.value
                continue
            if peer["host"] != mine["host"] or peer["shape"] != mine["shape"]:
                err = f"rank {r}: not on this node or different shape ({peer['host']}, {peer['shape']})"
                break
            p = C.c_void_p(0)
            src = (C.c_ubyte * 64).from_buffer_copy(peer["handle"])
            rc = lib.ctclip_symm_import(src, C.byref(p))
            if rc != 0:
                err = f"rank {r}: {_lib.last_error()}"
                break
            self.imported.append(p)
            ptrs[r] = p.value
        # every rank must take the same decision: the kernel waits for all peers
        ok = torch.tensor([0 if err else 1], device="cuda", dtype=torch.int32)
        if dist.get_backend(group) == "gloo":
            ok = ok.cpu()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.close()
            raise _lib.CtclipError(f"latent exchange: peer mapping failed on some rank ({err or 'another rank'})")
        self.table = ptrs
        # bit 0 is raised by the kernel when a peer did not arrive within the timeout (the loss of that step is NaN and the
        # optimiser kernel skips the update); CTClipTrainStep.raise_if_skipped reads it
        self.status = torch.zeros(1, device="cuda", dtype=torch.int32)

    def next_step(self) -> int:
        self.step += 1
        return self.step

    def close(self):
        lib = _lib.lib()
        for p in self.imported:
            lib.ctclip_symm_unimport(p)
        self.imported = []
        if self.own.value:
            lib.ctclip_symm_free(self.own)
            self.own = C.c_void_p(0)
        self.table = None


def mode() -> str:
    """CTCLIP_LATENT_EXCHANGE = "peer" (default: fused peer-memory kernel) | "nccl" (all_gather_into_tensor + loss kernel,
    kept as the cross-check of the fused path and for ranks that do not share a node)"""
    return os.environ.get("CTCLIP_LATENT_EXCHANGE", "peer").lower()


def get_exchange(b_local: int, d: int, group=None):
    """the cached exchange for this shape, or None when the NCCL path is selected / peer mapping is impossible"""
    if mode() != "peer":
        return None
    key = (b_local, d, dist.get_world_size(group), id(group))
    if key not in _EXCHANGES:
        try:
            _EXCHANGES[key] = LatentExchange(b_local, d, group)
        except _lib.CtclipError as e:  # collective decision (all ranks raise together): stay on the GPU, through NCCL
            print(f"[ctpa_clip_b200] {e}; using the NCCL all-gather for the latents", file=sys.stderr, flush=True)
            _EXCHANGES[key] = None
    return _EXCHANGES[key]


def set_timeout(seconds: float):
    """how long the exchange kernel waits for a peer (default 600 s, env CTCLIP_PEER_TIMEOUT_S)"""
    _lib.lib().ctclip_symm_set_timeout_ms.argtypes = [C.c_ulonglong]
    _lib.check(_lib.lib().ctclip_symm_set_timeout_ms(C.c_ulonglong(max(1, int(seconds * 1000)))), "ctclip_symm_set_timeout_ms")


def timed_out() -> bool:
    """True when any exchange of this process gave up waiting for a peer (synchronises the device)"""
    return any(ex is not None and int(ex.status.item()) != 0 for ex in _EXCHANGES.values())


def shutdown():
    for ex in _EXCHANGES.values():
        if ex is not None:
            ex.close()
    _EXCHANGES.clear()
