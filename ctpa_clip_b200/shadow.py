"""bf16 operand mirror of the fp32 parameters.

The tensor-core GEMMs read bf16 weights. Without a trainer every derived-operand cache casts its weights itself; with
`trainer.ParamArena` the fused clip + Adam kernel writes the bf16 image of every parameter in the same pass that updates it
(ctclip_adam_step `bf16_shadow`), into a flat arena with the parameters' own offsets, and `bf16_of(param)` is a zero-cost
view of that arena — no per-step cast / cat kernels between the optimiser and the next forward.

A view is only handed out while the parameter is untouched by torch since the mirror was last synchronised (`_version`
check: in-place torch writes such as load_state_dict bump it; the Adam kernel, which keeps the mirror in step, does not).
"""
from __future__ import annotations

import weakref

import torch

_arenas: list = []   # weakrefs to objects with .flat (fp32), .bf16, .versions {data_ptr: _version}


def register(arena) -> None:
    _arenas[:] = [r for r in _arenas if r() is not None]
    _arenas.append(weakref.ref(arena))


def _mirror_view(p: torch.Tensor):
    if not p.is_contiguous() or p.dtype != torch.float32:
        return None
    ptr = p.data_ptr()
    for r in _arenas:
        a = r()
        if a is None or a.bf16 is None:
            continue
        base = a.flat.data_ptr()
        if base <= ptr < base + 4 * a.flat.numel():
            if a.versions.get(ptr) != p._version:
                return None                      # modified behind the mirror's back: caller casts
            off = (ptr - base) // 4
            return a.bf16[off: off + p.numel()].view(p.shape)
    return None


def bf16_of(p: torch.Tensor) -> torch.Tensor:
    """bf16 [same shape] image of an fp32 weight: a view of the optimiser's mirror when one is in sync, else a fresh cast"""
    v = _mirror_view(p)
    if v is not None:
        return v
    return p.detach().to(torch.bfloat16).contiguous()


def bf16_rows_of(ps) -> torch.Tensor:
    """bf16 image of torch.cat(ps, 0) for 2-D weights of equal width; a single view when the mirror holds them back to back"""
    views = [_mirror_view(p) for p in ps]
    if all(v is not None for v in views):
        ok = all(views[i + 1].data_ptr() == views[i].data_ptr() + 2 * views[i].numel() and
                 views[i + 1].untyped_storage().data_ptr() == views[0].untyped_storage().data_ptr()
                 for i in range(len(views) - 1))
        if ok:
            rows = sum(v.shape[0] for v in views)
            return views[0].as_strided((rows, views[0].shape[1]), (views[0].shape[1], 1))
    return torch.cat([p.detach() for p in ps], 0).to(torch.bfloat16).contiguous()


def f32_cat_of(ps) -> torch.Tensor:
    """fp32 torch.cat(ps, 0) of 1-D parameters; a view when they already sit back to back in one storage (the arena)"""
    ok = all(p.is_contiguous() and p.dtype == torch.float32 for p in ps) and all(
        ps[i + 1].data_ptr() == ps[i].data_ptr() + 4 * ps[i].numel() and
        ps[i + 1].untyped_storage().data_ptr() == ps[0].untyped_storage().data_ptr() for i in range(len(ps) - 1))
    if ok:
        return ps[0].detach().as_strided((sum(p.numel() for p in ps),), (1,))
    return torch.cat([p.detach() for p in ps], 0).float().contiguous()
