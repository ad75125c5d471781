"""Flat aliasing of many small parameter tensors.

The reference keeps per-component parameters as separate tensors (n BatchNorm1d modules per norm,
n x #cols embedding tables per encoder; state-dict keys SURVEY.md §8b).  The kernels want one
contiguous vector per role.  ``alias_flat`` lays a list of tensors back to back in one buffer and
re-points each tensor's ``.data`` at its slice, so both views of the memory stay valid: the
state-dict / optimizer see the individual tensors, the kernels see the flat vector.  The layout
is re-established lazily whenever something (``module.to()``, ``param.data = ...`` as the
reference's reset_parameters does, unpickling) broke the aliasing.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


def is_packed(flat: Optional[torch.Tensor], tensors: Sequence[torch.Tensor], quick: bool = False) -> bool:
    """quick: only the first and the last tensor are checked (hot path; a group is moved / re-initialised as a whole)."""
    if flat is None or len(tensors) == 0:
        return False
    t0 = tensors[0]
    base = flat.data_ptr()
    esz = flat.element_size()
    if quick:
        tl = tensors[-1]
        return t0.data_ptr() == base and tl.data_ptr() == base + (flat.numel() - tl.numel()) * esz
    if flat.device != t0.device or flat.dtype != t0.dtype:
        return False
    off = 0
    for t in tensors:
        if t.data_ptr() != base + off * esz:      # also false after .to(device) / .data re-pointing
            return False
        off += t.numel()
    return off == flat.numel()


def view_if_contiguous(tensors: Sequence[torch.Tensor]) -> Optional[torch.Tensor]:
    """If ``tensors`` already sit back to back inside one storage (e.g. because a larger flat buffer — the optimizer's —
    contains them in this order), return a flat view over exactly that memory; otherwise None."""
    t0 = tensors[0]
    base, esz, off = t0.data_ptr(), t0.element_size(), 0
    for t in tensors:
        if t.device != t0.device or t.dtype != t0.dtype or not t.is_contiguous() or t.data_ptr() != base + off * esz:
            return None
        off += t.numel()
    st = t0.untyped_storage()
    start = base - st.data_ptr()
    if start < 0 or start % esz or start + off * esz > st.nbytes():
        return None
    return torch.empty(0, dtype=t0.dtype, device=t0.device).set_(st, start // esz, (off,))


def alias_flat(flat: Optional[torch.Tensor], tensors: Sequence[torch.Tensor], quick: bool = False) -> torch.Tensor:
    """Return a 1-D tensor whose storage IS the concatenation of ``tensors`` (re-packing if needed)."""
    if is_packed(flat, tensors, quick):
        return flat
    v = view_if_contiguous(tensors)
    if v is not None:
        return v
    with torch.no_grad():
        t0 = tensors[0]
        total = sum(t.numel() for t in tensors)
        new = torch.empty(total, dtype=t0.dtype, device=t0.device)
        off = 0
        for t in tensors:
            k = t.numel()
            view = new[off:off + k].view(t.shape)
            view.copy_(t.detach().to(device=t0.device, dtype=t0.dtype))
            t.data = view
            off += k
    return new
