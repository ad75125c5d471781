"""Host-side helpers of the hypercomplex stack: multiplication rules, initialisers, Kronecker
utilities (kept for API compatibility and tests — the kernels never build the Kronecker matrix)."""
from __future__ import annotations

import math
from typing import List

import numpy as np
import torch

# ---------------------------------------------------------------------------- multiplication rules


def _signed_unit_matrices(table: List[List[int]]) -> List[torch.Tensor]:
    """table[i][r] = +-(c+1): matrix i has sign(table)*1 at (r, |table|-1)."""
    n = len(table)
    mats = []
    for i in range(n):
        m = torch.zeros(n, n, dtype=torch.float32)
        for r, v in enumerate(table[i]):
            m[r, abs(v) - 1] = 1.0 if v > 0 else -1.0
        mats.append(m)
    return mats


# real left-multiplication matrices of 1,i,j,k acting on (r,i,j,k) columns (reference utils.py:5-22)
_QUATERNION = [[1, 2, 3, 4], [-2, 1, -4, 3], [-3, 4, 1, -2], [-4, -3, 2, 1]]
_COMPLEX = [[1, 2], [-2, 1]]                                   # reference utils.py:30-32


def get_multiplication_matrices(phm_dim: int, type: str = "standard") -> List[torch.Tensor]:
    """n contribution matrices A_i of shape [n,n] (reference phc/hypercomplex/utils.py:61-85).

    "standard": complex rule for n=2, Hamilton (quaternion) rule for n=4, otherwise the signed
    cyclic shifts  A_0 = I,  A_i = diag(+1,-1,+1,...) @ P^i  with P the right-shift permutation.
    "random": entries U(-1, 1).
    """
    assert type in ["standard", "random"]
    if type == "random":
        return [a for a in torch.empty(phm_dim, phm_dim, phm_dim, dtype=torch.float32).uniform_(-1, 1)]
    assert phm_dim >= 1
    if phm_dim == 2:
        return _signed_unit_matrices(_COMPLEX)
    if phm_dim == 4:
        return _signed_unit_matrices(_QUATERNION)
    shift = torch.roll(torch.eye(phm_dim, dtype=torch.float32), shifts=1, dims=1)
    signs = torch.diag(torch.tensor([1.0 if r % 2 == 0 else -1.0 for r in range(phm_dim)]))
    mats = [torch.eye(phm_dim, dtype=torch.float32)]
    for i in range(1, phm_dim):
        mats.append(signs @ torch.linalg.matrix_power(shift, i))
    return mats


# ---------------------------------------------------------------------------- initialisers
def unitary_init(phm_dim: int, in_features: int, out_features: int, low: float = 0, high: float = 1) -> torch.Tensor:
    """[n, in, out]: zero real part, U(low, high) imaginary parts, normalised over the component axis."""
    v = torch.zeros(phm_dim, in_features, out_features, dtype=torch.float32)
    for i in range(1, phm_dim):
        v[i].uniform_(low, high)
    return v / v.norm(p=2, dim=0)


def phm_init(phm_dim: int, in_features: int, out_features: int, low: float = 0, high: float = 1,
             criterion: str = "glorot", transpose: bool = True) -> torch.Tensor:
    """Hypercomplex polar initialisation (reference phc/hypercomplex/inits.py:16-44): chi-distributed
    modulus (df = n), uniform unit imaginary direction, uniform phase.  RNG streams are consumed in the
    reference's order (scipy/numpy chi -> torch uniform -> numpy uniform) so a seeded run draws the
    same weights."""
    from scipy.stats import chi
    if criterion == "glorot":
        s = math.sqrt(2.0 / (phm_dim * (in_features + out_features)))
    elif criterion == "he":
        s = math.sqrt(2.0 / (phm_dim * in_features))
    else:
        raise ValueError("Invalid criterion: " + criterion)
    shape = (in_features, out_features)
    modulus = torch.from_numpy(chi.rvs(phm_dim, loc=0, scale=s, size=shape)).to(torch.float32)
    direction = unitary_init(phm_dim, in_features, out_features, low, high)
    phase = torch.from_numpy(np.random.uniform(low=-np.pi, high=np.pi, size=shape)).to(torch.float32)
    w = torch.empty(phm_dim, in_features, out_features, dtype=torch.float32)
    w[0] = modulus * torch.cos(phase)
    w[1:] = modulus * direction[1:] * torch.sin(phase)
    return w.permute(0, 2, 1) if transpose else w


def glorot_uniform(t: torch.Tensor) -> torch.Tensor:
    return torch.nn.init.xavier_uniform_(t, gain=math.sqrt(2))


def glorot_normal(t: torch.Tensor) -> torch.Tensor:
    return torch.nn.init.xavier_normal_(t, gain=math.sqrt(2))


# ---------------------------------------------------------------------------- Kronecker utilities
def kronecker_product_einsum_batched(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """[b,a,c] x [b,k,p] -> [b, a*k, c*p]   (reference kronecker.py:35-48)."""
    assert A.dim() == 3 and B.dim() == 3
    b, a, c = A.shape
    _, k, p = B.shape
    return (A[:, :, None, :, None] * B[:, None, :, None, :]).reshape(b, a * k, c * p)


def kronecker_product(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Kronecker product over the last two axes with broadcast leading axes (reference kronecker.py:52-64)."""
    lead = torch.broadcast_shapes(a.shape[:-2], b.shape[:-2])
    res = a[..., :, None, :, None] * b[..., None, :, None, :]
    return res.reshape(lead + (a.shape[-2] * b.shape[-2], a.shape[-1] * b.shape[-1]))


def kronecker_product_single(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    assert A.dim() == B.dim() == 2
    return torch.kron(A, B)


def phm_cat(tensors: list, phm_dim: int, dim: int = -1) -> torch.Tensor:
    """Component-aware concat: the result's component c block is the concat of every input's
    component c block (reference utils.py:122-135)."""
    parts = [t.reshape(t.size(0), phm_dim, -1) for t in tensors]
    return torch.cat(parts, dim=-1).reshape(tensors[0].size(0), -1)
