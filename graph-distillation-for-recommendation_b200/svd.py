"""Truncated SVD embeddings of the interaction matrix.

``compute_svd_embeddings(R, dim, seed)`` of distill_recsys.py:124-155 — the step right before the k-means
stage of distill_recsys — returns ``(U sqrt(S), V sqrt(S))`` of the ``dim`` leading singular triplets.  The
reference calls scipy's ``svds`` (ARPACK, fp64, host).  Here the factorisation is a **block Krylov
Rayleigh-Ritz** whose sparse products R·Q and Rᵀ·Q run on the stage-2 CSR SpMM kernel (R and Rᵀ as device CSR),
i.e. the hot operation is the same HBM-bound gather kernel as the propagation; the small dense steps
(orthogonalisation of the N x b blocks, the (q·b)² Ritz problem) use torch's fp64 / fp32 linear algebra.

Parity with the reference is defined on what it determines uniquely: the singular values (relative 1e-4) and
the singular vectors of well separated values up to sign; the tail of a slowly decaying spectrum is only
determined as a subspace (the downstream k-means is invariant to rotations inside it only approximately —
the reference itself depends on ARPACK's starting vector there).
"""
from __future__ import annotations

from typing import Tuple, Union

import numpy as np
import scipy.sparse as sp
import torch

from ._dev import device_of
from .graph import CSR
from .propagation import spmm


def _orth_against(Y: torch.Tensor, basis: list) -> torch.Tensor:
    """Two passes of block Gram-Schmidt against the previous blocks, then a thin QR (fp64: the Krylov basis loses
    orthogonality quickly in fp32)."""
    for _ in range(2):
        for Q in basis:
            Y = Y - Q @ (Q.T @ Y)
    Qn, _ = torch.linalg.qr(Y)
    return Qn


def truncated_svd(R: CSR, k: int, seed: int = 42, block: int = None, n_blocks: int = None
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(U [n_rows, k], S [k] descending, V [n_cols, k]) of the k largest singular triplets of the device CSR ``R``."""
    n_rows, n_cols = R.shape
    dev = R.device
    Rt, _ = R.transpose()
    b = int(block or (k + 8))
    q = int(n_blocks or max(6, min(14, 600 // b)))
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    G = torch.randn((n_cols, b), generator=gen, dtype=torch.float32).to(dev)
    # Krylov blocks of R^T R on the column side:  V_0 = orth(G), V_{j+1} = orth(R^T (R V_j)) against all previous
    basis = []
    V = _orth_against(G.double(), basis)
    basis.append(V)
    for _ in range(1, q):
        Y = spmm(R, V.to(torch.float32).contiguous())                   # [n_rows, b]  sparse x dense on the CSR kernel
        Z = spmm(Rt, Y.contiguous())                                    # [n_cols, b]
        V = _orth_against(Z.double(), basis)
        basis.append(V)
    Q = torch.cat(basis, dim=1)                                         # [n_cols, q b], orthonormal
    B = spmm(R, Q.to(torch.float32).contiguous()).double()              # [n_rows, q b] = R Q
    # Rayleigh-Ritz: SVD of B through its Gram matrix would square the condition number; QR + small SVD instead
    Qb, Rb = torch.linalg.qr(B)
    Us, S, Vh = torch.linalg.svd(Rb)
    U = (Qb @ Us[:, :k]).to(torch.float32)
    Vk = (Q @ Vh.T[:, :k]).to(torch.float32)
    return U, S[:k].to(torch.float32), Vk


def compute_svd_embeddings(R: Union[sp.spmatrix, CSR], dim: int, seed: int = 42, device=None
                           ) -> Tuple[np.ndarray, np.ndarray]:
    """distill_recsys.py:124-155: ``(user_emb, item_emb) = (U sqrt(S), V sqrt(S))`` as float32 numpy arrays, singular
    values in descending order; raises the reference's ValueError when no factorisation is possible."""
    shape = R.shape
    k = min(int(dim), min(shape) - 1)
    if k <= 0:
        raise ValueError(f"Cannot compute SVD with shape={tuple(shape)} and dim={dim}")
    A = R if isinstance(R, CSR) else CSR.from_scipy(sp.csr_matrix(R, dtype=np.float32), device=device_of(device))
    U, S, V = truncated_svd(A, k, seed=seed)
    sqrt_s = torch.sqrt(torch.clamp_min(S, 1e-12)).reshape(1, -1)
    return (U * sqrt_s).cpu().numpy().astype(np.float32), (V * sqrt_s).cpu().numpy().astype(np.float32)
