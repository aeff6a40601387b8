"""Stage 1 — adjacency build.  Host-side mirror of the reference's graph utilities.

Function names, argument meaning and return types follow
``/root/reference/ClustGDD/deep_robust_utils.py`` (to_tensor :85-113, normalize_adj
:180-207, normalize_adj_tensor :245-265, sparse_mx_to_torch_sparse_tensor :389-396,
to_scipy :408-417, is_sparse_tensor :419-436) and
``/root/reference/ClustGDD/distill_recsys.py`` (build_interaction_matrix :110-117),
but every arithmetic step runs in libgdr_b200 on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from ._dev import device_of, need_cuda, ptr, stream, to_device_f32, to_device_i64, workspace


@dataclass
class CSR:
    """Device-resident CSR matrix: rowptr int32[n_rows+1], colidx int32[nnz] (sorted inside a
    row), vals float32[nnz]."""

    rowptr: torch.Tensor
    colidx: torch.Tensor
    vals: torch.Tensor
    shape: Tuple[int, int]
    deg: Optional[torch.Tensor] = None  # fp64 degrees when produced by sym_normalize

    _plan: Optional[torch.Tensor] = None  # nnz-balanced row blocks of the SpMM (pattern-only, cached)

    @property
    def nnz(self) -> int:
        return int(self.colidx.shape[0])

    def spmm_plan(self) -> torch.Tensor:
        """int32 block bounds for gdr_spmm_prop_planned (built once per sparsity pattern)."""
        if self._plan is None:
            nb = _lib.query("gdr_spmm_plan_blocks", self.shape[0], self.nnz)
            bounds = torch.empty(nb + 1, dtype=torch.int32, device=self.device)
            _lib.call("gdr_spmm_plan", self.shape[0], self.nnz, ptr(self.rowptr), ptr(bounds), stream())
            self._plan = bounds
        return self._plan

    @property
    def device(self) -> torch.device:
        return self.rowptr.device

    # -- conversions ---------------------------------------------------------------
    def coo_indices(self) -> torch.Tensor:
        """int64 [2, nnz] row-major indices (the layout of deep_robust_utils.py:389-396)."""
        idx = torch.empty((2, self.nnz), dtype=torch.int64, device=self.device)
        if self.nnz:
            _lib.call("gdr_csr_to_coo", self.shape[0], ptr(self.rowptr), ptr(self.colidx),
                      ptr(idx[0]), ptr(idx[1]), stream())
        return idx

    def to_torch_coo(self) -> torch.Tensor:
        """torch sparse COO f32 tensor; carries this CSR as ``_gdr_csr`` so that later
        stages do not rebuild it."""
        t = torch.sparse_coo_tensor(self.coo_indices(), self.vals, self.shape, check_invariants=False)
        t._gdr_csr = self
        return t

    def to_scipy(self) -> sp.csr_matrix:
        return sp.csr_matrix(
            (self.vals.cpu().numpy(), self.colidx.cpu().numpy(), self.rowptr.cpu().numpy()),
            shape=self.shape,
        )

    def transpose(self) -> Tuple["CSR", torch.Tensor]:
        """Transposed CSR and the permutation t_perm (position of each transposed entry in
        this matrix)."""
        n_rows, n_cols = self.shape
        dev = self.device
        t_rowptr = torch.empty(n_cols + 1, dtype=torch.int32, device=dev)
        t_colidx = torch.empty(self.nnz, dtype=torch.int32, device=dev)
        t_perm = torch.empty(self.nnz, dtype=torch.int32, device=dev)
        wsb = _lib.query("gdr_csr_transpose_ws_bytes", n_rows, n_cols, self.nnz)
        ws = workspace(wsb, dev)
        _lib.call("gdr_csr_transpose", n_rows, n_cols, self.nnz, ptr(self.rowptr), ptr(self.colidx),
                  ptr(t_rowptr), ptr(t_colidx), ptr(t_perm), ptr(ws), ws.numel(), stream())
        t_vals = self.vals[t_perm.long()] if self.nnz else self.vals.clone()
        return CSR(t_rowptr, t_colidx, t_vals, (n_cols, n_rows)), t_perm

    @staticmethod
    def from_scipy(mx, device=None) -> "CSR":
        """Upload a scipy sparse matrix.  Goes through the device COO->CSR build, so
        duplicates are summed and columns sorted exactly as scipy's own tocsr() would."""
        dev = device_of(device)
        coo = mx.tocoo()
        return coo_to_csr(coo.row, coo.col, coo.data, coo.shape, device=dev)

    @staticmethod
    def from_torch_coo(t: torch.Tensor) -> "CSR":
        cached = getattr(t, "_gdr_csr", None)
        if cached is not None and cached.device == t.device:
            return cached
        need_cuda(t, "adj")
        idx, val = t._indices(), t._values()
        return coo_to_csr(idx[0], idx[1], val, tuple(t.shape), device=t.device)


def coo_to_csr(row, col, val, shape, *, symmetrize: bool = False, binarize: bool = False,
               device=None) -> CSR:
    """COO -> CSR with duplicate entries summed and sorted columns.

    Same result as ``sp.csr_matrix((val, (row, col)), shape)`` (utils.py:66-67,
    distill_recsys.py:116-117); ``symmetrize`` + ``binarize`` give
    ``adj + adj.T; adj[adj > 1] = 1`` (utils_graphsaint.py:20-22).
    Raises ValueError for indices outside ``shape`` like scipy does.
    """
    dev = device_of(device if device is not None else (row.device if isinstance(row, torch.Tensor) and row.is_cuda else None))
    n_rows, n_cols = int(shape[0]), int(shape[1])
    row_d = to_device_i64(row, dev)
    col_d = to_device_i64(col, dev)
    if row_d.shape != col_d.shape or row_d.dim() != 1:
        raise ValueError("row and col must be 1-D arrays of equal length")
    nnz_in = int(row_d.shape[0])
    val_d = None if val is None else to_device_f32(val, dev)
    if val_d is not None and val_d.shape[0] != nnz_in:
        raise ValueError("val must have the same length as row/col")
    cap = nnz_in * (2 if symmetrize else 1)
    rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    colidx = torch.empty(cap, dtype=torch.int32, device=dev)
    vals = torch.empty(cap, dtype=torch.float32, device=dev)
    meta = torch.zeros(2, dtype=torch.int64, device=dev)  # [nnz_out, status(int32 in low half)]
    wsb = _lib.query("gdr_coo_to_csr_ws_bytes", n_rows, n_cols, nnz_in, int(symmetrize))
    ws = workspace(wsb, dev)
    _lib.call("gdr_coo_to_csr", n_rows, n_cols, nnz_in, ptr(row_d), ptr(col_d), ptr(val_d),
              int(symmetrize), int(binarize), ptr(rowptr), ptr(colidx), ptr(vals),
              ptr(meta[0:1]), ptr(meta[1:2]), ptr(ws), ws.numel(), stream())
    nnz_out, status = (int(v) for v in meta.cpu().tolist())
    if status & 0xFFFFFFFF:
        raise ValueError("row/col index exceeds matrix dimensions")
    return CSR(rowptr, colidx[:nnz_out], vals[:nnz_out], (n_rows, n_cols))


def sym_normalize(A: CSR, self_loop_mode: int = 2) -> CSR:
    """D^-1/2 (A [+ I]) D^-1/2 on a device CSR (deep_robust_utils.py:180-207).
    ``self_loop_mode`` 2 reproduces the reference's ``if mx[0, 0] == 0: mx = mx + I``."""
    n = A.shape[0]
    if A.shape[0] != A.shape[1]:
        raise ValueError("sym_normalize needs a square matrix")
    dev = A.device
    cap = A.nnz + n
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    colidx = torch.empty(cap, dtype=torch.int32, device=dev)
    vals = torch.empty(cap, dtype=torch.float32, device=dev)
    deg = torch.empty(n, dtype=torch.float64, device=dev)
    nnz_out = torch.zeros(1, dtype=torch.int64, device=dev)
    wsb = _lib.query("gdr_sym_normalize_ws_bytes", n, A.nnz)
    ws = workspace(wsb, dev)
    _lib.call("gdr_sym_normalize", n, A.nnz, ptr(A.rowptr), ptr(A.colidx), ptr(A.vals),
              int(self_loop_mode), ptr(rowptr), ptr(colidx), ptr(vals), ptr(deg), ptr(nnz_out),
              ptr(ws), ws.numel(), stream())
    m = int(nnz_out.item())
    return CSR(rowptr, colidx[:m], vals[:m], A.shape, deg=deg)


# ---------------------------------------------------------------------------------
# reference-signature functions (deep_robust_utils.py)
# ---------------------------------------------------------------------------------
def is_sparse_tensor(tensor) -> bool:
    """deep_robust_utils.py:419-436."""
    return tensor.layout == torch.sparse_coo


def sparse_mx_to_torch_sparse_tensor(sparse_mx, device=None) -> torch.Tensor:
    """scipy sparse -> torch sparse COO f32 with int64 row-major indices
    (deep_robust_utils.py:389-396).  With ``device`` given the tensor is built on that
    GPU (the reference builds on the host and ``to_tensor`` moves it)."""
    coo = sparse_mx.tocoo().astype(np.float32)
    idx = torch.from_numpy(np.vstack((coo.row, coo.col)).astype(np.int64))
    val = torch.from_numpy(coo.data)
    if device is not None:
        idx, val = idx.to(device), val.to(device)
    return torch.sparse_coo_tensor(idx, val, torch.Size(coo.shape), check_invariants=False)


def to_tensor(adj, features, labels=None, device="cuda"):
    """deep_robust_utils.py:85-113: scipy / array inputs -> torch tensors on ``device``."""
    if sp.issparse(adj):
        adj = sparse_mx_to_torch_sparse_tensor(adj, device)
    else:
        adj = torch.as_tensor(np.asarray(adj), dtype=torch.float32).to(device)
    if sp.issparse(features):
        features = sparse_mx_to_torch_sparse_tensor(features, device)
    else:
        features = torch.as_tensor(np.array(features), dtype=torch.float32).to(device)
    if labels is None:
        return adj, features
    labels = torch.as_tensor(np.asarray(labels), dtype=torch.int64).to(device)
    return adj, features, labels


def to_scipy(tensor) -> sp.csr_matrix:
    """deep_robust_utils.py:408-417 (device -> host copy; duplicates are summed by the
    device COO->CSR build, as scipy's csr_matrix constructor would)."""
    if is_sparse_tensor(tensor):
        return CSR.from_torch_coo(tensor).to_scipy()
    idx = tensor.nonzero().t()
    vals = tensor[idx[0], idx[1]]
    return coo_to_csr(idx[0], idx[1], vals, tuple(tensor.shape), device=tensor.device).to_scipy()


def normalize_adj_tensor(adj: torch.Tensor, sparse: bool = False) -> torch.Tensor:
    """deep_robust_utils.py:245-265.  sparse=True: torch sparse COO in, torch sparse COO out
    (row-major, nnz(A)+N entries when the identity is added).  sparse=False: dense n x n."""
    need_cuda(adj, "adj")
    if sparse:
        A = CSR.from_torch_coo(adj)
        return sym_normalize(A, self_loop_mode=2).to_torch_coo()
    if adj.dim() != 2 or adj.shape[0] != adj.shape[1]:
        raise ValueError("dense adjacency must be square")
    a = adj.to(torch.float32).contiguous()
    n = a.shape[0]
    out = torch.empty_like(a)
    ws = workspace(_lib.query("gdr_sym_normalize_dense_ws_bytes", n), a.device)
    _lib.call("gdr_sym_normalize_dense", n, ptr(a), a.stride(0), ptr(out), out.stride(0), ptr(ws),
              ws.numel(), stream())
    return out


def normalize_adj(mx, device=None) -> sp.csr_matrix:
    """deep_robust_utils.py:180-207 on a scipy matrix: upload, normalise on the GPU, download.
    Returns CSR float32 (the reference's float64 result is cast to float32 by its only
    consumer, sparse_mx_to_torch_sparse_tensor)."""
    return sym_normalize(CSR.from_scipy(mx, device), self_loop_mode=2).to_scipy()


def build_interaction_matrix(num_users: int, num_items: int, u, i, values=None, device=None,
                             return_device: bool = False):
    """distill_recsys.py:110-117: user-item interaction CSR, duplicate pairs summed."""
    R = coo_to_csr(u, i, values, (int(num_users), int(num_items)), device=device)
    return R if return_device else R.to_scipy()
