"""Loaders and on-disk formats either side of the distillation core.

Host-side readers / writers with the reference's file names, keys and return kinds, plus the one
compute step the loaders contain — the induced subgraph ``adj[np.ix_(idx, idx)]`` of the inductive
split — on the device:

  GraphSAINT layout     data/<name>/{adj_full.npz, feats.npy, role.json, class_map.json}
                        DataGraphSAINT.__init__            utils_graphsaint.py:15-57
  Rankformer text       <dir>/<name>/{train,valid,test}.txt  "user item" per line
                        _read_ui_txt / load_rankformer_dataset   distill_recsys.py:63-108
  distilled export      condensed_graph.npz (cu, ci, w, num_cu, num_ci) + u2cu.npy + i2ci.npy
                        distill_recsys.py:736-753  (written so that downstream reference code loads them)
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from ._dev import device_of, ptr, stream, to_device_i64, workspace
from .graph import CSR, coo_to_csr


# ---------------------------------------------------------------------------------
# induced subgraph on the device
# ---------------------------------------------------------------------------------
def induced_subgraph(A: CSR, idx) -> CSR:
    """``adj[np.ix_(idx, idx)]`` (utils_graphsaint.py:34-36, utils.py:127-129): row / column i of the result is
    node ``idx[i]``; entries keep their values; columns come out sorted."""
    dev = A.device
    idx_d = to_device_i64(idx, dev)
    m, n = int(idx_d.numel()), A.shape[0]
    if m and (int(idx_d.min()) < 0 or int(idx_d.max()) >= n):
        raise IndexError("index out of bounds")       # what scipy raises for adj[np.ix_(...)]
    cap = max(A.nnz, 1)
    row = torch.empty(cap, dtype=torch.int64, device=dev)
    col = torch.empty(cap, dtype=torch.int64, device=dev)
    val = torch.empty(cap, dtype=torch.float32, device=dev)
    nnz = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = workspace(_lib.query("gdr_induced_subgraph_ws_bytes", n, m), dev)
    _lib.call("gdr_induced_subgraph_coo", n, ptr(A.rowptr), ptr(A.colidx), ptr(A.vals), m, ptr(idx_d), ptr(row), ptr(col),
              ptr(val), ptr(nnz), ptr(ws), ws.numel(), stream())
    k = int(nnz.item())
    return coo_to_csr(row[:k], col[:k], val[:k], (m, m), device=dev)


# ---------------------------------------------------------------------------------
# GraphSAINT layout
# ---------------------------------------------------------------------------------
@dataclass
class GraphSaintData:
    """What DataGraphSAINT exposes (utils_graphsaint.py:15-57), adjacency matrices as device CSR."""
    nnodes: int
    adj_full: CSR
    adj_train: CSR
    adj_val: CSR
    adj_test: CSR
    feat: torch.Tensor
    feat_train: torch.Tensor
    feat_val: torch.Tensor
    feat_test: torch.Tensor
    labels: np.ndarray
    labels_train: np.ndarray
    labels_val: np.ndarray
    labels_test: np.ndarray
    idx_train: np.ndarray
    idx_val: np.ndarray
    idx_test: np.ndarray
    nclass: int


def process_labels(class_map: dict, nnodes: int) -> Tuple[np.ndarray, int]:
    """DataGraphSAINT.process_labels: class_map.json -> int labels (multi-label lists are kept as a 0/1 matrix)."""
    first = next(iter(class_map.values()))
    if isinstance(first, list):
        nclass = len(first)
        lab = np.zeros((nnodes, nclass))
        for k, v in class_map.items():
            lab[int(k)] = v
        return lab, nclass
    lab = np.zeros(nnodes, dtype=np.int32)
    for k, v in class_map.items():
        lab[int(k)] = v
    lab = lab - lab.min()          # the minimum over ALL vertices, labelled or not, as the reference takes it
    return lab, int(lab.max()) + 1


def load_graphsaint(dataset_dir: str, symmetrize: bool = False, label_rate: Optional[float] = None,
                    device=None) -> GraphSaintData:
    """DataGraphSAINT.__init__ (utils_graphsaint.py:15-57).  ``symmetrize`` is the reference's ogbn-arxiv branch
    (``adj + adj.T`` clipped to 1, :20-22).  Features are z-scored with the statistics of the TRAINING rows
    (:38-43); the induced train / val / test adjacency blocks (:34-36) are cut on the device."""
    from .kmeans import standard_scale  # noqa: F401  (same StandardScaler arithmetic as gdr_standard_scale)
    dev = device_of(device)
    adj = sp.load_npz(os.path.join(dataset_dir, "adj_full.npz")).tocoo()
    n = adj.shape[0]
    A = coo_to_csr(adj.row, adj.col, None if symmetrize else adj.data, (n, n), symmetrize=symmetrize, binarize=symmetrize,
                   device=dev)
    role = json.load(open(os.path.join(dataset_dir, "role.json")))
    idx_train, idx_val, idx_test = (np.asarray(role[k], dtype=np.int64) for k in ("tr", "va", "te"))
    if label_rate is not None and label_rate < 1:
        idx_train = idx_train[: int(label_rate * len(idx_train))]
    feat = torch.from_numpy(np.load(os.path.join(dataset_dir, "feats.npy")).astype(np.float32)).to(dev)
    it, iv, ie = (torch.from_numpy(i).to(dev) for i in (idx_train, idx_val, idx_test))
    # StandardScaler().fit(feat[idx_train]) then transform(feat): statistics from the library kernel, fp64
    ft = feat[it].contiguous()
    D = feat.shape[1]
    mean = torch.empty(D, dtype=torch.float64, device=dev)
    scale = torch.empty(D, dtype=torch.float64, device=dev)
    tmp = torch.empty_like(ft)
    ws = workspace(_lib.query("gdr_standard_scale_ws_bytes", ft.shape[0], D), dev)
    _lib.call("gdr_standard_scale", ft.shape[0], D, ptr(ft), ft.stride(0), ptr(tmp), tmp.stride(0), ptr(mean), ptr(scale),
              ptr(ws), ws.numel(), stream())
    feat = (feat - mean.to(torch.float32)) / scale.to(torch.float32)
    class_map = json.load(open(os.path.join(dataset_dir, "class_map.json")))
    labels, nclass = process_labels(class_map, n)
    return GraphSaintData(nnodes=n, adj_full=A, adj_train=induced_subgraph(A, it), adj_val=induced_subgraph(A, iv),
                          adj_test=induced_subgraph(A, ie), feat=feat, feat_train=feat[it], feat_val=feat[iv],
                          feat_test=feat[ie], labels=labels, labels_train=labels[idx_train], labels_val=labels[idx_val],
                          labels_test=labels[idx_test], idx_train=idx_train, idx_val=idx_val, idx_test=idx_test,
                          nclass=nclass)


# ---------------------------------------------------------------------------------
# Rankformer text format
# ---------------------------------------------------------------------------------
@dataclass
class RecDataset:
    """distill_recsys.py:44-60."""
    num_users: int
    num_items: int
    train_u: np.ndarray
    train_i: np.ndarray
    valid_u: np.ndarray
    valid_i: np.ndarray
    test_u: np.ndarray
    test_i: np.ndarray

    @property
    def num_edges_train(self) -> int:
        return int(self.train_u.shape[0])


def read_ui_txt(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """'user item' per line, space separated (distill_recsys.py:63-73)."""
    arr = np.loadtxt(path, dtype=np.int64)
    if arr.ndim == 1:
        arr = arr.reshape(1, 2)
    if arr.shape[1] < 2:
        raise ValueError(f"Bad interaction file format: {path} (need 2 columns: user item)")
    return arr[:, 0].astype(np.int64, copy=False), arr[:, 1].astype(np.int64, copy=False)


def load_rankformer_dataset(data_dir: str, dataset: str) -> RecDataset:
    """distill_recsys.py:76-108."""
    root = os.path.join(data_dir, dataset)
    paths = [os.path.join(root, f"{s}.txt") for s in ("train", "valid", "test")]
    for p in paths:
        if not os.path.exists(p):
            raise FileNotFoundError(f"Missing file: {p}")
    (tu, ti), (vu, vi), (eu, ei) = (read_ui_txt(p) for p in paths)
    num_users = int(max(tu.max(initial=0), vu.max(initial=0), eu.max(initial=0)) + 1)
    num_items = int(max(ti.max(initial=0), vi.max(initial=0), ei.max(initial=0)) + 1)
    return RecDataset(num_users, num_items, tu, ti, vu, vi, eu, ei)


# ---------------------------------------------------------------------------------
# distilled export
# ---------------------------------------------------------------------------------
def save_condensed(out_dir: str, edge_index, w, num_cu: int, num_ci: int, u2cu, i2ci) -> None:
    """The artefacts distill_recsys.py:736-753 writes: condensed_graph.npz + u2cu.npy + i2ci.npy."""
    os.makedirs(out_dir, exist_ok=True)
    ei = edge_index.detach().cpu().numpy() if isinstance(edge_index, torch.Tensor) else np.asarray(edge_index)
    wv = w.detach().cpu().numpy() if isinstance(w, torch.Tensor) else np.asarray(w)
    np.savez_compressed(os.path.join(out_dir, "condensed_graph.npz"), cu=ei[0], ci=ei[1], w=wv,
                        num_cu=np.int64(num_cu), num_ci=np.int64(num_ci))
    np.save(os.path.join(out_dir, "u2cu.npy"), np.asarray(u2cu).astype(np.int64))
    np.save(os.path.join(out_dir, "i2ci.npy"), np.asarray(i2ci).astype(np.int64))


def load_condensed(out_dir: str):
    """Inverse of save_condensed: (scipy CSR num_cu x num_ci of the stored weights, u2cu, i2ci)."""
    g = np.load(os.path.join(out_dir, "condensed_graph.npz"))
    C = sp.coo_matrix((g["w"], (g["cu"], g["ci"])), shape=(int(g["num_cu"]), int(g["num_ci"]))).tocsr()
    return C, np.load(os.path.join(out_dir, "u2cu.npy")), np.load(os.path.join(out_dir, "i2ci.npy"))
