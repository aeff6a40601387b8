#!/usr/bin/env python
"""Generate golden input/output vectors by running the REFERENCE's own functions.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box); the resulting small ``.npz`` fixtures are committed next to this script and are what
``tests/`` reads.  Nothing here is copied from the reference: its modules are imported
and executed as-is (missing optional wheels are stubbed with MagicMock as SURVEY §8c
describes) together with the third-party libraries that own the arithmetic
(scipy sparsetools, torch CPU sparse, scikit-learn 1.9.0).

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import scipy.sparse as sp
import torch

REF = "/root/reference/ClustGDD"
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub_missing_wheels():
    names = ["torch_geometric", "torch_geometric.transforms", "torch_geometric.utils", "torch_geometric.data",
             "torch_geometric.datasets", "torch_geometric.nn", "torch_geometric.nn.inits",
             "torch_geometric.typing", "torch_geometric.nn.conv", "torch_geometric.loader", "ogb",
             "ogb.nodeproppred", "torch_sparse", "torch_scatter", "matplotlib", "matplotlib.colors",
             "matplotlib.pyplot", "deeprobust", "deeprobust.graph", "deeprobust.graph.data",
             "deeprobust.graph.utils", "networkx"]
    for n in names:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = MagicMock()
    tgd = sys.modules["torch_geometric.data"]
    if isinstance(tgd, MagicMock):
        tgd.InMemoryDataset = type("InMemoryDataset", (), {})
        tgd.Data = type("Data", (), {})


def sha16(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def main():
    sys.path.insert(0, REF)
    _stub_missing_wheels()
    import deep_robust_utils as dru
    import distill_recsys as dr
    from sklearn.cluster import KMeans
    from sklearn.preprocessing import StandardScaler

    rng = np.random.RandomState(20251018)

    # ---------------- stage 1: CSR construction ----------------
    n, e = 60, 500
    r = rng.randint(0, n, e).astype(np.int64)
    c = rng.randint(0, n, e).astype(np.int64)
    A = sp.csr_matrix((np.ones(e), (r, c)), shape=(n, n))  # utils.py:66-67 (duplicates summed)
    B = sp.csr_matrix((np.ones(e), (r, c)), shape=(n, n))
    B.data[:] = 1.0                                         # stored graph of ones, as adj_full.npz is
    B = B + B.T
    B[B > 1] = 1                                            # utils_graphsaint.py:20-22
    B = sp.csr_matrix(B)
    B.sort_indices()
    np.savez_compressed(os.path.join(OUT, "csr_build.npz"), n=n, row=r, col=c,
                        a_indptr=A.indptr, a_indices=A.indices, a_data=A.data.astype(np.float32),
                        b_indptr=B.indptr, b_indices=B.indices, b_data=B.data.astype(np.float32))

    # ---------------- stage 1/4: recsys builders on the shipped Ali-Display file ----------------
    ali = np.loadtxt("/root/reference/Rankformer/data/Ali-Display/train.txt", dtype=np.int64)
    u_all, i_all = ali[:, 0], ali[:, 1]
    nu, ni = int(u_all.max()) + 1, int(i_all.max()) + 1
    R = dr.build_interaction_matrix(nu, ni, u_all, i_all)
    Cc = dr.build_condensed_bipartite(u_all, i_all, np.arange(nu) % 100, np.arange(ni) % 64, 100, 64)
    kat = dict(ali_shape=np.array([nu, ni]), ali_nnz=R.nnz, ali_lines=u_all.shape[0],
               ali_R_sha=sha16(R.indptr, R.indices, R.data), ali_R_max=float(R.data.max()),
               ali_C_sha=sha16(Cc.indptr, Cc.indices, Cc.data), ali_C_nnz=Cc.nnz, ali_C_sum=float(Cc.data.sum()),
               ali_C_max=float(Cc.data.max()))
    print("Ali-Display KAT:", kat)
    # a committed SUBSET (first 20 000 lines; ids kept) so the GPU box can run the same functions
    us, is_ = u_all[:20000], i_all[:20000]
    nus, nis = int(us.max()) + 1, int(is_.max()) + 1
    Rs = dr.build_interaction_matrix(nus, nis, us, is_)
    u2cu = (np.arange(nus) * 7 + 3) % 211
    i2ci = (np.arange(nis) * 5 + 1) % 97
    Cs = dr.build_condensed_bipartite(us, is_, u2cu, i2ci, 211, 97)
    np.savez_compressed(os.path.join(OUT, "recsys_ali_subset.npz"), u=us.astype(np.int32), i=is_.astype(np.int32),
                        nu=nus, ni=nis, r_indptr=Rs.indptr, r_indices=Rs.indices, r_data=Rs.data,
                        u2cu=u2cu.astype(np.int32), i2ci=i2ci.astype(np.int32), c_indptr=Cs.indptr,
                        c_indices=Cs.indices, c_data=Cs.data, **{k: np.asarray(v) for k, v in kat.items()})

    # ---------------- stage 1: symmetric normalisation ----------------
    def rand_sym(n, m, seed, self_loop0=False, isolated=None, weighted=False):
        rs = np.random.RandomState(seed)
        rr, cc = rs.randint(0, n, m), rs.randint(0, n, m)
        keep = rr != cc
        rr, cc = rr[keep], cc[keep]
        if isolated is not None:
            keep = (rr != isolated) & (cc != isolated)
            rr, cc = rr[keep], cc[keep]
        M = sp.csr_matrix((np.ones(rr.shape[0]), (rr, cc)), shape=(n, n))
        M = M + M.T
        M[M > 1] = 1
        M = sp.csr_matrix(M)
        if weighted:
            M.data = (rs.rand(M.nnz) + 0.5)
            M = sp.csr_matrix((M + M.T) / 2)
        if self_loop0:
            M = sp.lil_matrix(M)
            M[0, 0] = 1.0
            M[5, 5] = 2.0
            M = sp.csr_matrix(M)
        M.sort_indices()
        return M

    norm_cases = {}
    for name, kw in dict(plain=dict(seed=1), loop0=dict(seed=2, self_loop0=True), isolated=dict(seed=3, isolated=7),
                         loop0_isolated=dict(seed=4, self_loop0=True, isolated=9),
                         weighted=dict(seed=5, weighted=True)).items():
        M = rand_sym(300, 1500, **kw)
        adj, _ = dru.to_tensor(M, np.zeros((300, 1), dtype=np.float32))
        out = dru.normalize_adj_tensor(adj, sparse=True)       # deep_robust_utils.py:245-256
        norm_cases[name] = out
        norm_cases_in = M.tocoo()
        np.savez_compressed(os.path.join(OUT, f"normalize_{name}.npz"), n=300,
                            in_row=norm_cases_in.row.astype(np.int64), in_col=norm_cases_in.col.astype(np.int64),
                            in_val=norm_cases_in.data.astype(np.float32),
                            out_idx=out._indices().numpy(), out_val=out._values().numpy())
    Dn = torch.from_numpy(rand_sym(40, 120, seed=11).toarray().astype(np.float32))
    np.savez_compressed(os.path.join(OUT, "normalize_dense.npz"), a=Dn.numpy(),
                        out=dru.normalize_adj_tensor(Dn).numpy())  # :257-264

    # ---------------- stage 2: propagation loop ----------------
    M = rand_sym(300, 1500, seed=21)
    X = rng.randn(300, 13).astype(np.float32)
    adj, feat = dru.to_tensor(M, X)
    adj_norm = dru.normalize_adj_tensor(adj, sparse=True)
    T, alpha = 4, 0.8
    for t in range(T):                                       # clustgdd_agent_transduct.py:59-65
        if t == 0:
            prop_feat = feat
            target_feat = (1 - alpha) * prop_feat
        else:
            prop_feat = alpha * adj_norm @ prop_feat
            target_feat = target_feat + (1 - alpha) * prop_feat
    mc = M.tocoo()
    np.savez_compressed(os.path.join(OUT, "propagate.npz"), n=300, in_row=mc.row.astype(np.int64),
                        in_col=mc.col.astype(np.int64), in_val=mc.data.astype(np.float32), x=X, T=T, alpha=alpha,
                        prop=prop_feat.numpy(), target=target_feat.numpy())

    # ---------------- stage 3: k-means (sklearn with fixed init) ----------------
    def blobs(N, D, k, seed):
        rs = np.random.RandomState(seed)
        cen = rs.randn(k, D) * 3
        return (cen[rs.randint(0, k, N)] + rs.randn(N, D)).astype(np.float32)

    Xk = blobs(3000, 7, 25, 31)
    perm = np.random.RandomState(31).permutation(3000)
    C0 = Xk[perm[:40]].copy()
    km1 = KMeans(n_clusters=40, init=C0, n_init=1, max_iter=1, tol=0, algorithm="lloyd").fit(Xk)
    km = KMeans(n_clusters=40, init=C0, n_init=1, max_iter=300, tol=1e-4, algorithm="lloyd").fit(Xk)
    km0 = KMeans(n_clusters=40, init=C0, n_init=1, max_iter=12, tol=0, algorithm="lloyd").fit(Xk)
    # duplicate initial centres -> empty clusters -> relocation path
    C0e = C0.copy()
    C0e[5] = C0e[4]
    C0e[6] = C0e[4]
    kme = KMeans(n_clusters=40, init=C0e, n_init=1, max_iter=1, tol=0, algorithm="lloyd").fit(Xk)
    np.savez_compressed(os.path.join(OUT, "kmeans.npz"), x=Xk, c0=C0, c0_empty=C0e,
                        it1_labels=km1.labels_, it1_centers=km1.cluster_centers_, it1_inertia=km1.inertia_,
                        fit_labels=km.labels_, fit_centers=km.cluster_centers_, fit_inertia=km.inertia_,
                        fit_n_iter=km.n_iter_, tol0_labels=km0.labels_, tol0_centers=km0.cluster_centers_,
                        tol0_inertia=km0.inertia_, tol0_n_iter=km0.n_iter_,
                        empty_labels=kme.labels_, empty_centers=kme.cluster_centers_, empty_inertia=kme.inertia_)
    Xe = (rng.randn(500, 16) * rng.rand(16) * 3 + rng.randn(16)).astype(np.float32)
    Xe[:, 3] = 2.5  # zero-variance column
    np.savez_compressed(os.path.join(OUT, "standard_scaler.npz"), x=Xe,
                        out=StandardScaler(with_mean=True, with_std=True).fit_transform(Xe))  # distill_recsys.py:172

    # ---------------- stage 3/4: cluster means + graph_compress (agent methods) ----------------
    import clustgdd_agent_transduct as agent
    M = rand_sym(300, 1500, seed=41)
    adj, feat = dru.to_tensor(M, rng.randn(300, 9).astype(np.float32))
    adj_norm = dru.normalize_adj_tensor(adj, sparse=True)
    labels = torch.from_numpy(np.random.RandomState(41).randint(0, 20, 300).astype(np.int32))
    labels[labels == 13] = 12  # an empty cluster in the middle
    glist, adj_syn = agent.ClustGDD.graph_compress(None, labels, adj_norm, [adj_norm])  # transduct :234-250
    lf = labels.float()
    means = torch.stack([feat[torch.where(lf == i)[0]].mean(dim=0) for i in range(20)], dim=0)  # :121-125
    mc = M.tocoo()
    np.savez_compressed(os.path.join(OUT, "graph_compress.npz"), n=300, in_row=mc.row.astype(np.int64),
                        in_col=mc.col.astype(np.int64), in_val=mc.data.astype(np.float32), labels=labels.numpy(),
                        feat=feat.numpy(), means=means.numpy(), syn_dense=adj_syn.to_dense().numpy(),
                        syn_idx=adj_syn._indices().numpy(), syn_val=adj_syn._values().numpy(),
                        list0_dense=glist[0].to_dense().numpy())

    # ---------------- stages 1-2 bipartite: LightGCNCondensed.propagate ----------------
    torch.manual_seed(7)
    Cs_small = dr.build_condensed_bipartite(us[:5000], is_[:5000], u2cu, i2ci, 211, 97)
    ei, ew = dr.condensed_csr_to_edge_index(Cs_small, torch.device("cpu"))
    model = dr.LightGCNCondensed(211, 97, 16, 2, ei, ew, torch.device("cpu"))
    with torch.no_grad():
        uo, io = model.propagate()                           # distill_recsys.py:319-353
        w = model.edge_weight()
    np.savez_compressed(os.path.join(OUT, "lightgcn.npz"), edge_index=ei.numpy(), w=w.numpy(),
                        u0=(model.user_emb.weight + model.user_delta).detach().numpy(),
                        i0=(model.item_emb.weight + model.item_delta).detach().numpy(), layers=2,
                        u_out=uo.numpy(), i_out=io.numpy())
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
