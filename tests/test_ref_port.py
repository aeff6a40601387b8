"""oracle/ref_port.py — the CPU port that bench.py times as `cpu_baseline` / `--impl reference` — against the
golden vectors generated from the real reference (tests/golden/make_golden*.py).  CPU only."""
import numpy as np
import scipy.sparse as sp
import torch

from conftest import golden


def _coo(g, prefix="in_"):
    n = int(g["n"])
    idx = torch.from_numpy(np.stack([g[prefix + "row"], g[prefix + "col"]]).astype(np.int64))
    return n, torch.sparse_coo_tensor(idx, torch.from_numpy(g[prefix + "val"].astype(np.float32)), (n, n))


def test_build_adjacency_matches_reference():
    from oracle import ref_port as rp
    g = golden("csr_build.npz")
    B = rp.build_adjacency(g["row"], g["col"], int(g["n"]))              # utils_graphsaint.py:20-22 path
    B.sort_indices()
    assert np.array_equal(B.indptr, g["b_indptr"]) and np.array_equal(B.indices, g["b_indices"])
    assert np.array_equal(B.data.astype(np.float32), g["b_data"])


def test_normalize_and_propagate_match_reference():
    from oracle import ref_port as rp
    g = golden("normalize_plain.npz")
    n, adj = _coo(g)
    out = rp.normalize_adj_tensor_sparse(adj).coalesce()
    assert np.array_equal(out._indices().numpy(), g["out_idx"])
    assert np.array_equal(out._values().numpy(), g["out_val"])           # same scipy arithmetic: bit-identical
    p = golden("propagate.npz")
    n, adj = _coo(p)
    adj_norm = rp.normalize_adj_tensor_sparse(adj)                        # the golden loop runs on the normalised graph
    prop, target = rp.propagate(adj_norm, torch.from_numpy(p["x"]), int(p["T"]), float(p["alpha"]))
    assert np.array_equal(prop.numpy(), p["prop"]) and np.array_equal(target.numpy(), p["target"])


def test_kmeans_fit_matches_reference():
    from oracle import ref_port as rp
    g = golden("kmeans.npz")
    km = rp.kmeans_fit(g["x"], g["c0"], 300, tol=1e-4)
    assert km.n_iter_ == int(g["fit_n_iter"]) and np.array_equal(km.labels_, g["fit_labels"])
    np.testing.assert_allclose(km.cluster_centers_, g["fit_centers"], rtol=1e-6, atol=1e-6)
    assert abs(km.inertia_ - float(g["fit_inertia"])) <= 1e-6 * float(g["fit_inertia"])


def test_cluster_means_and_graph_compress_match_reference():
    from oracle import ref_port as rp
    g = golden("graph_compress.npz")
    n, adj = _coo(g)
    adj_norm = rp.normalize_adj_tensor_sparse(adj)
    labels = g["labels"]
    means = rp.cluster_means(torch.from_numpy(g["feat"]), labels, 20).numpy()
    fin = np.isfinite(g["means"])
    np.testing.assert_allclose(means[fin], g["means"][fin], rtol=1e-6, atol=1e-7)
    S = rp.graph_compress_sparse(labels.astype(np.int64), adj_norm).toarray()
    ref = g["syn_dense"]
    ok = np.isfinite(ref)                                               # the empty cluster gives NaN rows in the reference
    np.testing.assert_allclose(S[: ref.shape[0], : ref.shape[1]][ok], ref[ok], rtol=1e-5, atol=1e-8)


def test_sparsifier_port_matches_reference():
    from oracle import ref_port as rp
    g = golden("sparsify.npz")
    n = int(g["n"])
    idx = torch.from_numpy(np.stack([g["src"], g["dst"]]).astype(np.int64))
    adj = torch.sparse_coo_tensor(idx, torch.from_numpy(g["val"]), (n, n))
    src, dst = idx[0], idx[1]
    assert np.array_equal(rp.er_estimator(adj, src, dst).numpy(), g["er"])
    er_att, rew = rp.attaw_er_estimator(adj, torch.from_numpy(g["ebd"]), src, dst)
    assert np.array_equal(er_att.numpy(), g["er_att"]) and np.array_equal(rew.coalesce()._values().numpy(), g["rew_val"])
    out = rp.graph_sparse_attaw(adj, float(g["ratio"]), torch.from_numpy(g["ebd"]))
    assert len(out) == int(g["C"])
    for i, t in enumerate(out):
        ii = t.coalesce()._indices().numpy()
        assert np.array_equal(ii[0], g[f"att{i}_row"]) and np.array_equal(ii[1], g[f"att{i}_col"])


def test_build_condensed_bipartite_matches_reference():
    from oracle import ref_port as rp
    g = golden("recsys_ali_subset.npz")
    C = rp.build_condensed_bipartite(g["u"], g["i"], g["u2cu"], g["i2ci"], int(g["c_indptr"].shape[0] - 1),
                                     int(g["c_indices"].max()) + 1)
    C.sort_indices()
    assert np.array_equal(C.indptr, g["c_indptr"]) and np.array_equal(C.indices, g["c_indices"])
    assert np.array_equal(C.data.astype(np.float32), g["c_data"].astype(np.float32))
