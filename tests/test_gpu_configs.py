"""Full-pipeline parity at the BASELINE.json shapes that are parity-test cases, not bench lines:
config A (Cora-shaped, the reference's own CPU-runnable case) end to end against the oracle, and
config C (Yelp2018-shaped bipartite) through the distill_recsys call sites with size-independent
properties.  GPU only."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def np_(t):
    return t.detach().cpu().numpy()


def test_config_a_cora_shaped_end_to_end(oracle):
    import gdr
    from gdr import synth
    cfg = synth.CONFIGS["A"]
    n, f, k, hops, d = cfg["n"], cfg["f"], cfg["k"], cfg["hops"], cfg["d_logit"]
    u, v = synth.uniform_graph(n, cfg["pairs"], seed=1234)
    X = synth.features(n, f, seed=1334, kind="l1")
    # stage 1 through the reference-signature functions (scipy in, torch sparse COO out)
    import scipy.sparse as sp
    A_sp = sp.csr_matrix((np.ones(u.shape[0]), (u, v)), shape=(n, n))
    A_sp = A_sp + A_sp.T
    A_sp[A_sp > 1] = 1
    adj, feat = gdr.to_tensor(sp.csr_matrix(A_sp), X, device=DEV)
    adj_norm = gdr.normalize_adj_tensor(adj, sparse=True)
    rp, ci, va = oracle.coo_to_csr(u, v, None, (n, n), symmetrize=True, binarize=True)
    rpo, cio, vo, _ = oracle.sym_normalize(rp, ci, va, n)
    rows = np.repeat(np.arange(n), np.diff(rpo))
    assert np.array_equal(np_(adj_norm._indices()), np.stack([rows, cio]))
    assert np.array_equal(np_(adj_norm._values()), vo)
    # stage 2: prop_num = hops + 1
    prop, target = gdr.propagate(adj_norm, feat, hops + 1, 0.8)
    p_ref, t_ref = oracle.propagate(rpo, cio, vo, X, hops + 1, 0.8)
    np.testing.assert_allclose(np_(target), t_ref, rtol=1e-5, atol=1e-6 * np.abs(X).max())
    assert tuple(target.shape) == (n, f)
    # stage 3 on 7-d "logits" (random projection of the propagated features), K = 140
    W = np.random.RandomState(7).standard_normal((f, d)).astype(np.float32)
    logits = (t_ref @ W).astype(np.float32)
    C0 = synth.kmeans_init(logits, k, seed=1234)
    km = gdr.KMeans(n_clusters=k, init=C0, n_init=1, max_iter=50, tol=1e-4).fit(logits)
    ref = oracle.kmeans_fit(logits, C0, max_iter=50, tol=1e-4)
    assert km.n_iter_ == ref["n_iter"]
    ok, n_band, n_bad = oracle.labels_match(km.labels_, logits - ref["mean"], ref["centers_centered"])
    assert ok, n_bad
    assert abs(km.inertia_ - ref["inertia"]) <= 1e-4 * ref["inertia"]
    labels = torch.from_numpy(km.labels_).to(DEV)
    # cluster means on the 1433-wide propagated features + coarsened graph
    means = np_(gdr.cluster_means(target, labels, k))
    m_ref = oracle.cluster_means(t_ref, km.labels_, k)
    okm = ~np.isnan(m_ref)
    np.testing.assert_allclose(means[okm], m_ref[okm], rtol=1e-5, atol=1e-7)
    _, adj_syn = gdr.graph_compress(labels, adj_norm, [])
    S = oracle.graph_compress_dense(km.labels_.astype(np.int64), rpo, cio, vo, int(km.labels_.max()) + 1)
    fin = np.isfinite(S)
    np.testing.assert_allclose(np_(adj_syn.to_dense())[fin], S[fin], rtol=1e-5, atol=1e-9)


def test_config_c_yelp_shaped_recsys_path(oracle):
    import gdr
    from gdr import synth
    cfg = synth.BIPARTITE["C"]
    U, I, E, d, L = cfg["users"], cfg["items"], cfg["inter"], cfg["d"], cfg["layers"]
    u, i = synth.bipartite_interactions(U, I, E, seed=1236)
    R = gdr.build_interaction_matrix(U, I, u, i, device=DEV)                     # distill_recsys.py:558
    assert R.shape == (U, I) and R.data.sum() == E and R.nnz < E               # duplicates were summed
    assert (np.diff(R.indptr) >= 0).all() and R.has_sorted_indices
    rp, ci, va = oracle.coo_to_csr(u, i, None, (U, I))
    assert np.array_equal(R.indptr, rp) and np.array_equal(R.indices, ci) and np.array_equal(R.data, va)
    rs = np.random.RandomState(3)
    emb_u = (rs.standard_normal((U, d)) * 0.1).astype(np.float32)
    emb_i = (rs.standard_normal((I, d)) * 0.1).astype(np.float32)
    ncu, nci = int(np.ceil(U * 0.1)), int(np.ceil(I * 0.1))
    u2cu, cu_centers = gdr.kmeans_cluster(emb_u, ncu, seed=42, init="random")   # :569-583
    i2ci, ci_centers = gdr.kmeans_cluster(emb_i, nci, seed=42, init="random")
    assert cu_centers.shape == (ncu, d) and u2cu.max() < ncu and i2ci.max() < nci
    C = gdr.build_condensed_bipartite(u, i, u2cu, i2ci, ncu, nci, device=DEV)   # :587
    rpc, cic, cnt, _ = oracle.coarsen_counts(u, i, u2cu, i2ci, ncu, nci)
    assert np.array_equal(C.indptr, rpc) and np.array_equal(C.indices, cic) and np.array_equal(C.data, cnt.astype(np.float32))
    assert C.data.sum() == E                                                     # counts LINES
    ei, ew = gdr.condensed_csr_to_edge_index(C, DEV)
    g = gdr.BipartiteGraph(ei, ew, ncu, nci)
    u0 = torch.from_numpy((rs.standard_normal((ncu, d)) * 0.1).astype(np.float32)).to(DEV)
    i0 = torch.from_numpy((rs.standard_normal((nci, d)) * 0.1).astype(np.float32)).to(DEV)
    uo, io = gdr.lightgcn_propagate(g, u0, i0, L)                                # :319-353
    u_ref, i_ref = oracle.lightgcn_propagate(rpc, cic, cnt.astype(np.float32), np_(u0), np_(i0), L)
    np.testing.assert_allclose(np_(uo), u_ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np_(io), i_ref, rtol=1e-5, atol=1e-6)
    # teacher pooling (:623-636)
    pooled = gdr.segment_mean_pool(torch.from_numpy(emb_u).to(DEV), torch.from_numpy(u2cu).to(DEV), ncu)
    s_ref, c_ref = oracle.segment_sum(emb_u, u2cu.astype(np.int32), ncu)
    np.testing.assert_allclose(np_(pooled), s_ref / np.maximum(c_ref, 1)[:, None], rtol=1e-6, atol=1e-7)


def test_rankformer_gcn_forward():
    import gdr
    from gdr import synth
    U, I, E, d = 700, 500, 20000, 32
    u, i = synth.bipartite_interactions(U, I, E, seed=5)
    x = np.random.RandomState(1).standard_normal((U + I, d)).astype(np.float32)
    for alpha, beta in [(1.0, 0.0), (0.5, 0.5)]:
        g = gdr.RankformerGCNGraph(torch.from_numpy(u).to(DEV), torch.from_numpy(i).to(DEV), U, I, alpha, beta)
        out = np_(gdr.rankformer_gcn_forward(g, torch.from_numpy(x).to(DEV)))
        # Rankformer/code/rec.py:118-137 restated with numpy scatter-adds
        du = np.maximum(np.bincount(u, minlength=U), 1).astype(np.float32)
        di = np.maximum(np.bincount(i, minlength=I), 1).astype(np.float32)
        w1 = (1.0 / du[u] ** np.float32(alpha) / di[i] ** np.float32(beta)).astype(np.float32)
        w2 = (1.0 / du[u] ** np.float32(beta) / di[i] ** np.float32(alpha)).astype(np.float32)
        zu = np.zeros((U, d), np.float64)
        zi = np.zeros((I, d), np.float64)
        np.add.at(zu, u, x[U:][i].astype(np.float64) * w1[:, None])
        np.add.at(zi, i, x[:U][u].astype(np.float64) * w2[:, None])
        np.testing.assert_allclose(out, np.concatenate([zu, zi]), rtol=2e-5, atol=1e-6)


def test_minibatch_kmeans_follows_sklearn():
    """MiniBatchKMeans (clustgdd_agent_transduct.py:103, distill_recsys.py:174-176): sklearn's algorithm with sklearn's
    random stream — same subsets, same batches, same reassignment draws.  A near-tie in an E-step can part the two
    trajectories, so the contract is WCSS; on well separated data the fits coincide."""
    import gdr
    from gdr import synth
    from sklearn.cluster import MiniBatchKMeans as SkMB
    X = synth.clustered_features(30000, 40, 50, seed=2)
    for kw in (dict(n_clusters=100, random_state=0, batch_size=1000),
               dict(n_clusters=50, random_state=3, batch_size=2048, init="random", n_init=1),
               dict(n_clusters=64, random_state=1, batch_size=512, max_iter=5, max_no_improvement=None)):
        mb = gdr.MiniBatchKMeans(**kw).fit(X)
        sk = SkMB(**kw).fit(X)
        assert mb.labels_.shape == (30000,) and mb.cluster_centers_.shape == (kw["n_clusters"], 40)
        assert mb.labels_.dtype == np.int32 and mb.n_steps_ >= 1
        assert abs(mb.inertia_ - sk.inertia_) <= 0.05 * sk.inertia_, (kw, mb.inertia_, sk.inertia_, mb.n_steps_, sk.n_steps_)
        # the inertia reported is the WCSS of the returned labels against the returned centres
        d = ((X - mb.cluster_centers_[mb.labels_]) ** 2).sum()
        assert abs(d - mb.inertia_) <= 1e-4 * mb.inertia_
    # distill_recsys.kmeans_cluster switches to mini-batches above 20 000 rows (distill_recsys.py:173-176)
    labels, centers = gdr.kmeans_cluster(X, 100, seed=0, minibatch=True, batch_size=2048)
    assert labels.dtype == np.int64 and centers.shape == (100, 40) and len(np.unique(labels)) > 50
    with pytest.raises(ValueError):
        gdr.MiniBatchKMeans(n_clusters=10, reassignment_ratio=-1.0).fit(X)
