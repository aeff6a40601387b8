"""Oracle parity at the FULL sizes of BASELINE.json's configs (SURVEY §8c protocol), through the C ABI:

  config B (ogbn-arxiv-shaped, 169 343 nodes) — every stage against scipy / the oracle / scikit-learn;
  config E (ogbn-products-shaped, 2.45 M nodes, 126 M non-zeros, K = 10 000) — CSR `array_equal` with the
      scipy recipe of utils_graphsaint.py:18-22, normalised values bit-equal to the oracle, one hop against the
      oracle's CSR loop on sampled rows, ONE E-step against the exact fp64 argmin (1e-6 band) on 60 000 sampled rows
      and against sklearn's own labels on all rows, coarsened counts against scipy P^T A P;
  config C / D (Yelp2018 / Amazon-book shaped bipartite) — interaction CSR, LightGCN propagation, per-side
      k-means single step, condensed counts against the oracle / scipy.

The host side of these tests is the expensive part (scipy builds, sklearn E-steps): about two minutes in total on
the GPU box.  Needs a B200: python -m pytest tests -m gpu."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gdr():
    import gdr as g
    assert torch.cuda.is_available()
    return g


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def np_(t):
    return t.detach().cpu().numpy()


def scipy_sym_binary(u, v, n):
    """utils.py:66-67 + utils_graphsaint.py:20-22: csr of ones (dups summed), A + A^T, A[A > 1] = 1."""
    A = sp.csr_matrix((np.ones(u.shape[0], dtype=np.float32), (u, v)), shape=(n, n))
    A = A + A.T
    A.data[:] = 1.0                                   # == A[A > 1] = 1 (all stored values are >= 1)
    A.sort_indices()
    return A


def exact_band_check(labels, rows, Xc, Cc, band=1e-6):
    """labels[rows] against the exact fp64 argmin of |x - c|^2 (BLAS fp64 GEMM on the host), 1e-6 relative band:
    outside the band the label must be the argmin, inside either of the two nearest is accepted."""
    C64 = Cc.astype(np.float64)
    cn = (C64 * C64).sum(1)
    n_band = n_bad = 0
    for s in range(0, rows.shape[0], 4096):
        r = rows[s:s + 4096]
        x = Xc[r].astype(np.float64)
        d = (x * x).sum(1)[:, None] - 2.0 * (x @ C64.T) + cn[None, :]
        two = np.argpartition(d, 1, axis=1)[:, :2]
        dd = np.take_along_axis(d, two, axis=1)
        swap = dd[:, 1] < dd[:, 0]
        j1 = np.where(swap, two[:, 1], two[:, 0])
        j2 = np.where(swap, two[:, 0], two[:, 1])
        d1, d2 = dd.min(1), dd.max(1)
        # the GEMM form of the distance is itself only accurate to ~1e-13 |x||c|: re-evaluate both candidates exactly
        e1 = ((x - C64[j1]) ** 2).sum(1)
        e2 = ((x - C64[j2]) ** 2).sum(1)
        sw = e2 < e1
        j1, j2, e1, e2 = np.where(sw, j2, j1), np.where(sw, j1, j2), np.minimum(e1, e2), np.maximum(e1, e2)
        margin = (e2 - e1) / np.maximum(e2, 1e-30)
        inband = margin <= band
        lab = labels[r]
        good = (lab == j1) | (inband & (lab == j2))
        n_band += int(inband.sum())
        n_bad += int((~good).sum())
    return n_band, n_bad


def single_step_protocol(gdr, dev, oracle, X, C0, sample=None):
    """SURVEY 8c protocol 1 against scikit-learn's own E-step (sklearn/cluster/_kmeans.py:761 _labels_inertia, the routine
    behind KMeans.fit) on SHARED mean-centred inputs:
      labels     ours == sklearn's, except rows inside the 1e-6 relative margin band (either of the two nearest accepted,
                 checked in exact fp64); ours additionally against the exact fp64 argmin (all rows, or ``sample`` rows)
      M-step     per-cluster sums / counts on our labels bit-identical to the sequential fp32 oracle
      C_{t+1}    KMeans(max_iter=1) centres within 1e-5 of sklearn's for every cluster whose membership is identical,
                 WCSS within 1e-4
    Returns our first-step labels."""
    import os
    from sklearn.cluster import KMeans as SkKMeans
    from sklearn.cluster._kmeans import _labels_inertia_threadpool_limit
    from gdr._dev import padded_rows
    n, K = X.shape[0], C0.shape[0]
    mean = X.mean(axis=0)
    Xc, C0c = np.ascontiguousarray(X - mean), np.ascontiguousarray(C0 - mean)
    Xd, Cd = padded_rows(torch.from_numpy(Xc).to(dev)), padded_rows(torch.from_numpy(C0c).to(dev))
    lab_d = torch.empty(n, dtype=torch.int32, device=dev)
    gdr.assign_labels(Xd, Cd, lab_d, tc_operand=gdr.kmeans.TcOperand(Xd) if X.shape[1] <= 128 else None)
    lab = np_(lab_d)
    rows = np.arange(n) if sample is None else np.sort(np.random.RandomState(1).choice(n, sample, replace=False))
    _, n_bad = exact_band_check(lab, rows, Xc, C0c)
    assert n_bad == 0, f"{n_bad} rows outside the 1e-6 band disagree with the exact fp64 argmin"
    sk_lab = _labels_inertia_threadpool_limit(Xc, np.ones(n, dtype=np.float32), C0c, n_threads=len(os.sched_getaffinity(0)),
                                              return_inertia=False)
    diff = np.flatnonzero(lab != sk_lab)
    assert diff.size <= max(8, int(2e-4 * n)), f"{diff.size} rows differ from sklearn's E-step"
    if diff.size:
        _, bad_ours = exact_band_check(lab, diff, Xc, C0c)
        _, bad_sk = exact_band_check(sk_lab, diff, Xc, C0c)
        assert bad_ours == 0, f"{bad_ours} of {diff.size} differing rows: our label is outside the band"
        assert bad_sk == 0      # (sklearn's fp32 GEMM is itself only exact inside the band)
    sums, counts = gdr.segment_sum(Xd, lab_d, K)
    s_ref, c_ref = oracle.segment_sum(Xc, lab, K)
    assert np.array_equal(np_(counts), c_ref) and np.array_equal(np_(sums), s_ref)
    del Xd, lab_d, sums
    km1 = gdr.KMeans(n_clusters=K, init=C0, n_init=1, max_iter=1, tol=0).fit(X)
    sk1 = SkKMeans(n_clusters=K, init=C0, n_init=1, max_iter=1, tol=0, algorithm="lloyd").fit(X)
    same = np.ones(K, dtype=bool)
    same[lab[diff]] = False
    same[sk_lab[diff]] = False
    same &= c_ref > 0
    assert same.sum() >= K - 4 * max(1, diff.size) - (c_ref == 0).sum()
    np.testing.assert_allclose(km1.cluster_centers_[same], sk1.cluster_centers_[same], rtol=1e-5, atol=1e-5)
    assert abs(km1.inertia_ - sk1.inertia_) <= 1e-4 * sk1.inertia_
    assert (km1.labels_ != sk1.labels_).mean() <= 1e-3
    return lab


# ---------------------------------------------------------------------------------------------- config B
@pytest.fixture(scope="module")
def cfg_b(gdr, dev):
    from gdr import synth
    cfg = dict(synth.CONFIGS["B"])
    n = cfg["n"]
    u, v = synth.uniform_graph(n, cfg["pairs"], seed=1235)
    X = synth.features(n, cfg["f"], seed=1335)
    A = gdr.coo_to_csr(u, v, None, (n, n), symmetrize=True, binarize=True, device=dev)
    An = gdr.sym_normalize(A, 2)
    cfg.update(u=u, v=v, X=X, A=A, An=An)
    return cfg


def test_config_b_stage1_equals_scipy_and_oracle(gdr, dev, oracle, cfg_b):
    n, A, An = cfg_b["n"], cfg_b["A"], cfg_b["An"]
    S = scipy_sym_binary(cfg_b["u"], cfg_b["v"], n)
    assert np.array_equal(np_(A.rowptr), S.indptr) and np.array_equal(np_(A.colidx), S.indices)
    assert np.array_equal(np_(A.vals), S.data)
    rpo, cio, vo, deg = oracle.sym_normalize(S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data, n)
    assert np.array_equal(np_(An.rowptr), rpo) and np.array_equal(np_(An.colidx), cio)
    assert np.array_equal(np_(An.vals), vo)                       # fp64-then-round recipe: bit-exact
    assert np.array_equal(np_(An.deg), deg)


def test_config_b_stage2_equals_oracle(gdr, dev, oracle, cfg_b):
    An, X = cfg_b["An"], cfg_b["X"]
    prop, target = gdr.propagate(An, torch.from_numpy(X).to(dev), cfg_b["hops"] + 1, 0.8)
    p_ref, t_ref = oracle.propagate(np_(An.rowptr), np_(An.colidx), np_(An.vals), X, cfg_b["hops"] + 1, 0.8)
    scale = np.abs(X).max()
    np.testing.assert_allclose(np_(prop), p_ref, rtol=1e-5, atol=1e-6 * scale)
    np.testing.assert_allclose(np_(target), t_ref, rtol=1e-5, atol=1e-6 * scale)
    cfg_b["target"] = target


@pytest.mark.parametrize("space", ["feature", "logit"])
def test_config_b_stage3_equals_sklearn(gdr, dev, oracle, cfg_b, space):
    """Feature space (D = 128) and the reference-faithful logit space (D = 40, clustgdd_agent_transduct.py:88-105)."""
    from sklearn.cluster import KMeans as SkKMeans
    from gdr import synth
    n, K = cfg_b["n"], cfg_b["k"]
    if space == "feature":
        if "target" not in cfg_b:
            cfg_b["target"] = gdr.propagate(cfg_b["An"], torch.from_numpy(cfg_b["X"]).to(dev), cfg_b["hops"] + 1, 0.8)[1]
        X = np.ascontiguousarray(np_(cfg_b["target"]))
    else:
        X = synth.clustered_features(n, cfg_b["d_logit"], cfg_b["d_logit"], seed=77)
    C0 = synth.kmeans_init(X, K, seed=1235)
    single_step_protocol(gdr, dev, oracle, X, C0)
    # protocol 2: end to end, 20 iterations tol = 0 — same iteration count, WCSS within 1e-4
    km = gdr.KMeans(n_clusters=K, init=C0, n_init=1, max_iter=20, tol=0).fit(X)
    sk = SkKMeans(n_clusters=K, init=C0, n_init=1, max_iter=20, tol=0, algorithm="lloyd").fit(X)
    assert km.n_iter_ == sk.n_iter_
    # SURVEY 8c protocol 2: identical labels when no row ever entered the band, else WCSS within 1e-4 (a flipped in-band
    # row perturbs the rest of the trajectory: unclustered propagated features end ~98.6 % equal after 20 iterations)
    assert abs(km.inertia_ - sk.inertia_) <= 1e-4 * sk.inertia_
    assert (km.labels_ == sk.labels_).mean() > 0.95
    if space == "feature":
        cfg_b["labels"] = km.labels_


def test_config_b_stage4_equals_scipy(gdr, dev, oracle, cfg_b):
    n, K, An = cfg_b["n"], cfg_b["k"], cfg_b["An"]
    labels = cfg_b.get("labels")
    if labels is None:
        labels = np.random.RandomState(3).randint(0, K, n).astype(np.int32)
    lab64 = labels.astype(np.int64)
    k = int(lab64.max()) + 1
    # cluster means (clustgdd_agent_transduct.py:116-127) against the oracle
    if "target" in cfg_b:
        means = np_(gdr.cluster_means(cfg_b["target"], torch.from_numpy(labels).to(dev), k))
        ref = oracle.cluster_means(np_(cfg_b["target"]), lab64, k)
        ok = ~np.isnan(ref)
        assert np.array_equal(np.isnan(means), np.isnan(ref))
        np.testing.assert_allclose(means[ok], ref[ok], rtol=1e-5, atol=1e-6)
    # integer cell counts: P^T B P with B the pattern of A_hat and P the one-hot of the labels (scipy, exact)
    S = An.to_scipy()
    P = sp.csr_matrix((np.ones(n, dtype=np.int64), (np.arange(n), lab64)), shape=(n, k))
    Bp = sp.csr_matrix((np.ones(S.nnz, dtype=np.int64), S.indices, S.indptr), shape=S.shape)
    Cnt = (P.T @ Bp @ P).tocsr()
    Cnt.setdiag(0)
    Cnt.eliminate_zeros()
    Cnt.sort_indices()
    lab_d = torch.from_numpy(labels).to(dev)
    rpc, cic, cnt, _ = gdr.coarsen_edges(lab_d, lab_d, k, k, csr=An, drop_diag=True)
    assert np.array_equal(np_(rpc), Cnt.indptr) and np.array_equal(np_(cic), Cnt.indices)
    assert np.array_equal(np_(cnt), Cnt.data)
    # values S = P^T A_hat P / (n_a n_b), diagonal removed (clustgdd_agent_transduct.py:234-250), 1e-5
    sizes = np.bincount(lab64, minlength=k).astype(np.float64)
    Pn = sp.csr_matrix((1.0 / sizes[lab64], (np.arange(n), lab64)), shape=(n, k))
    Sv = (Pn.T @ S.astype(np.float64) @ Pn).tocsr()
    Sv.setdiag(0)
    Sv.eliminate_zeros()
    Sv.sort_indices()
    _, adj_syn = gdr.graph_compress(lab_d, An, [])
    got = sp.csr_matrix((np_(adj_syn._values()), (np_(adj_syn._indices()[0]), np_(adj_syn._indices()[1]))), shape=(k, k))
    got.sort_indices()
    assert np.array_equal(got.indptr, Sv.indptr) and np.array_equal(got.indices, Sv.indices)
    np.testing.assert_allclose(got.data, Sv.data, rtol=1e-5, atol=1e-12)


# ---------------------------------------------------------------------------------------------- config E
@pytest.fixture(scope="module")
def cfg_e(gdr, dev):
    from gdr import synth
    cfg = dict(synth.CONFIGS["E"])
    n = cfg["n"]
    u, v = synth.uniform_graph(n, cfg["pairs"], seed=1238)
    A = gdr.coo_to_csr(torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev), None, (n, n), symmetrize=True,
                       binarize=True)
    An = gdr.sym_normalize(A, 2)
    cfg.update(u=u, v=v, A=A, An=An)
    return cfg


def test_config_e_stage1_equals_scipy_and_oracle(gdr, dev, oracle, cfg_e):
    n, A, An = cfg_e["n"], cfg_e["A"], cfg_e["An"]
    S = scipy_sym_binary(cfg_e["u"], cfg_e["v"], n)
    assert np.array_equal(np_(A.rowptr), S.indptr) and np.array_equal(np_(A.colidx), S.indices)
    rpo, cio, vo, deg = oracle.sym_normalize(S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data, n)
    del S
    assert np.array_equal(np_(An.rowptr), rpo) and np.array_equal(np_(An.colidx), cio)
    assert np.array_equal(np_(An.vals), vo)
    assert np.array_equal(np_(An.deg), deg)
    cfg_e["csr_host"] = (rpo, cio, vo)


def test_config_e_one_hop_equals_oracle_on_sampled_rows(gdr, dev, oracle, cfg_e):
    from gdr import synth
    n, f, An = cfg_e["n"], cfg_e["f"], cfg_e["An"]
    X = synth.features(n, f, seed=1338)
    Xd = torch.from_numpy(X).to(dev)
    prop, target = gdr.propagate(An, Xd, 2, 0.8)                  # one hop + the t = 0 term
    rp, ci, va = cfg_e.get("csr_host") or (np_(An.rowptr), np_(An.colidx), np_(An.vals))
    rows = np.sort(np.random.RandomState(0).choice(n, 200_000, replace=False))
    cnt = (rp[rows + 1] - rp[rows]).astype(np.int64)
    sub_rp = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    take = np.repeat(rp[rows].astype(np.int64) - sub_rp[:-1], cnt) + np.arange(int(sub_rp[-1]))
    T0 = (np.float32(1.0 - 0.8) * X[rows]).astype(np.float32)
    y_ref = oracle.spmm_prop(sub_rp, ci[take], va[take], np.float32(0.8), X, T=T0, beta=np.float32(1.0 - 0.8))
    t_ref = T0                                                    # updated in place: T += beta * Y
    scale = np.abs(X).max()
    rows_d = torch.from_numpy(rows).to(dev)
    np.testing.assert_allclose(np_(prop[rows_d]), y_ref, rtol=1e-5, atol=1e-6 * scale)
    np.testing.assert_allclose(np_(target[rows_d]), t_ref, rtol=1e-5, atol=1e-6 * scale)
    cfg_e["target"] = target


def test_config_e_estep_equals_exact_argmin_and_sklearn(gdr, dev, oracle, cfg_e):
    """K = 10 000 (10 240 padded) centres x 2.45 M rows through the two-level tcgen05 screen: one E-step from shared
    centres against sklearn's own E-step on ALL rows and against the exact fp64 argmin on 60 000 sampled rows, the
    M-step against the sequential oracle, and KMeans(max_iter=1) against sklearn's fit (single_step_protocol)."""
    from gdr import synth
    n, K, f = cfg_e["n"], cfg_e["k"], cfg_e["f"]
    target = cfg_e.get("target")
    if target is None:
        target = gdr.propagate(cfg_e["An"], torch.from_numpy(synth.features(n, f, seed=1338)).to(dev), 2, 0.8)[1]
    X = np.ascontiguousarray(np_(target))
    C0 = synth.kmeans_init(X, K, seed=1238)
    cfg_e["labels"] = single_step_protocol(gdr, dev, oracle, X, C0, sample=60_000)


def test_config_e_stage4_counts_equal_scipy(gdr, dev, cfg_e):
    n, K, An = cfg_e["n"], cfg_e["k"], cfg_e["An"]
    labels = cfg_e.get("labels")
    if labels is None:
        labels = np.random.RandomState(5).randint(0, K, n).astype(np.int32)
    lab64 = labels.astype(np.int64)
    k = int(lab64.max()) + 1
    rp, ci, _ = cfg_e.get("csr_host") or (np_(An.rowptr), np_(An.colidx), None)
    # exact integer counts of (label[src], label[dst]) cells, diagonal dropped: numpy on packed keys
    src_lab = np.repeat(lab64, np.diff(rp.astype(np.int64)))
    key = src_lab * k + lab64[ci]
    key = key[src_lab != lab64[ci]]
    cells, cnt_ref = np.unique(key, return_counts=True)
    lab_d = torch.from_numpy(labels).to(dev)
    rpc, cic, cnt, _ = gdr.coarsen_edges(lab_d, lab_d, k, k, csr=An, drop_diag=True)
    crow = np.repeat(np.arange(k, dtype=np.int64), np.diff(np_(rpc).astype(np.int64)))
    assert np.array_equal(crow * k + np_(cic), cells)
    assert np.array_equal(np_(cnt), cnt_ref)


# ---------------------------------------------------------------------------------------------- configs C / D
@pytest.mark.parametrize("name,idx", [("C", 2), ("D", 3)])
def test_bipartite_config_equals_oracle(gdr, dev, oracle, name, idx):
    from sklearn.cluster import KMeans as SkKMeans
    from sklearn.preprocessing import StandardScaler
    from gdr import synth
    cfg = synth.BIPARTITE[name]
    nu, ni, d, seed = cfg["users"], cfg["items"], cfg["d"], 1234 + idx
    u, i = synth.bipartite_interactions(nu, ni, cfg["inter"], seed)
    # a1: interaction CSR with duplicate lines summed (distill_recsys.py:110-117)
    R = gdr.build_interaction_matrix(nu, ni, u, i, device=dev, return_device=True)
    S = sp.csr_matrix((np.ones(u.shape[0], dtype=np.float32), (u, i)), shape=(nu, ni))
    S.sum_duplicates()
    S.sort_indices()
    assert np.array_equal(np_(R.rowptr), S.indptr) and np.array_equal(np_(R.colidx), S.indices)
    assert np.array_equal(np_(R.vals), S.data)
    # a6: LightGCN propagation over the whole interaction graph (distill_recsys.py:319-353)
    rs = np.random.RandomState(seed)
    u0 = (0.1 * rs.standard_normal((nu, d))).astype(np.float32)
    i0 = (0.1 * rs.standard_normal((ni, d))).astype(np.float32)
    graph = gdr.BipartiteGraph(R.coo_indices(), R.vals, nu, ni)
    uo, io = gdr.lightgcn_propagate(graph, torch.from_numpy(u0).to(dev), torch.from_numpy(i0).to(dev), cfg["layers"])
    ur, ir = oracle.lightgcn_propagate(S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data, u0, i0, cfg["layers"])
    np.testing.assert_allclose(np_(uo), ur, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np_(io), ir, rtol=1e-5, atol=1e-6)
    # a7: per side, StandardScaler + one Lloyd step from shared centres (distill_recsys.py:158-181)
    maps = []
    for emb, n_side in ((u0, nu), (i0, ni)):
        K = int(np.ceil(0.1 * n_side))
        Xs = StandardScaler().fit_transform(emb).astype(np.float32)
        got = np_(gdr.standard_scale(torch.from_numpy(emb).to(dev)))
        np.testing.assert_allclose(got, Xs, rtol=1e-6, atol=1e-6)
        C0 = synth.kmeans_init(got, K, seed)
        lab = single_step_protocol(gdr, dev, oracle, got, C0)
        maps.append((lab.astype(np.int64), K))
    # a10: condensed bipartite counts (distill_recsys.py:184-201), duplicates included: sum C = #lines
    (u2cu, ncu), (i2ci, nci) = maps
    C = gdr.build_condensed_bipartite(u, i, u2cu, i2ci, ncu, nci, device=dev)
    Cr = sp.coo_matrix((np.ones(u.shape[0], dtype=np.float32), (u2cu[u], i2ci[i])), shape=(ncu, nci))
    Cr.sum_duplicates()
    Cr = Cr.tocsr()
    Cr.sort_indices()
    assert np.array_equal(C.indptr, Cr.indptr) and np.array_equal(C.indices, Cr.indices)
    assert np.array_equal(C.data, Cr.data) and int(C.data.sum()) == u.shape[0]
