"""Parity of the CUDA path (through the C ABI) against the oracle and the golden vectors.
Needs a B200: run with  python -m pytest tests -m gpu."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import golden

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


def csr_np(A):
    return np_(A.rowptr), np_(A.colidx), np_(A.vals)


# ---------------------------------------------------------------- primitives
@pytest.mark.parametrize("n,bits", [(0, 20), (1, 20), (31, 8), (4096, 16), (4097, 33), (100003, 40), (1 << 20, 44)])
def test_sort_pairs_stable(gdr, dev, n, bits):
    from gdr import _lib
    from gdr._dev import ptr, stream, workspace
    rs = np.random.RandomState(n + bits)
    keys = rs.randint(0, 1 << min(bits, 62), size=n, dtype=np.int64).astype(np.uint64)
    if n > 10:
        keys[: n // 2] = keys[n // 2: n // 2 + n // 2]  # many duplicates -> stability matters
    vals = np.arange(n, dtype=np.uint32)
    k = torch.from_numpy(keys.view(np.int64)).to(dev)
    v = torch.from_numpy(vals.view(np.int32)).to(dev)
    ws = workspace(_lib.query("gdr_sort_pairs_ws_bytes", n), dev)
    _lib.call("gdr_sort_pairs", n, bits, ptr(k), ptr(v), ptr(ws), ws.numel(), stream())
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(np_(k).view(np.uint64), keys[order])
    assert np.array_equal(np_(v).view(np.uint32), vals[order])


# ---------------------------------------------------------------- stage 1
def test_coo_to_csr_golden(gdr, dev):
    g = golden("csr_build.npz")
    n = int(g["n"])
    A = gdr.coo_to_csr(g["row"], g["col"], None, (n, n), device=dev)
    rp, ci, v = csr_np(A)
    assert np.array_equal(rp, g["a_indptr"]) and np.array_equal(ci, g["a_indices"]) and np.array_equal(v, g["a_data"])
    B = gdr.coo_to_csr(g["row"], g["col"], None, (n, n), symmetrize=True, binarize=True, device=dev)
    rp, ci, v = csr_np(B)
    assert np.array_equal(rp, g["b_indptr"]) and np.array_equal(ci, g["b_indices"]) and np.array_equal(v, g["b_data"])


def test_recsys_builders_golden(gdr, dev):
    g = golden("recsys_ali_subset.npz")
    nu, ni = int(g["nu"]), int(g["ni"])
    R = gdr.build_interaction_matrix(nu, ni, g["u"].astype(np.int64), g["i"].astype(np.int64), device=dev)
    assert isinstance(R, sp.csr_matrix) and R.dtype == np.float32 and R.indices.dtype == np.int32
    assert np.array_equal(R.indptr, g["r_indptr"]) and np.array_equal(R.indices, g["r_indices"])
    assert np.array_equal(R.data, g["r_data"])
    C = gdr.build_condensed_bipartite(g["u"].astype(np.int64), g["i"].astype(np.int64), g["u2cu"], g["i2ci"], 211, 97,
                                      device=dev)
    assert np.array_equal(C.indptr, g["c_indptr"]) and np.array_equal(C.indices, g["c_indices"])
    assert np.array_equal(C.data, g["c_data"])
    assert C.data.sum() == g["u"].shape[0]
    ei, ew = gdr.condensed_csr_to_edge_index(C, dev)
    coo = C.tocoo()
    assert np.array_equal(np_(ei), np.stack([coo.row, coo.col])) and np.array_equal(np_(ew), coo.data)


def test_coo_to_csr_edge_cases(gdr, dev, oracle):
    A = gdr.coo_to_csr(np.zeros(0, np.int64), np.zeros(0, np.int64), None, (5, 7), device=dev)
    assert A.nnz == 0 and np.array_equal(np_(A.rowptr), np.zeros(6, np.int32))
    with pytest.raises(ValueError):
        gdr.coo_to_csr(np.array([0, 9]), np.array([1, 1]), None, (5, 7), device=dev)
    # ragged rows, one very long row (exercises the long-run reduction), float values
    rs = np.random.RandomState(3)
    r = np.concatenate([np.full(5000, 3), rs.randint(0, 50, 3000)]).astype(np.int64)
    c = np.concatenate([np.full(5000, 2), rs.randint(0, 40, 3000)]).astype(np.int64)
    v = np.ones(r.shape[0], np.float32)
    A = gdr.coo_to_csr(r, c, v, (50, 40), device=dev)
    rp, ci, vv = oracle.coo_to_csr(r, c, v, (50, 40))
    grp, gci, gv = csr_np(A)
    assert np.array_equal(grp, rp) and np.array_equal(gci, ci) and np.array_equal(gv, vv)


@pytest.mark.parametrize("case", ["plain", "loop0", "isolated", "loop0_isolated", "weighted"])
def test_normalize_adj_tensor_golden(gdr, dev, case):
    g = golden(f"normalize_{case}.npz")
    n = int(g["n"])
    idx = torch.from_numpy(np.stack([g["in_row"], g["in_col"]])).to(dev)
    adj = torch.sparse_coo_tensor(idx, torch.from_numpy(g["in_val"]).to(dev), (n, n))
    out = gdr.normalize_adj_tensor(adj, sparse=True)
    assert gdr.is_sparse_tensor(out) and out.device.type == "cuda" and out.dtype == torch.float32
    assert np.array_equal(np_(out._indices()), g["out_idx"])
    if case in ("plain", "isolated"):
        assert np.array_equal(np_(out._values()), g["out_val"])  # bit-exact with scipy's fp64 path
    else:
        np.testing.assert_allclose(np_(out._values()), g["out_val"], rtol=5e-7, atol=0)  # numpy powf: <= 2 ulp
    back = gdr.to_scipy(out)
    assert back.nnz == g["out_val"].shape[0]


def test_normalize_dense_golden(gdr, dev):
    g = golden("normalize_dense.npz")
    out = gdr.normalize_adj_tensor(torch.from_numpy(g["a"]).to(dev))
    np.testing.assert_allclose(np_(out), g["out"], rtol=3e-7, atol=0)


def test_stage1_midsize_vs_oracle(gdr, dev, oracle):
    from gdr import synth
    n = 20000
    u, v = synth.skewed_graph(n, 150000, seed=5)
    A = gdr.coo_to_csr(u, v, None, (n, n), symmetrize=True, binarize=True, device=dev)
    rp, ci, va = oracle.coo_to_csr(u, v, None, (n, n), symmetrize=True, binarize=True)
    grp, gci, gva = csr_np(A)
    assert np.array_equal(grp, rp) and np.array_equal(gci, ci) and np.array_equal(gva, va)
    An = gdr.sym_normalize(A, 2)
    rpo, cio, vo, deg = oracle.sym_normalize(rp, ci, va, n)
    grp, gci, gva = csr_np(An)
    assert np.array_equal(grp, rpo) and np.array_equal(gci, cio)
    assert np.array_equal(gva, vo)
    assert np.array_equal(np_(An.deg), deg)
    # idempotence of the build: CSR -> COO -> CSR is the identity
    idx = An.coo_indices()
    A2 = gdr.coo_to_csr(idx[0], idx[1], An.vals, (n, n), device=dev)
    assert np.array_equal(np_(A2.rowptr), grp) and np.array_equal(np_(A2.colidx), gci) and np.array_equal(np_(A2.vals), gva)
    # transpose of a symmetric matrix is itself
    AT, perm = An.transpose()
    assert np.array_equal(np_(AT.rowptr), grp) and np.array_equal(np_(AT.colidx), gci)


# ---------------------------------------------------------------- stage 2
def test_propagate_golden(gdr, dev):
    g = golden("propagate.npz")
    n = int(g["n"])
    idx = torch.from_numpy(np.stack([g["in_row"], g["in_col"]])).to(dev)
    adj = torch.sparse_coo_tensor(idx, torch.from_numpy(g["in_val"]).to(dev), (n, n))
    adj_norm = gdr.normalize_adj_tensor(adj, sparse=True)
    prop, target = gdr.propagate(adj_norm, torch.from_numpy(g["x"]).to(dev), int(g["T"]), float(g["alpha"]))
    scale = np.abs(g["x"]).max()
    np.testing.assert_allclose(np_(prop), g["prop"], rtol=1e-5, atol=1e-6 * scale)
    np.testing.assert_allclose(np_(target), g["target"], rtol=1e-5, atol=1e-6 * scale)


@pytest.mark.parametrize("f", [1, 7, 32, 64, 100, 128, 200, 500, 1433])
def test_propagate_widths_vs_oracle(gdr, dev, oracle, f):
    from gdr import synth
    n = 3000
    u, v = synth.skewed_graph(n, 20000, seed=f)
    A = gdr.sym_normalize(gdr.coo_to_csr(u, v, None, (n, n), symmetrize=True, binarize=True, device=dev), 2)
    X = synth.features(n, f, seed=f + 1, kind="none")
    rp, ci, va = csr_np(A)
    for T in (1, 2, 4):
        prop, target = gdr.propagate(A, torch.from_numpy(X).to(dev), T, 0.7)
        p_ref, t_ref = oracle.propagate(rp, ci, va, X, T, 0.7)
        scale = np.abs(X).max()
        np.testing.assert_allclose(np_(prop), p_ref, rtol=1e-5, atol=1e-6 * scale)
        np.testing.assert_allclose(np_(target), t_ref, rtol=1e-5, atol=1e-6 * scale)


def test_spmm_linearity_and_empty_rows(gdr, dev):
    from gdr import synth
    n = 5000
    u, v = synth.uniform_graph(n, 10000, seed=9)  # sparse: many empty rows
    A = gdr.coo_to_csr(u, v, None, (n, n), device=dev)
    x = torch.randn(n, 64, device=dev)
    y = torch.randn(n, 64, device=dev)
    lhs = gdr.spmm(A, x + y)
    rhs = gdr.spmm(A, x) + gdr.spmm(A, y)
    torch.testing.assert_close(lhs, rhs, rtol=1e-5, atol=1e-5)
    deg = np.diff(np_(A.rowptr))
    assert (np_(lhs)[deg == 0] == 0).all()


# ---------------------------------------------------------------- stage 3
def test_kmeans_golden_fit(gdr, dev):
    g = golden("kmeans.npz")
    km = gdr.KMeans(n_clusters=40, init=g["c0"], n_init=1, max_iter=1, tol=0).fit(g["x"])
    assert isinstance(km.labels_, np.ndarray) and km.labels_.dtype == np.int32
    assert np.array_equal(km.labels_, g["it1_labels"])
    np.testing.assert_allclose(km.cluster_centers_, g["it1_centers"], rtol=1e-5, atol=1e-6)
    assert abs(km.inertia_ - float(g["it1_inertia"])) <= 1e-4 * float(g["it1_inertia"])
    km = gdr.KMeans(n_clusters=40, init=g["c0"], n_init=1, max_iter=300, tol=1e-4).fit(g["x"])
    assert km.n_iter_ == int(g["fit_n_iter"])
    assert np.array_equal(km.labels_, g["fit_labels"])
    np.testing.assert_allclose(km.cluster_centers_, g["fit_centers"], rtol=1e-5, atol=1e-5)
    assert abs(km.inertia_ - float(g["fit_inertia"])) <= 1e-4 * float(g["fit_inertia"])
    km = gdr.KMeans(n_clusters=40, init=g["c0"], n_init=1, max_iter=12, tol=0).fit(g["x"])
    assert km.n_iter_ == int(g["tol0_n_iter"]) and np.array_equal(km.labels_, g["tol0_labels"])
    assert np.array_equal(km.predict(g["x"]), km.labels_)


def test_kmeans_empty_cluster_relocation_golden(gdr, dev):
    g = golden("kmeans.npz")
    km = gdr.KMeans(n_clusters=40, init=g["c0_empty"], n_init=1, max_iter=1, tol=0).fit(g["x"])
    a = np.sort(km.cluster_centers_.round(4), axis=0)
    b = np.sort(g["empty_centers"].round(4), axis=0)
    np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-4)
    assert abs(km.inertia_ - float(g["empty_inertia"])) <= 1e-4 * float(g["empty_inertia"])


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_kmeans_empty_clusters_filled_in_place_golden(gdr, dev, precision):
    """sklearn's own fit on the degenerate case (n_clusters > distinct rows): skipped relocation, in-place
    _average_centers (an empty cluster below the largest one gets its raw SUM), then a real relocation."""
    g = golden("kmeans_edge.npz")
    for it in (1, 2, 5):
        km = gdr.KMeans(n_clusters=6, init=g["c0"], n_init=1, max_iter=it, tol=0, precision=precision).fit(g["x"])
        assert km.n_iter_ == int(g[f"it{it}_n_iter"])
        assert np.array_equal(km.labels_, g[f"it{it}_labels"])
        np.testing.assert_allclose(km.cluster_centers_, g[f"it{it}_centers"], rtol=1e-5, atol=2e-6)


def test_kmeans_errors(gdr, dev):
    x = np.random.RandomState(0).randn(10, 3).astype(np.float32)
    with pytest.raises(ValueError):
        gdr.KMeans(n_clusters=11, init="random").fit(x)
    with pytest.raises(ValueError):
        gdr.KMeans(n_clusters=3, init=np.zeros((3, 4), np.float32)).fit(x)
    with pytest.raises(ValueError):
        gdr.kmeans_cluster(x, 0, seed=0)


@pytest.mark.parametrize("N,K,D", [(1, 1, 1), (300, 7, 3), (5000, 140, 7), (20000, 1000, 40), (20000, 333, 128),
                                   (4099, 129, 130)])
def test_lloyd_step_vs_oracle(gdr, dev, oracle, N, K, D):
    """Single Lloyd step from shared centres (the parity protocol of SURVEY §8c)."""
    from gdr import synth
    from gdr._dev import padded_rows
    X = synth.clustered_features(N, D, max(2, K // 3), seed=N + K)
    X -= X.mean(axis=0)
    C = synth.kmeans_init(X, K, seed=1)
    Xd, Cd = padded_rows(torch.from_numpy(X).to(dev)), padded_rows(torch.from_numpy(C).to(dev))
    labels = torch.empty(N, dtype=torch.int32, device=dev)
    best = torch.empty(N, dtype=torch.float32, device=dev)
    gdr.assign_labels(Xd, Cd, labels, best=best)
    lab = np_(labels)
    ok, n_band, n_bad = oracle.labels_match(lab, X, C, band=1e-6)
    assert ok, f"{n_bad} rows outside the 1e-6 margin band disagree with the exact argmin"
    l32, b32 = oracle.kmeans_assign(X, C)
    np.testing.assert_allclose(np_(best), b32, rtol=1e-4, atol=1e-4 * np.abs(b32).max())
    # M-step: bit-exact with the sequential fp32 oracle given equal labels
    sums, counts = gdr.segment_sum(Xd, labels, K)
    s_ref, c_ref = oracle.segment_sum(X, lab, K)
    assert np.array_equal(np_(counts), c_ref)
    assert np.array_equal(np_(sums), s_ref)


def test_cluster_means_and_graph_compress_golden(gdr, dev):
    g = golden("graph_compress.npz")
    n = int(g["n"])
    idx = torch.from_numpy(np.stack([g["in_row"], g["in_col"]])).to(dev)
    adj = torch.sparse_coo_tensor(idx, torch.from_numpy(g["in_val"]).to(dev), (n, n))
    adj_norm = gdr.normalize_adj_tensor(adj, sparse=True)
    labels = torch.from_numpy(g["labels"]).to(dev)
    means = np_(gdr.cluster_means(torch.from_numpy(g["feat"]).to(dev), labels, 20))
    assert np.isnan(means[13]).all()
    ok = ~np.isnan(g["means"])
    np.testing.assert_allclose(means[ok], g["means"][ok], rtol=1e-5, atol=1e-6)
    glist, adj_syn = gdr.graph_compress(labels, adj_norm, [adj_norm])
    assert len(glist) == 1 and gdr.is_sparse_tensor(adj_syn) and tuple(adj_syn.shape) == g["syn_dense"].shape
    ref = g["syn_dense"]
    fin = np.isfinite(ref)
    got = np_(adj_syn.to_dense())
    np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(np_(glist[0].to_dense())[fin], g["list0_dense"][fin], rtol=1e-5, atol=1e-8)
    # structure: same non-zero cells as the reference's .to_sparse() (NaN cells excluded)
    ref_nz = set(map(tuple, g["syn_idx"].T[np.isfinite(g["syn_val"])]))
    got_nz = set(map(tuple, np_(adj_syn._indices()).T))
    assert got_nz == ref_nz


def test_coarsen_counts_vs_oracle(gdr, dev, oracle):
    from gdr import synth
    u, i = synth.bipartite_interactions(3000, 4000, 200000, seed=2)
    rs = np.random.RandomState(2)
    for (ku, ki) in [(1, 1), (3, 2), (300, 400), (3000, 4000)]:
        u2 = rs.randint(0, ku, 3000)
        i2 = rs.randint(0, ki, 4000)
        C = gdr.build_condensed_bipartite(u, i, u2, i2, ku, ki, device=dev)
        rp, ci, cnt, _ = oracle.coarsen_counts(u, i, u2, i2, ku, ki)
        assert np.array_equal(C.indptr, rp) and np.array_equal(C.indices, ci)
        assert np.array_equal(C.data, cnt.astype(np.float32))
        assert C.data.sum() == u.shape[0]


# ---------------------------------------------------------------- bipartite propagation
def test_lightgcn_golden(gdr, dev):
    g = golden("lightgcn.npz")
    ei = torch.from_numpy(g["edge_index"]).to(dev)
    graph = gdr.BipartiteGraph(ei, torch.from_numpy(g["w"]).to(dev), g["u0"].shape[0], g["i0"].shape[0])
    u, i = gdr.lightgcn_propagate(graph, torch.from_numpy(g["u0"]).to(dev), torch.from_numpy(g["i0"]).to(dev),
                                  int(g["layers"]))
    np.testing.assert_allclose(np_(u), g["u_out"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np_(i), g["i_out"], rtol=1e-5, atol=1e-6)


def test_bipartite_normalize_vs_oracle(gdr, dev, oracle):
    from gdr import synth
    u, i = synth.bipartite_interactions(500, 300, 20000, seed=4)
    rp, ci, w = oracle.coo_to_csr(u, i, None, (500, 300))
    norm, du, di = oracle.bipartite_normalize(rp, ci, w, 500, 300)
    R = gdr.coo_to_csr(u, i, None, (500, 300), device=dev)
    graph = gdr.BipartiteGraph(R.coo_indices(), R.vals, 500, 300)
    assert np.array_equal(np_(graph.deg_u), du) and np.array_equal(np_(graph.deg_i), di)
    assert np.array_equal(np_(graph.A.vals), norm)  # IEEE add/sqrt/mul/div: bit-exact


# ---------------------------------------------------------------- full-size properties (config B)
def test_config_b_properties(gdr, dev):
    from gdr import synth
    cfg = synth.CONFIGS["B"]
    n = cfg["n"]
    u, v = synth.uniform_graph(n, cfg["pairs"], seed=1235)
    A = gdr.coo_to_csr(u, v, None, (n, n), symmetrize=True, binarize=True, device=dev)
    rp, ci = np_(A.rowptr).astype(np.int64), np_(A.colidx)
    assert rp[0] == 0 and rp[-1] == A.nnz and (np.diff(rp) >= 0).all()
    rows = np.repeat(np.arange(n), np.diff(rp))
    key = rows * n + ci
    assert (np.diff(key) > 0).all()                       # sorted, no duplicates
    An = gdr.sym_normalize(A, 2)
    assert An.nnz == A.nnz + n                            # identity added (node 0 has no self-loop)
    # D^1/2 A_hat D^1/2 1 = deg  =>  sum_j A_hat_ij sqrt(d_j) = sqrt(d_i)
    sq = An.deg.sqrt().to(torch.float32).unsqueeze(1).contiguous()
    lhs = gdr.spmm(An, sq.expand(n, 4).contiguous())[:, 0]
    torch.testing.assert_close(lhs, sq[:, 0], rtol=1e-5, atol=1e-5)
    X = torch.from_numpy(synth.clustered_features(n, cfg["f"], 200, seed=7)).to(dev)
    prev = None
    C0 = synth.kmeans_init(np_(X), cfg["k"], 1235)
    for it in (1, 2, 4):
        km = gdr.KMeans(n_clusters=cfg["k"], init=C0, n_init=1, max_iter=it, tol=0).fit(X)
        assert prev is None or km.inertia_ <= prev * (1 + 1e-6)   # WCSS monotone non-increasing
        prev = km.inertia_
    labels = km.labels_
    _, adj_syn = gdr.graph_compress(labels, An, [])
    k = int(labels.max()) + 1
    rpc, cic, cnt, _ = gdr.coarsen_edges(labels, labels, k, k, csr=An)
    assert int(cnt.sum()) == An.nnz                        # every edge lands in exactly one cell


def test_block_build_world1_equals_global_build(gdr, dev):
    """dist_build_adjacency with a single rank (no collectives) runs the row-block kernels
    (gdr_sym_normalize_block_*) on the whole matrix: same CSR as coo_to_csr + sym_normalize."""
    from gdr import parallel as par
    from gdr import synth
    for n, pairs, seed in [(5003, 40000, 1), (2000, 300, 2)]:      # the second one has isolated nodes
        u, v = synth.skewed_graph(n, pairs, seed=seed)
        u_d, v_d = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
        A = gdr.sym_normalize(gdr.coo_to_csr(u_d, v_d, None, (n, n), symmetrize=True, binarize=True), 2)
        B = par.dist_build_adjacency(par.Comm(), par.RowPartition(n, 1, 0), u_d, v_d, n)
        assert torch.equal(A.rowptr, B.rowptr) and torch.equal(A.colidx, B.colidx) and torch.equal(A.vals, B.vals)


def test_config_e_properties(gdr, dev):
    """BASELINE full size (config E, ogbn-products-shaped: 2.45 M nodes, 61.9 M pairs): size-independent properties
    of stages 1, 2 and 4 — sorted duplicate-free CSR, nnz(A_hat) = nnz(A) + N, symmetric pattern, the normalisation
    identity sum_j A_hat_ij sqrt(d_j) = sqrt(d_i) through the SpMM kernel, and the coarsening checksum."""
    from gdr import synth
    cfg = synth.CONFIGS["E"]
    n = cfg["n"]
    u, v = synth.uniform_graph(n, cfg["pairs"], seed=1238)
    A = gdr.coo_to_csr(torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev), None, (n, n), symmetrize=True, binarize=True)
    rp = A.rowptr.long()
    assert int(rp[0]) == 0 and int(rp[-1]) == A.nnz and bool((rp[1:] >= rp[:-1]).all())
    rows = torch.repeat_interleave(torch.arange(n, device=dev), rp[1:] - rp[:-1])
    key = rows * n + A.colidx.long()
    assert bool((key[1:] > key[:-1]).all())                                   # sorted, no duplicates
    tkey = A.colidx.long() * n + rows                                         # pattern of the transpose
    assert torch.equal(torch.sort(tkey).values, key)                          # symmetric
    assert bool((A.vals == 1).all())
    del key, tkey
    An = gdr.sym_normalize(A, 2)
    assert An.nnz == A.nnz + n
    sq = An.deg.sqrt().to(torch.float32).unsqueeze(1)
    lhs = gdr.spmm(An, sq.expand(n, 4).contiguous())[:, 0]
    torch.testing.assert_close(lhs, sq[:, 0], rtol=2e-5, atol=2e-5)
    labels = torch.from_numpy(np.random.RandomState(5).randint(0, cfg["k"], n).astype(np.int32)).to(dev)
    k = cfg["k"]
    rpc, cic, cnt, _ = gdr.coarsen_edges(labels, labels, k, k, csr=An)
    assert int(cnt.long().sum()) == An.nnz                                     # every edge lands in exactly one cell
    crow = torch.repeat_interleave(torch.arange(k, device=dev), (rpc[1:] - rpc[:-1]).long())
    ckey = crow * k + cic.long()
    assert bool((ckey[1:] > ckey[:-1]).all())
