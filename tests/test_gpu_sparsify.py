"""Edge scoring + top-k sparsification on the GPU against the oracle and the
golden vectors generated from the reference's ER_estimator / attaw_ER_estimator / graph_sparse."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def gdr():
    import gdr as g
    assert torch.cuda.is_available()
    return g


def np_(t):
    return t.detach().cpu().numpy()


def _golden_adj(gdr):
    g = golden("sparsify.npz")
    n = int(g["n"])
    idx = torch.from_numpy(np.stack([g["src"], g["dst"]])).to(DEV)
    adj = torch.sparse_coo_tensor(idx, torch.from_numpy(g["val"]).to(DEV), (n, n))
    return g, n, adj


def _edge_set(t):
    t = t.coalesce()
    i = np_(t._indices())
    return set(zip(i[0].tolist(), i[1].tolist()))


def test_er_estimators_match_reference_golden(gdr):
    g, n, adj = _golden_adj(gdr)
    er = gdr.ER_estimator(adj, None, None)
    assert np.array_equal(np_(er), g["er"])                       # bit-exact: same fp32 sequence as the reference
    ebd = torch.from_numpy(g["ebd"]).to(DEV)
    er_att, rew = gdr.attaw_ER_estimator(adj, ebd, None, None)
    np.testing.assert_allclose(np_(rew.coalesce()._values()), g["rew_val"], rtol=0, atol=2e-7 * np.abs(g["val"]).max())
    np.testing.assert_allclose(np_(er_att), g["er_att"], rtol=1e-3, atol=1e-4 * np.abs(g["er_att"]).max())


def test_graph_sparse_matches_reference_golden(gdr):
    g, n, adj = _golden_adj(gdr)
    ratio, k = float(g["ratio"]), int(g["val"].shape[0] * float(g["ratio"]))
    ebd = torch.from_numpy(g["ebd"]).to(DEV)
    van = gdr.graph_sparse(adj, ratio, sp_type="vanilla")[0]
    assert van._nnz() == k
    assert len(_edge_set(van) ^ set(zip(g["van_row"].tolist(), g["van_col"].tolist()))) <= 2
    sin = gdr.graph_sparse(adj, ratio, ebd=ebd, sp_type="single")[0]
    assert len(_edge_set(sin) ^ set(zip(g["sin_row"].tolist(), g["sin_col"].tolist()))) <= 2
    att = gdr.graph_sparse(adj, ratio, ebd=ebd, sp_type="attaw")
    assert len(att) == int(g["C"])
    for i, t in enumerate(att):
        assert t._nnz() == k
        assert len(_edge_set(t) ^ set(zip(g[f"att{i}_row"].tolist(), g[f"att{i}_col"].tolist()))) <= 4
    assert gdr.graph_sparse(adj, ratio, sp_type="no_sp")[0] is adj


@pytest.mark.parametrize("n,pairs,C,ratio", [(3000, 40000, 7, 0.3), (20000, 300000, 40, 0.05), (500, 2000, 3, 1.0), (500, 2000, 3, 0.0)])
def test_scores_and_topk_against_oracle(gdr, oracle, n, pairs, C, ratio):
    from gdr import synth
    u, v = synth.skewed_graph(n, pairs, seed=n)
    A = gdr.sym_normalize(gdr.coo_to_csr(torch.from_numpy(u).to(DEV), torch.from_numpy(v).to(DEV), None, (n, n),
                                         symmetrize=True, binarize=True), 2)
    rp, ci, va = np_(A.rowptr), np_(A.colidx), np_(A.vals)
    ebd = (np.random.RandomState(n).randn(n, C) * 2).astype(np.float32)
    ebd_d = torch.from_numpy(ebd).to(DEV)
    # scores
    assert np.array_equal(np_(gdr.er_lower(A)), oracle.er_lower(rp, ci, va))
    R = gdr.cosine_reweight(A, ebd_d)
    er_o, rew_o = oracle.attaw_er_lower(rp, ci, va, ebd)
    np.testing.assert_allclose(np_(R.vals), rew_o, rtol=0, atol=3e-7 * np.abs(va).max())
    np.testing.assert_allclose(np_(gdr.softmax_rows(ebd_d)), oracle.softmax_rows(ebd), rtol=2e-6, atol=1e-9)
    # top-k on GIVEN weights is integer work: bit-exact against the oracle, ties included
    rs = np.random.RandomState(n + 1)
    w = rs.rand(va.shape[0]).astype(np.float32)
    w[rs.randint(0, w.shape[0], w.shape[0] // 3)] = 0.5            # many ties, also at the threshold for some k
    w[rs.randint(0, w.shape[0], 5)] = -1.0
    k = int(va.shape[0] * ratio)
    S = gdr.topk_filter(A, torch.from_numpy(w).to(DEV), k)
    rpo, cio, vo = oracle.filter_csr(rp, ci, va, oracle.topk_edges(w, k))
    assert np.array_equal(np_(S.rowptr), rpo) and np.array_equal(np_(S.colidx), cio) and np.array_equal(np_(S.vals), vo)
    # class weights: same fp32 product order
    P = gdr.softmax_rows(ebd_d)
    er_d = gdr.er_lower(R)
    wd = gdr.class_edge_weight(R, er_d, P, C - 1)
    assert np.array_equal(np_(wd), oracle.class_edge_weight(rp, ci, np_(er_d), np_(P)[:, C - 1]))


def test_graph_sparse_then_compress_pipeline(gdr):
    """graph_sparse -> graph_compress, the way ClustGDD.train chains them (clustgdd_agent_transduct.py:380-384)."""
    from gdr import synth
    n, C = 4000, 5
    u, v = synth.uniform_graph(n, 30000, seed=3)
    A = gdr.sym_normalize(gdr.coo_to_csr(torch.from_numpy(u).to(DEV), torch.from_numpy(v).to(DEV), None, (n, n),
                                         symmetrize=True, binarize=True), 2)
    ebd = torch.randn(n, C, device=DEV)
    labels = torch.randint(0, 50, (n,), device=DEV, dtype=torch.int32)
    glist = gdr.graph_sparse(A.to_torch_coo(), 0.2, ebd=ebd, sp_type="attaw")
    syn_list, syn = gdr.graph_compress(labels, A, glist)
    assert len(syn_list) == C and syn.shape == (50, 50)
    for t in syn_list:
        assert t.shape == (50, 50) and torch.isfinite(t._values()).all()


def test_rand_sparsifier_follows_torch_randperm(gdr):
    from gdr import synth
    n = 1000
    u, v = synth.uniform_graph(n, 6000, seed=5)
    A = gdr.sym_normalize(gdr.coo_to_csr(torch.from_numpy(u).to(DEV), torch.from_numpy(v).to(DEV), None, (n, n),
                                         symmetrize=True, binarize=True), 2)
    ebd = torch.randn(n, 2, device=DEV)
    torch.manual_seed(123)
    out = gdr.graph_sparse(A, 0.25, ebd=ebd, sp_type="rand")
    torch.manual_seed(123)
    k = int(A.nnz * 0.25)
    coo = A.coo_indices().cpu()
    for t in out:
        pick = torch.randperm(A.nnz)[:k]
        ref = set(zip(coo[0][pick].tolist(), coo[1][pick].tolist()))
        i = t.coalesce()._indices().cpu()
        assert set(zip(i[0].tolist(), i[1].tolist())) == ref


def test_sparsify_classes_multi_batch_equals_single_batch(gdr):
    """Large graphs process the classes in several batches (scratch is bounded): force 3 classes per batch on a
    7-class problem and compare with the single-batch result, entry for entry."""
    from gdr import synth, _lib
    n, C = 3000, 7
    u, v = synth.uniform_graph(n, 25000, seed=9)
    A = gdr.sym_normalize(gdr.coo_to_csr(torch.from_numpy(u).to(DEV), torch.from_numpy(v).to(DEV), None, (n, n),
                                         symmetrize=True, binarize=True), 2)
    ebd = torch.randn(n, C, device=DEV)
    R = gdr.cosine_reweight(A, ebd)
    er, P, k = gdr.er_lower(R), gdr.softmax_rows(ebd), A.nnz // 5
    from gdr.sparsify import sparsify_classes
    one = sparsify_classes(R, er, P, k)
    _lib.call("gdr_debug_set", b"sparsify_batch", 3)
    try:
        many = sparsify_classes(R, er, P, k)
    finally:
        _lib.call("gdr_debug_set", b"sparsify_batch", 0)
    assert len(one) == len(many) == C
    for a, b in zip(one, many):
        assert torch.equal(a.rowptr, b.rowptr) and torch.equal(a.colidx, b.colidx) and torch.equal(a.vals, b.vals)
        assert int(a.rowptr[-1]) == k
