"""Truncated SVD embeddings against the reference's compute_svd_embeddings run on its own
shipped Ali-Display file (tests/golden/svd_ali.npz) and against scipy's svds on a synthetic matrix."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


def test_svd_matches_scipy_svds_on_synthetic():
    import gdr
    from scipy.sparse.linalg import svds
    rs = np.random.RandomState(3)
    nu, ni, k = 3000, 2000, 16
    # low-rank structure + noise so that the leading values are separated
    P = rs.rand(nu, 24) < 0.08
    Qm = rs.rand(24, ni) < 0.08
    M = sp.csr_matrix(((P.astype(np.float32) @ Qm.astype(np.float32)) > 0).astype(np.float32))
    ue, ie = gdr.compute_svd_embeddings(M, k, seed=1)
    U, S, VT = svds(M.astype(np.float64), k=k)
    S = np.sort(S)[::-1]
    sig = (ue.astype(np.float64) ** 2).sum(0)
    np.testing.assert_allclose(sig, S, rtol=1e-4)
    # R ~ U S V^T on the leading subspace: user_emb @ item_emb^T reproduces the rank-k reconstruction
    ref = (U[:, ::-1] * np.sqrt(S)) @ (VT[::-1].T * np.sqrt(S)).T
    got = ue.astype(np.float64) @ ie.astype(np.float64).T
    assert np.linalg.norm(got - ref) <= 2e-3 * np.linalg.norm(ref)
    with pytest.raises(ValueError):
        gdr.compute_svd_embeddings(sp.csr_matrix((1, 5), dtype=np.float32), 4)


def test_svd_matches_reference_on_ali_display_subset():
    """Golden: the reference's own compute_svd_embeddings on the Ali-Display subset of the build-stage fixtures."""
    import gdr
    g = golden("svd_ali.npz")
    sub = golden("recsys_ali_subset.npz")
    R = gdr.build_interaction_matrix(int(sub["nu"]), int(sub["ni"]), sub["u"], sub["i"])
    ue, ie = gdr.compute_svd_embeddings(R, int(g["dim"]), seed=42)
    assert ue.shape == (int(sub["nu"]), int(g["dim"])) and ie.shape == (int(sub["ni"]), int(g["dim"])) and ue.dtype == np.float32
    sig = (ue.astype(np.float64) ** 2).sum(0)
    np.testing.assert_allclose(sig, g["sigma"], rtol=1e-4)
    np.testing.assert_allclose((ie.astype(np.float64) ** 2).sum(0), g["sigma"], rtol=1e-4)
    for j in range(4):   # well separated leading vectors: equal up to sign
        for mine, ref in ((ue[:, j], g["user_emb4"][:, j]), (ie[:, j], g["item_emb4"][:, j])):
            a, b = mine.astype(np.float64), ref.astype(np.float64)
            assert abs(a @ b) / (np.linalg.norm(a) * np.linalg.norm(b)) > 0.999
