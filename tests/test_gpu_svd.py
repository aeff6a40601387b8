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


@pytest.mark.parametrize("shape,k", [((300, 100), 16), ((2000, 400), 64), ((120, 900), 24), ((40, 30), 8), ((50, 9), 8)])
def test_svd_small_matrices_match_scipy(shape, k):
    """Matrices whose smaller side is below the default Krylov dimension (q b ~ 600): the block count is clamped to the
    side (the unclamped recursion factorised rounding noise) and the result equals scipy's svds."""
    import gdr
    from scipy.sparse.linalg import svds
    rs = np.random.RandomState(shape[0] + k)
    M = sp.random(shape[0], shape[1], density=0.15, format="csr", random_state=rs, dtype=np.float32)
    M.data[:] = 1.0
    ue, ie = gdr.compute_svd_embeddings(M, k, seed=3)
    kk = min(k, min(shape) - 1)
    assert ue.shape == (shape[0], kk) and ie.shape == (shape[1], kk)
    U, S, VT = svds(M.astype(np.float64), k=kk)
    S = np.sort(S)[::-1]
    np.testing.assert_allclose((ue.astype(np.float64) ** 2).sum(0), S, rtol=1e-4)
    np.testing.assert_allclose((ie.astype(np.float64) ** 2).sum(0), S, rtol=1e-4)
    ref = (U[:, ::-1] * np.sqrt(S)) @ (VT[::-1].T * np.sqrt(S)).T
    got = ue.astype(np.float64) @ ie.astype(np.float64).T
    assert np.linalg.norm(got - ref) <= 2e-3 * np.linalg.norm(ref)


def test_dense_fp64_kernels_against_numpy():
    """The hand-written dense steps (csrc/svd.cu) one by one: Gram / projection products, Cholesky, triangular solve,
    tall x small product and the Jacobi eigen-solver."""
    import ctypes
    import gdr
    from gdr import _lib
    from gdr._dev import ptr, stream, workspace
    from gdr import svd as gsvd
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(0)
    A = rs.standard_normal((5003, 37))
    B = rs.standard_normal((5003, 70))
    Ad, Bd = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    np.testing.assert_allclose(gsvd._gram(Ad, Bd).cpu().numpy(), A.T @ B, rtol=1e-12, atol=1e-10)
    # column slices of a wider matrix (leading dimension > width)
    wide = torch.from_numpy(np.concatenate([A, B], axis=1)).to(dev)
    np.testing.assert_allclose(gsvd._gram(wide[:, :37], wide[:, 37:]).cpu().numpy(), A.T @ B, rtol=1e-12, atol=1e-10)
    # CholeskyQR2: orthonormal columns spanning the same space
    Y = torch.from_numpy(B.copy()).to(dev)
    assert gsvd._cholqr2(Y)
    Yh = Y.cpu().numpy()
    np.testing.assert_allclose(Yh.T @ Yh, np.eye(70), atol=1e-13)
    np.testing.assert_allclose(Yh @ (Yh.T @ B), B, rtol=1e-10, atol=1e-10)
    # a rank-deficient block is reported, not factorised
    Bad = torch.from_numpy(np.concatenate([B[:, :10], B[:, :10]], axis=1)).to(dev)
    assert not gsvd._cholqr2(Bad)
    # block Gram-Schmidt
    Z = torch.from_numpy(A.copy()).to(dev)
    gsvd._project_out(Z, Y)
    assert np.abs(Yh.T @ Z.cpu().numpy()).max() < 1e-11
    # Jacobi eigen-decomposition of a 301 x 301 (odd size) symmetric matrix
    n = 301
    Mx = rs.standard_normal((n, n))
    Sym = Mx @ Mx.T
    T = torch.from_numpy(Sym.copy()).to(dev)
    W = torch.empty((n, n), dtype=torch.float64, device=dev)
    ev = torch.empty(n, dtype=torch.float64, device=dev)
    order = torch.empty(n, dtype=torch.int32, device=dev)
    sweeps = ctypes.c_int32(0)
    ws = workspace(_lib.query("gdr_sym_eig_jacobi_ws_bytes", n), dev)
    _lib.call("gdr_sym_eig_jacobi", n, ptr(T), ptr(W), ptr(ev), ptr(order), 30, 1e-14, ctypes.addressof(sweeps), ptr(ws),
              ws.numel(), stream())
    ref = np.sort(np.linalg.eigvalsh(Sym))[::-1]
    np.testing.assert_allclose(ev.cpu().numpy(), ref, rtol=1e-10, atol=1e-9 * ref[0])
    Wh, oh = W.cpu().numpy(), order.cpu().numpy()
    np.testing.assert_allclose(Wh.T @ Wh, np.eye(n), atol=1e-12)
    Vs = Wh[:, oh]
    np.testing.assert_allclose(Sym @ Vs[:, :5], Vs[:, :5] * ev.cpu().numpy()[:5], rtol=1e-9, atol=1e-8 * ref[0])
    assert 1 <= sweeps.value < 30
