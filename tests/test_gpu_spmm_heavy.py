"""Hub rows (nnz > 1024) take the CTA-cooperative path of k_spmm; isolated nodes, a dense row and
every feature-width class are checked against the oracle.  GPU only."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("f", [4, 20, 64, 100, 128, 256, 1100])
def test_spmm_hub_rows_vs_oracle(oracle, f):
    import gdr
    rs = np.random.RandomState(f)
    n = 6000
    # three hubs (degrees ~5000, ~2500, ~1030) + a sparse background + isolated tail
    src = np.concatenate([np.full(5000, 7), np.full(2500, 64), np.full(1030, 4999), rs.randint(0, 5000, 20000)])
    dst = np.concatenate([rs.choice(n, 5000, replace=False), rs.choice(n, 2500, replace=False),
                          rs.choice(n, 1030, replace=False), rs.randint(0, 5000, 20000)])
    val = rs.rand(src.shape[0]).astype(np.float32)
    A = gdr.coo_to_csr(src, dst, val, (n, n), device=DEV)
    rp, ci, va = (t.cpu().numpy() for t in (A.rowptr, A.colidx, A.vals))
    assert np.diff(rp).max() > 1024
    X = rs.standard_normal((n, f)).astype(np.float32)
    T0 = rs.standard_normal((n, f)).astype(np.float32)
    Xd = torch.from_numpy(X).to(DEV)
    from gdr._dev import padded_rows
    Td = padded_rows(torch.from_numpy(T0).to(DEV)).clone() if f % 4 == 0 else None
    if Td is None:
        buf = torch.zeros((n, (f + 3) // 4 * 4), device=DEV)
        buf[:, :f] = torch.from_numpy(T0).to(DEV)
        Td = buf[:, :f]
    y = gdr.spmm(A, Xd, alpha=0.7, accumulate_into=Td, beta=0.3)
    T_ref = T0.copy()
    y_ref = oracle.spmm_prop(rp, ci, va, np.float32(0.7), X, T=T_ref, beta=np.float32(0.3))
    scale = np.abs(y_ref).max()
    np.testing.assert_allclose(y.cpu().numpy(), y_ref, rtol=1e-5, atol=1e-6 * scale)
    np.testing.assert_allclose(Td.cpu().numpy(), T_ref, rtol=1e-5, atol=1e-6 * scale)
    # run-to-run determinism of the split reduction
    y2 = gdr.spmm(A, Xd, alpha=0.7)
    y3 = gdr.spmm(A, Xd, alpha=0.7)
    assert torch.equal(y2, y3)
