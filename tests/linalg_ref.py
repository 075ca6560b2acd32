"""Dense numpy/LAPACK restatement of the reference's linear algebra downstream of the partials, used to check
the GPU's KKT sweep on the GPU's OWN inputs (so finite-difference noise of the partials cancels out):
  lambda = (J~ H~^-1 J~^T)^-1 (h - (H~^-1 J~^T)^T g~)      CalcLagrangeMultipliers, cc:1371-1396
  gm     = g~ + J~^T lambda                                 CalcMeritFunctionGradient, cc:1435-1443
  dqH    = H~^-1 (-gm)   (= pH * Delta, cc:2137-2152; not rescaled by D)
TEST INFRASTRUCTURE (imports nothing from the product)."""
import numpy as np


def dense_penta(A, B, C):
    """Symmetric block penta-diagonal matrix from its lower bands [nblk][k*k] (column-major blocks;
    penta_diagonal_matrix.cc:64-105 MakeSymmetric)."""
    nblk, kk = C.shape
    k = int(round(np.sqrt(kk)))
    H = np.zeros((nblk * k, nblk * k))
    blk = lambda X, i: X[i].reshape(k, k).T  # column-major -> (row, col)
    for i in range(nblk):
        H[i * k:(i + 1) * k, i * k:(i + 1) * k] = blk(C, i)
        if i >= 1:
            H[i * k:(i + 1) * k, (i - 1) * k:i * k] = blk(B, i)
            H[(i - 1) * k:i * k, i * k:(i + 1) * k] = blk(B, i).T
        if i >= 2:
            H[i * k:(i + 1) * k, (i - 2) * k:(i - 1) * k] = blk(A, i)
            H[(i - 2) * k:(i - 1) * k, i * k:(i + 1) * k] = blk(A, i).T
    return H


def kkt_reference(HsA, HsB, HsC, gs, J, h):
    """Returns lambda, gm, dqH, cond(H~), cond(S) by the reference's route (dense, LAPACK)."""
    H = dense_penta(HsA, HsB, HsC)
    n = H.shape[0]
    if J is None or J.size == 0:
        lam = np.zeros(0)
        gm = gs.copy()
        condS = 1.0
    else:
        nh = h.size
        Jd = J.reshape(n, nh).T  # stored column-major (nh x n)
        X = np.linalg.solve(H, Jd.T)
        S = Jd @ X
        lam = np.linalg.solve(S, h - X.T @ gs)
        gm = gs + Jd.T @ lam
        condS = np.linalg.cond(S)
    dqH = np.linalg.solve(H, -gm)
    return lam, gm, dqH, np.linalg.cond(H), condS
