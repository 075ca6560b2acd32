"""KKT residuals of the GPU solve (both linear solvers) on the wiggled cheetah test state."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idto_b200 import capi, problems
from idto_b200.types import GRAD_CENTRAL, LINSOLVE_THOMAS, LINSOLVE_TWISTED

name = sys.argv[1] if len(sys.argv) > 1 else "mini_cheetah"
m, dt, prob, params, guess = getattr(problems, name)(gradients_method=GRAD_CENTRAL)
rng = np.random.default_rng(7)
q = np.array(guess, float).copy()
q[1:] += rng.normal(0, 0.03, q[1:].shape)
res = {}
for ls in (LINSOLVE_THOMAS, LINSOLVE_TWISTED):
    params.linear_solver = ls
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 1)
    gs.set_q(q)
    gs.eval(3)
    nq, T = m.nq, prob.num_steps
    nblk, n = T + 1, (T + 1) * m.nq
    A, B, C = (gs.get(f)[0].reshape(nblk, nq, nq).transpose(0, 2, 1) for f in ("Hs_A", "Hs_B", "Hs_C"))
    H = np.zeros((n, n))
    for i in range(nblk):
        H[i*nq:(i+1)*nq, i*nq:(i+1)*nq] = C[i]
        if i >= 1: H[i*nq:(i+1)*nq, (i-1)*nq:i*nq] = B[i]; H[(i-1)*nq:i*nq, i*nq:(i+1)*nq] = B[i].T
        if i >= 2: H[i*nq:(i+1)*nq, (i-2)*nq:(i-1)*nq] = A[i]; H[(i-2)*nq:(i-1)*nq, i*nq:(i+1)*nq] = A[i].T
    nh = gs.get("h").shape[1]
    J = gs.get("J")[0].reshape(n, nh).T
    g, h, lam, x = gs.get("gs")[0], gs.get("h")[0], gs.get("lambda")[0], gs.get("dqH")[0]
    K = np.block([[H, J.T], [J, np.zeros((nh, nh))]])
    sol = np.linalg.solve(K, -np.concatenate([g, h]))
    r1 = H @ x + J.T @ lam + g
    r2 = J @ x + h
    print(f"solver {ls}: |r1|={np.abs(r1).max():.3e} |r2|={np.abs(r2).max():.3e} |g|={np.abs(g).max():.3e} "
          f"|lam-dense|={np.abs(lam - sol[n:]).max():.3e} |x-dense|={np.abs(x - sol[:n]).max():.3e} "
          f"max|lam|={np.abs(lam).max():.3e} cond(K)={np.linalg.cond(K):.2e}")
    res[ls] = (lam, x)
    nu = nh // T
    d = np.abs(lam - sol[n:]).reshape(T, nu).max(axis=1)
    print("  per-step |lam - dense|:", np.array2string(d, precision=1, max_line_width=200))
    if ls == LINSOLVE_THOMAS and os.environ.get("KKT_DUMP"):
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"kkt_{name}.npz"), H=H, J=J, g=g, h=h, lam=lam, x=x,
                            nq=nq, T=T, nu=nu)
