"""Compare the GPU twisted sweep's per-block Y, Z, r with a numpy twisted elimination of the same matrix."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idto_b200 import capi, problems
from idto_b200.types import GRAD_CENTRAL, LINSOLVE_TWISTED
m, dt, prob, params, guess = problems.mini_cheetah(gradients_method=GRAD_CENTRAL)
rng = np.random.default_rng(7)
q = np.array(guess, float).copy(); q[1:] += rng.normal(0, 0.03, q[1:].shape)
params.linear_solver = LINSOLVE_TWISTED
gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 1)
gs.set_q(q); gs.eval(3)
nq, T = m.nq, prob.num_steps
nu = gs.get("h").shape[1] // T
kb = nq + nu
FY = gs.get("dbg_FY")[0].reshape(T + 1, kb, kb).transpose(0, 2, 1)
FZ = gs.get("dbg_FZ")[0].reshape(T + 1, kb, kb).transpose(0, 2, 1)
Fr = gs.get("dbg_Fr")[0].reshape(T + 1, kb)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "kkt_dbg.npz"), FY=FY, FZ=FZ, Fr=Fr, lam=gs.get("lambda")[0],
                    x=gs.get("dqH")[0], gs=gs.get("gs")[0], h=gs.get("h")[0])
print("saved")
