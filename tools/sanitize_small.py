"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): every kernel family once.
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idto_b200 import capi, problems
from idto_b200.types import GRAD_CENTRAL, GRAD_FORWARD

for name, T, method, B in (("hopper", 9, GRAD_CENTRAL, 2), ("mini_cheetah", 8, GRAD_CENTRAL, 2),
                           ("spinner", 8, GRAD_FORWARD, 1), ("allegro_hand", 6, GRAD_FORWARD, 2),
                           ("allegro_hand", 5, GRAD_CENTRAL, 1)):
    m, dt, prob, params, guess = getattr(problems, name)(T=T, gradients_method=method)
    params.max_iterations = 2
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, B)
    gs.set_q(np.array(guess))
    it, reason, stats = gs.solve(2)
    q, v, tau = gs.solution()
    gs.mpc_advance(0.5 * dt, q[:, 0], v[:, 0])
    gs.resolve_async(1)
    gs.synchronize()
    # the same through pinned host buffers (read / written by kernels through their mapped aliases)
    import torch
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hq0, hv0, hel = pin(q[:, 0]), pin(v[:, 0]), pin(np.full(B, 0.5 * dt))
    oq, ov, ot = pin(np.zeros_like(q)), pin(np.zeros_like(v)), pin(np.zeros_like(tau))
    for _ in range(3):  # eager, capture, replay
        gs.mpc_advance(hel.numpy(), hq0.numpy(), hv0.numpy())
        gs.resolve_async(1, q_out=oq.data_ptr(), v_out=ov.data_ptr(), tau_out=ot.data_ptr())
        gs.synchronize()
    assert np.isfinite(oq.numpy()).all()
    print(name, "iters", it.tolist(), "cost", stats[0, -1, 0])
