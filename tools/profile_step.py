"""Runs a few resident steps of the benchmark workload (for ncu / quick timing).

  python tools/profile_step.py [steps] [method] [batch]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from idto_b200 import capi, problems  # noqa: E402
from idto_b200.types import GRAD_CENTRAL, GRAD_FORWARD  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
method = GRAD_FORWARD if (len(sys.argv) > 2 and sys.argv[2] == "forward") else GRAD_CENTRAL
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
m, dt, prob, params, guess = problems.mini_cheetah(T=40, gradients_method=method, max_iterations=1)
model = capi.Model(m)
gs = capi.BatchSolver(model, dt, prob, params, B)
q0, v0, qg = problems.perturbed_batch(m, prob, B)
gs.reset_initial_conditions(q0, v0)
gs.set_q(qg)
for _ in range(3):
    gs.invalidate()
    gs.resolve_async(1)
gs.synchronize()
for ns in (1, 2, 4, 8):
    gs.set_substreams(ns)
    gs.invalidate(); gs.resolve_async(1); gs.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        gs.invalidate()
        gs.resolve_async(1)
    gs.synchronize()
    print(f"substreams={ns}: {(time.perf_counter() - t0) / steps * 1e3:.3f} ms/step (pipelined, wall)")
gs.set_substreams(4)
gs.profile_enable(True)
t0 = time.perf_counter()
for _ in range(steps):
    gs.invalidate()
    gs.resolve_async(1)
gs.synchronize()
el = time.perf_counter() - t0
names = ("trajectory", "id_partials", "assemble", "lagrange", "dogleg", "trajectory_scratch", "trust_update")
print(f"{steps} steps, batch {B}: {el / steps * 1e3:.3f} ms/step wall;",
      {n: round(gs.profile_read(n)[0] / max(gs.profile_read(n)[1], 1), 4) for n in names})

# ---- end-to-end step (pinned host buffers through idto_resolve_async), with a host-side breakdown ----
try:
    import torch
    from idto_b200.types import NUM_STATS
    T, T1 = 40, 41
    hq = torch.from_numpy(qg.copy()).pin_memory()
    hq0, hv0 = torch.from_numpy(q0.copy()).pin_memory(), torch.from_numpy(v0.copy()).pin_memory()
    oq = torch.empty((B, T1, m.nq), dtype=torch.float64).pin_memory()
    ov = torch.empty((B, T1, m.nv), dtype=torch.float64).pin_memory()
    ot = torch.empty((B, T, m.nv), dtype=torch.float64).pin_memory()
    ost = torch.empty((B, 1, NUM_STATS), dtype=torch.float64).pin_memory()
    gs.profile_enable(False)
    tt = [0.0, 0.0, 0.0]
    for it in range(steps + 3):
        a = time.perf_counter()
        gs.resolve_async(1, q_guess=hq.data_ptr(), q_init=hq0.data_ptr(), v_init=hv0.data_ptr(),
                         q_out=oq.data_ptr(), v_out=ov.data_ptr(), tau_out=ot.data_ptr(), stats_out=ost.data_ptr())
        b_ = time.perf_counter()
        gs.synchronize()
        c = time.perf_counter()
        hq.numpy()[...] = oq.numpy()
        d = time.perf_counter()
        if it >= 3:
            tt[0] += b_ - a; tt[1] += c - b_; tt[2] += d - c
    print("e2e per step [ms]: enqueue %.3f  wait %.3f  host copy %.3f" % tuple(x / steps * 1e3 for x in tt),
          "rho", float(ost[0, 0, 5]), "cost", float(ost[0, 0, 0]))
except ImportError:
    pass
