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
for _ in range(2):
    gs.invalidate()
    gs.resolve_async(1)
gs.synchronize()
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
