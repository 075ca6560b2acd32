"""N resident re-solves of the benchmark workload and nothing else (ncu target).
   python tools/simple_steps.py [steps] [central|forward] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idto_b200 import capi, problems  # noqa: E402
from idto_b200.types import GRAD_CENTRAL, GRAD_FORWARD  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
method = GRAD_FORWARD if (len(sys.argv) > 2 and sys.argv[2] == "forward") else GRAD_CENTRAL
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
m, dt, prob, params, guess = problems.mini_cheetah(T=40, gradients_method=method, max_iterations=1)
gs = capi.BatchSolver(capi.Model(m), dt, prob, params, B)
q0, v0, qg = problems.perturbed_batch(m, prob, B)
gs.reset_initial_conditions(q0, v0)
gs.set_q(qg)
for _ in range(steps):
    gs.invalidate()
    gs.resolve_async(1)
gs.synchronize()
print("done", steps, "steps; launches", gs.launch_count())
