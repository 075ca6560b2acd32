"""Where the end-to-end step spends its time beyond the resident step (wall clock, B200)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idto_b200 import capi, problems
from idto_b200.types import GRAD_CENTRAL, NUM_STATS
B, T = 64, 40
m, dt, prob, params, guess = problems.mini_cheetah(T=T, gradients_method=GRAD_CENTRAL, max_iterations=1)
gs = capi.BatchSolver(capi.Model(m), dt, prob, params, B)
q0, v0, qg = problems.perturbed_batch(m, prob, B)
gs.reset_initial_conditions(q0, v0); gs.set_q(qg)
flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hq0, hv0, hel = pin(q0), pin(v0), pin(np.zeros(B))
oq, ov = torch.empty((B, T + 1, m.nq), dtype=torch.float64).pin_memory(), torch.empty((B, T + 1, m.nv), dtype=torch.float64).pin_memory()
ot, ost = torch.empty((B, T, m.nv), dtype=torch.float64).pin_memory(), torch.empty((B, 1, NUM_STATS), dtype=torch.float64).pin_memory()
outs = dict(q_out=oq.data_ptr(), v_out=ov.data_ptr(), tau_out=ot.data_ptr(), stats_out=ost.data_ptr())
def run(name, fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); print(f"{name:58s} {(time.perf_counter() - t0) / n * 1e3:.3f} ms/step")
def A(): gs.flush_l2(flush.data_ptr(), flush.numel()); gs.invalidate(); gs.resolve_async(1)
def A2(): A(); gs.synchronize()
def Bv(): gs.flush_l2(flush.data_ptr(), flush.numel()); gs.invalidate(); gs.resolve_async(1, **outs); gs.synchronize()
def C(): gs.flush_l2(flush.data_ptr(), flush.numel()); gs.mpc_advance(hel.numpy(), hq0.numpy(), hv0.numpy()); gs.resolve_async(1); gs.synchronize()
def D(): gs.flush_l2(flush.data_ptr(), flush.numel()); gs.mpc_advance(hel.numpy(), hq0.numpy(), hv0.numpy()); gs.resolve_async(1, **outs); gs.synchronize()
def E(): gs.mpc_advance(hel.numpy(), hq0.numpy(), hv0.numpy()); gs.resolve_async(1, **outs); gs.synchronize()
def G(): gs.flush_l2(flush.data_ptr(), flush.numel()); gs.mpc_resolve_async(hel.numpy(), hq0.numpy(), hv0.numpy(), 1, **outs); gs.synchronize()
def F(): gs.flush_l2(flush.data_ptr(), flush.numel()); gs.synchronize()
run("resident, no host sync (flush + invalidate + resolve)", A)
run("  + synchronize every step", A2)
run("  + D2H of the solution", Bv)
run("mpc_advance instead of invalidate, sync, no D2H", C)
run("full e2e step, two calls (advance, re-solve)", D)
run("full e2e step, one call (bench)", G)
run("full e2e step without the L2 flush", E)
run("L2 flush + sync alone", F)
