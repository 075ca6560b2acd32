"""Resident Gauss-Newton iteration rate and per-stage times of a parity-test configuration (not the bench line):
    python tools/config_rate.py allegro_hand --T 60 --batch 64
One iteration per step from a stale cache (like bench.py's resident arm), CUDA-event stage times from the
library's own profiler."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from idto_b200 import capi, problems  # noqa: E402
from idto_b200.types import GRAD_CENTRAL, GRAD_FORWARD  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name")
    ap.add_argument("--T", type=int, default=None)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--method", default="forward", choices=["forward", "central"])
    args = ap.parse_args()
    kw = {"gradients_method": GRAD_FORWARD if args.method == "forward" else GRAD_CENTRAL}
    if args.T:
        kw["T"] = args.T
    m, dt, prob, params, guess = getattr(problems, args.name)(**kw)
    B = args.batch
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, B)
    q0, v0, qg = problems.perturbed_batch(m, prob, B) if args.name == "mini_cheetah" else (None, None, None)
    gs.set_q(np.stack([np.array(guess)] * B) if qg is None else qg)
    it, _, stats = gs.solve(3)  # move off the initial guess

    def step():
        gs.invalidate()
        gs.resolve_async(1)

    for _ in range(3):
        step()
    gs.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    gs.synchronize()
    el = time.perf_counter() - t0
    gs.profile_enable(True)
    for _ in range(5):
        step()
    gs.synchronize()
    stage = {}
    for name in ("id_partials", "trajectory", "assemble", "factor", "lagrange", "dogleg", "trajectory_scratch",
                 "trust_update"):
        tot, n = gs.profile_read(name)
        stage[name] = round(tot / max(n, 1), 4)
    print(json.dumps({"config": args.name, "T": prob.num_steps, "nq": m.nq, "pairs": int(m.npairs), "batch": B,
                      "method": args.method, "iters_per_s": B * args.steps / el, "ms_per_step": el / args.steps * 1e3,
                      "cost_after_3": float(stats[0, -1, 0]), "stage_ms": stage}))


if __name__ == "__main__":
    main()
