"""ID partials with the path columns (single-lane subtree evaluations) against the full-evaluation kernel
(IDTO_PATH_COLS=0), same inputs: python tools/path_check.py [model] [method]"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
name = sys.argv[1] if len(sys.argv) > 1 else "mini_cheetah"
method = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if os.environ.get("PATH_CHECK_CHILD"):
    from idto_b200 import capi, problems
    m, dt, prob, params, guess = getattr(problems, name)(gradients_method=method)
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
    rng = np.random.default_rng(7)
    q = np.array(guess, float)[None].repeat(2, 0)
    q[:, 1:] += rng.normal(0, 0.03, q[:, 1:].shape)
    gs.set_q(q)
    gs.eval(1)
    np.savez(os.environ["PATH_CHECK_CHILD"], **{f: gs.get(f) for f in ("dtau_dqm", "dtau_dqt", "dtau_dqp", "tau")})
    sys.exit(0)
out = {}
for tag, env in (("full", "0"), ("path", "1")):
    f = f"/tmp/path_check_{tag}.npz"
    subprocess.run([sys.executable, __file__, name, str(method)], check=True,
                   env=dict(os.environ, PATH_CHECK_CHILD=f, IDTO_PATH_COLS=env))
    out[tag] = np.load(f)
for k in ("dtau_dqm", "dtau_dqt", "dtau_dqp"):
    a, b = out["full"][k], out["path"][k]
    mask = ~(np.isnan(a) & np.isnan(b))
    d = np.abs(a - b)[mask]
    print(f"{name} method {method} {k}: max|full| {np.nanmax(np.abs(a)):.3e}  max|diff| {d.max():.3e}  "
          f"identical entries {np.mean(a[mask] == b[mask]) * 100:.2f}%  nan mismatch {np.sum(np.isnan(a) != np.isnan(b))}")
