"""Generates the committed fixtures of tests/golden/ (run in the build container; needs /root/reference):

  reference_anchors.json   the numeric known answers the reference's own tests hold for this path, extracted
                           from the reference test sources (file:line recorded per anchor)
  oracle_<case>.npz        inputs and outputs of the CPU oracle on seeded states (q, v, a, tau, cost, the three
                           ID-partial bands, g, lambda, dq) — regression anchors for the oracle itself and a
                           second, box-independent target for the GPU parity tests

    python tests/golden/make_golden.py
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def reference_anchors():
    out = {}
    src = open(os.path.join(REF, "optimizer/test/trajectory_optimizer_test.cc")).read().splitlines()
    # CalcCostFromState: the 11 hard-coded positions (trajectory_optimizer_test.cc:1189-1199)
    q, first = [], None
    for n, line in enumerate(src, 1):
        mm = re.search(r"q\.push_back\(drake::Vector1d\(([0-9.]+)\)\);", line)
        if mm and 1150 < n < 1260:
            q.append(float(mm.group(1)))
            first = first or n
    out["cost_from_state_q"] = {"values": q, "source": f"optimizer/test/trajectory_optimizer_test.cc:{first}-{first + len(q) - 1}"}
    # pendulum constants of the closed-form checks (m, l, b)
    for n, line in enumerate(src, 1):
        mm = re.search(r"const double (m|l|b) = ([0-9.]+);", line)
        if mm and 1210 < n < 1230:
            out.setdefault("pendulum_constants", {"source": "optimizer/test/trajectory_optimizer_test.cc:1219-1221"})[mm.group(1)] = float(mm.group(2))
    py = open(os.path.join(REF, "python_bindings/test/trajectory_optimizer_test.py")).read().splitlines()
    for n, line in enumerate(py, 1):
        mm = re.search(r"np\.array\(\[([0-9., ]+)\]\)", line)
        if mm and "expected" in line.lower() or (mm and 80 <= n <= 86):
            out["spinner_final_q"] = {"values": [float(x) for x in mm.group(1).split(",")], "tolerance": 1e-3,
                                      "source": f"python_bindings/test/trajectory_optimizer_test.py:{n}"}
    return out


def oracle_case(name, T, seed, amp):
    from idto_b200 import problems
    from idto_b200.types import GRAD_CENTRAL
    from oracle import oracle
    m, dt, prob, params, guess = getattr(problems, name)(T=T, gradients_method=GRAD_CENTRAL)
    oc = oracle.Oracle(m, dt, prob, params)
    rng = np.random.default_rng(seed)
    q = np.array(guess, float)
    q[1:] += rng.normal(0, amp, q[1:].shape)
    oc.set_q(q)
    oc.eval(4)
    d = {"q": q, "T": T, "seed": seed}
    for f in ("v", "a", "tau", "cost", "h", "dtau_dqm", "dtau_dqt", "dtau_dqp", "g", "D", "lambda", "merit", "dq"):
        d[f] = oc.get(f)
    return d


if __name__ == "__main__":
    with open(os.path.join(HERE, "reference_anchors.json"), "w") as f:
        json.dump(reference_anchors(), f, indent=1)
    for name, T, seed, amp in (("hopper", 10, 21, 0.05), ("mini_cheetah", 6, 22, 0.03), ("spinner", 8, 23, 0.05),
                               ("allegro_hand", 4, 24, 0.01)):
        np.savez_compressed(os.path.join(HERE, f"oracle_{name}.npz"), **oracle_case(name, T, seed, amp))
    print(open(os.path.join(HERE, "reference_anchors.json")).read()[:600])
    print({f: os.path.getsize(os.path.join(HERE, f)) for f in sorted(os.listdir(HERE))})
