"""ctypes loader for the CPU oracle (oracle/idto_oracle.cc).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under idto_b200/ imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from idto_b200.bake import ModelDesc, model_desc  # noqa: E402
from idto_b200.types import NUM_STATS, Params, ProblemDesc  # noqa: E402

_LIB = None
_D = ctypes.POINTER(ctypes.c_double)
_I = ctypes.POINTER(ctypes.c_int)


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "build", "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.oracle_create.restype = ctypes.c_void_p
        L.oracle_create.argtypes = [ctypes.POINTER(ModelDesc), ctypes.POINTER(ProblemDesc), ctypes.POINTER(Params)]
        L.oracle_destroy.argtypes = [ctypes.c_void_p]
        L.oracle_set_num_threads.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oracle_num_unactuated.argtypes = [ctypes.c_void_p]
        L.oracle_unactuated_dofs.argtypes = [ctypes.c_void_p, _I]
        L.oracle_set_q.argtypes = [ctypes.c_void_p, _D]
        L.oracle_reset_initial_conditions.argtypes = [ctypes.c_void_p, _D, _D]
        L.oracle_update_nominal_trajectory.argtypes = [ctypes.c_void_p, _D, _D]
        L.oracle_set_delta.argtypes = [ctypes.c_void_p, ctypes.c_double]
        L.oracle_get_delta.argtypes = [ctypes.c_void_p]
        L.oracle_get_delta.restype = ctypes.c_double
        L.oracle_eval.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oracle_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, _D]
        L.oracle_get.restype = ctypes.c_long
        L.oracle_solve.argtypes = [ctypes.c_void_p, ctypes.c_int, _I, _D]
        L.oracle_solution.argtypes = [ctypes.c_void_p, _D, _D, _D]
        L.oracle_inverse_dynamics.argtypes = [ctypes.c_void_p, _D, _D, _D, _D, _I]
        L.oracle_mass_matrix.argtypes = [ctypes.c_void_p, _D, _D]
        L.oracle_body_poses.argtypes = [ctypes.c_void_p, _D, _D, _D]
        L.oracle_penta_multiply.argtypes = [ctypes.c_int, ctypes.c_int, _D, _D, _D, _D, _D]
        L.oracle_penta_solve.argtypes = [ctypes.c_int, ctypes.c_int, _D, _D, _D, _D, ctypes.c_int]
        L.oracle_penta_dense.argtypes = [ctypes.c_int, ctypes.c_int, _D, _D, _D, _D]
        L.oracle_penta_scale.argtypes = [ctypes.c_int, ctypes.c_int, _D, _D, _D, _D]
        L.oracle_point_distance.argtypes = [ctypes.c_int, _D, _D, _D, _D, _D]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(_D)


def point_distance(gtype, dims, R_WG, p_WG, p_WQ):
    """(distance, p_GN, grad_W) of the restated Drake closed forms (sphere, box, capsule, cylinder)."""
    a = [np.ascontiguousarray(np.asarray(x, float).reshape(-1)) for x in (dims, R_WG, p_WG, p_WQ)]
    out = np.zeros(7)
    lib().oracle_point_distance(int(gtype), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(out))
    return out[0], out[1:4].copy(), out[4:7].copy()


class Oracle:
    """One TrajectoryOptimizer + one WarmStart on the CPU."""

    def __init__(self, model, time_step, prob, params, num_threads=1):
        self.model, self.prob, self.params = model, prob, params
        self.T, self.nq, self.nv = prob.num_steps, model.nq, model.nv
        md, self._k1 = model_desc(model)
        pd, self._k2 = prob.to_c(time_step, model.nq, model.nv)
        pc = params.to_c()
        self.h = lib().oracle_create(ctypes.byref(md), ctypes.byref(pd), ctypes.byref(pc))
        if not self.h:
            raise RuntimeError("oracle_create failed")
        lib().oracle_set_num_threads(self.h, int(num_threads))
        self.nu = lib().oracle_num_unactuated(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_destroy(self.h)
            self.h = None

    def unactuated_dofs(self):
        out = np.zeros(max(self.nu, 1), np.int32)
        lib().oracle_unactuated_dofs(self.h, out.ctypes.data_as(_I))
        return out[:self.nu].tolist()

    def set_q(self, q):
        q = np.ascontiguousarray(np.asarray(q, float).reshape(self.T + 1, self.nq))
        lib().oracle_set_q(self.h, _p(q))

    def reset_initial_conditions(self, q0, v0):
        q0, v0 = np.ascontiguousarray(q0, float), np.ascontiguousarray(v0, float)
        lib().oracle_reset_initial_conditions(self.h, _p(q0), _p(v0))

    def update_nominal_trajectory(self, qn, vn):
        qn = np.ascontiguousarray(np.asarray(qn, float).reshape(self.T + 1, self.nq))
        vn = np.ascontiguousarray(np.asarray(vn, float).reshape(self.T + 1, self.nv))
        lib().oracle_update_nominal_trajectory(self.h, _p(qn), _p(vn))

    def set_delta(self, d):
        lib().oracle_set_delta(self.h, float(d))

    def get_delta(self):
        return lib().oracle_get_delta(self.h)

    def eval(self, stage=4):
        lib().oracle_eval(self.h, int(stage))

    def get(self, name):
        n = lib().oracle_get(self.h, name.encode(), None)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n)
        lib().oracle_get(self.h, name.encode(), _p(out))
        return out

    def solve(self, max_iterations):
        stats = np.zeros((max_iterations, NUM_STATS))
        reason = ctypes.c_int(0)
        k = lib().oracle_solve(self.h, int(max_iterations), ctypes.byref(reason), _p(stats))
        return k, reason.value, stats[:k]

    def solution(self):
        q = np.zeros((self.T + 1, self.nq))
        v = np.zeros((self.T + 1, self.nv))
        tau = np.zeros((self.T, self.nv))
        lib().oracle_solution(self.h, _p(q), _p(v), _p(tau))
        return q, v, tau

    def inverse_dynamics(self, q, v, a):
        q, v, a = (np.ascontiguousarray(x, float) for x in (q, v, a))
        tau = np.zeros(self.nv)
        act = np.zeros(max(self.model.npairs, 1), np.int32)
        lib().oracle_inverse_dynamics(self.h, _p(q), _p(v), _p(a), _p(tau), act.ctypes.data_as(_I))
        return tau, act[:self.model.npairs]

    def mass_matrix(self, q):
        q = np.ascontiguousarray(q, float)
        M = np.zeros((self.nv, self.nv))
        lib().oracle_mass_matrix(self.h, _p(q), _p(M))
        return M.T.copy()  # column-major -> (row, col)

    def body_poses(self, q):
        q = np.ascontiguousarray(q, float)
        R = np.zeros((self.model.nbodies, 3, 3))
        p = np.zeros((self.model.nbodies, 3))
        lib().oracle_body_poses(self.h, _p(q), _p(R), _p(p))
        return R, p
