"""MPC shell around the re-solve: the classes of python_examples/mpc_utils.py without the Drake systems
framework (a LeafSystem's periodic event becomes an explicit `UpdateAbstractState(time, x0)` call).

  StoredTrajectory            mpc_utils.py:13-21   start_time + C2 cubic interpolants of q, v, tau
  Interpolator                mpc_utils.py:24-84   actuated state / input reference at time t for a PD controller
  ModelPredictiveController   mpc_utils.py:87-217  re-solve from the measured state, warm-started by the shifted
                                                   previous solution

The interpolant is Drake's PiecewisePolynomial.CubicWithContinuousSecondDerivatives (not-a-knot end
conditions, values clamped outside the knots).  The guess shift of a re-plan runs on the device
(`idto_mpc_advance`); the host-side interpolant below only serves the low-level controller, which samples the
stored trajectory between re-plans.
"""
from __future__ import annotations

import numpy as np

from .types import TrajectoryOptimizerSolution, TrajectoryOptimizerStats


class CubicSpline:
    """C2 cubic through uniformly spaced knots y[j] at j*h, not-a-knot end conditions, clamped evaluation."""

    def __init__(self, h: float, y):
        self.h, self.y = float(h), np.asarray(y, float)
        N = self.y.shape[0] - 1
        M = np.zeros_like(self.y)
        if N >= 2:
            rhs = np.zeros_like(self.y)
            rhs[1:N] = 6.0 * (self.y[2:] - 2.0 * self.y[1:N] + self.y[:N - 1]) / (h * h)
            if N == 2:
                M[:] = rhs[1] / 6.0
            else:
                M[1], M[N - 1] = rhs[1] / 6.0, rhs[N - 1] / 6.0
                if N >= 4:  # Thomas recurrence on M_2..M_{N-2} (diagonal 4, off-diagonals 1)
                    n = N - 3
                    d = rhs[2:N - 1].copy()
                    d[0] -= M[1]
                    d[-1] -= M[N - 1]
                    cp = np.zeros(n)
                    cp[0] = 0.25
                    d[0] = d[0] * 0.25
                    for k in range(1, n):
                        cp[k] = 1.0 / (4.0 - cp[k - 1])
                        d[k] = (d[k] - d[k - 1]) * cp[k]
                    for k in range(n - 2, -1, -1):
                        d[k] -= cp[k] * d[k + 1]
                    M[2:N - 1] = d
                M[0] = 2.0 * M[1] - M[2]
                M[N] = 2.0 * M[N - 1] - M[N - 2]
        self.M = M

    def value(self, t: float):
        N = self.y.shape[0] - 1
        t = min(max(float(t), 0.0), N * self.h)
        k = min(int(t / self.h), N - 1)
        s = t - k * self.h
        y, M, h = self.y, self.M, self.h
        b = (y[k + 1] - y[k]) / h - h * (2.0 * M[k] + M[k + 1]) / 6.0
        return y[k] + s * (b + s * (0.5 * M[k] + s * ((M[k + 1] - M[k]) / (6.0 * h))))


class StoredTrajectory:
    """mpc_utils.py:13-21."""
    start_time = None
    q = None
    v = None
    tau = None


class Interpolator:
    """mpc_utils.py:24-84: x(t) = [Bq q(t); Bv v(t)], u(t) = Bv tau(t) of a StoredTrajectory."""

    def __init__(self, Bq, Bv):
        self.Bq, self.Bv = np.asarray(Bq, float), np.asarray(Bv, float)
        assert self.Bq.shape[0] == self.Bv.shape[0]

    def SendState(self, time, trajectory: StoredTrajectory):
        t = time - trajectory.start_time
        return np.concatenate((self.Bq @ trajectory.q.value(t), self.Bv @ trajectory.v.value(t)))

    def SendControl(self, time, trajectory: StoredTrajectory):
        return self.Bv @ trajectory.tau.value(time - trajectory.start_time)


class ModelPredictiveController:
    """mpc_utils.py:87-217.  `optimizer` is an idto_b200.pyidto.TrajectoryOptimizer."""

    def __init__(self, optimizer, q_guess, nq, nv, mpc_rate, q_nom_relative_to_q_init=None):
        self.optimizer, self.nq, self.nv = optimizer, int(nq), int(nv)
        self.q_guess = [np.array(q, float) for q in q_guess]
        self.warm_start = optimizer.CreateWarmStart(self.q_guess)
        self.time_step, self.num_steps = optimizer.time_step(), optimizer.num_steps()
        self.mpc_rate = float(mpc_rate)
        self.selector = None if q_nom_relative_to_q_init is None else np.asarray(q_nom_relative_to_q_init, float)
        solution, stats = TrajectoryOptimizerSolution(), TrajectoryOptimizerStats()
        optimizer.SolveFromWarmStart(self.warm_start, solution, stats)  # mpc_utils.py:126-128
        self.stored_trajectory = self.StoreOptimizerSolution(solution, 0.0)

    def StoreOptimizerSolution(self, solution, start_time) -> StoredTrajectory:
        """mpc_utils.py:150-180: knots every time_step; the last control input is repeated."""
        tr = StoredTrajectory()
        tr.start_time = float(start_time)
        tau = list(solution.tau) + [solution.tau[-1]]
        tr.q = CubicSpline(self.time_step, np.asarray(solution.q, float))
        tr.v = CubicSpline(self.time_step, np.asarray(solution.v, float))
        tr.tau = CubicSpline(self.time_step, np.asarray(tau, float))
        return tr

    def UpdateNominalTrajectory(self, time, q0):
        """mpc_utils.py:211-217: a hook; the C++ controller's q_nom_relative_to_q_init shift
        (examples/mpc_controller.cc:62-69) is applied on the device when a selector was given."""

    def UpdateAbstractState(self, time, x0) -> StoredTrajectory:
        """mpc_utils.py:183-209: re-solve from the measured state x0 = [q0; v0] at `time`."""
        x0 = np.asarray(x0, float)
        q0, v0 = x0[:self.nq], x0[self.nq:]
        elapsed = time - self.stored_trajectory.start_time
        # ResetInitialConditions + shifted guess (+ nominal shift) without a host round trip of the trajectory
        self.warm_start._s.mpc_advance(elapsed, q0[None], v0[None], self.selector)
        self.optimizer._prob.q_init, self.optimizer._prob.v_init = q0.copy(), v0.copy()
        self.UpdateNominalTrajectory(time, q0)
        solution, stats = TrajectoryOptimizerSolution(), TrajectoryOptimizerStats()
        self.optimizer.SolveFromWarmStart(self.warm_start, solution, stats)
        self.stored_trajectory = self.StoreOptimizerSolution(solution, time)
        self.last_stats = stats
        return self.stored_trajectory
