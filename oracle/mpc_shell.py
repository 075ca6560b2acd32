"""CPU restatement of the reference's MPC shell (TEST INFRASTRUCTURE ONLY, like everything under oracle/).

Follows ModelPredictiveController::UpdateAbstractState / UpdateInitialGuess / StoreOptimizerSolution
(examples/mpc_controller.cc:43-137) and its Python twin (python_examples/mpc_utils.py:150-217):
the stored solution is a C2 cubic interpolant of the knots (Drake's
PiecewisePolynomial::CubicWithContinuousSecondDerivatives = "not-a-knot" end conditions), the new guess
is that interpolant sampled at elapsed + i*dt (PiecewisePolynomial::value clamps outside its domain),
q_guess[0] = the measured q0, and q_nom is shifted by selector o (q0 - q_nom[0]).

Drake is not installed here, so the spline is pinned against scipy.interpolate.CubicSpline(bc_type=
"not-a-knot") in tests/test_mpc_shell.py instead of against PiecewisePolynomial itself ("parity
unpinned" at the Drake boundary; the definition of the interpolant is the published one).
"""
from __future__ import annotations

import numpy as np


def not_a_knot_second_derivatives(y: np.ndarray, h: float) -> np.ndarray:
    """M_j = S''(x_j) of the not-a-knot cubic spline through y[j] at x_j = j*h (columns independent)."""
    y = np.asarray(y, float)
    N = y.shape[0] - 1
    M = np.zeros_like(y)
    if N < 2:
        return M
    rhs = np.zeros_like(y)
    rhs[1:N] = 6.0 * (y[2:] - 2.0 * y[1:N] + y[:N - 1]) / (h * h)
    if N == 2:
        M[:] = rhs[1] / 6.0
        return M
    # not-a-knot: M_0 = 2 M_1 - M_2 and M_N = 2 M_{N-1} - M_{N-2} turn rows 1 and N-1 into 6 M_j = rhs_j
    M[1], M[N - 1] = rhs[1] / 6.0, rhs[N - 1] / 6.0
    if N >= 4:
        n = N - 3  # unknowns M_2..M_{N-2}
        A = np.zeros((n, n))
        d = rhs[2:N - 1].copy()
        for k in range(n):
            A[k, k] = 4.0
            if k > 0:
                A[k, k - 1] = 1.0
            if k < n - 1:
                A[k, k + 1] = 1.0
        d[0] -= M[1]
        d[-1] -= M[N - 1]
        M[2:N - 1] = np.linalg.solve(A, d.reshape(n, -1)).reshape(d.shape)
    M[0] = 2.0 * M[1] - M[2]
    M[N] = 2.0 * M[N - 1] - M[N - 2]
    return M


def spline_value(y: np.ndarray, M: np.ndarray, h: float, t: float) -> np.ndarray:
    N = y.shape[0] - 1
    t = min(max(t, 0.0), N * h)
    k = min(int(t / h), N - 1)
    s = t - k * h
    b = (y[k + 1] - y[k]) / h - h * (2.0 * M[k] + M[k + 1]) / 6.0
    return y[k] + s * (b + s * (0.5 * M[k] + s * ((M[k + 1] - M[k]) / (6.0 * h))))


def shifted_guess(q_sol: np.ndarray, dt: float, elapsed: float, q0: np.ndarray) -> np.ndarray:
    """New q_guess [T+1][nq] from the previous solution (mpc_controller.cc:100-110, 56-58)."""
    q_sol = np.asarray(q_sol, float)
    M = not_a_knot_second_derivatives(q_sol, dt)
    out = np.empty_like(q_sol)
    out[0] = q0
    for i in range(1, q_sol.shape[0]):
        out[i] = spline_value(q_sol, M, dt, elapsed + i * dt)
    return out


def shifted_nominal(q_nom: np.ndarray, q0: np.ndarray, selector: np.ndarray) -> np.ndarray:
    """q_nom_t + selector o (q0 - q_nom_0)   (mpc_controller.cc:62-69)."""
    q_nom = np.asarray(q_nom, float)
    return q_nom + np.asarray(selector, float) * (np.asarray(q0, float) - q_nom[0])
