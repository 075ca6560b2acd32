// MPC shell, device side: what ModelPredictiveController::UpdateAbstractState does between two re-solves
// (examples/mpc_controller.cc:43-98, python_examples/mpc_utils.py:183-217), batched and without leaving HBM:
//   * the previous solution q_0..q_T is interpolated by a C2 cubic spline per coordinate
//     (PiecewisePolynomial::CubicWithContinuousSecondDerivatives = not-a-knot end conditions,
//     mpc_controller.cc:129-137) and the new guess is q_i = spline(elapsed + i dt), i = 1..T, clamped to the
//     spline's domain as PiecewisePolynomial::value does; q_0 = the measured q0 (mpc_controller.cc:56-58);
//   * the nominal trajectory moves with the initial condition for the selected coordinates:
//     q_nom_t += selector o (q0 - q_nom_0)   (mpc_controller.cc:62-69);
//   * ResetInitialConditions(q0, v0) (mpc_controller.cc:72) and every cache entry goes stale.
// One thread per (problem, coordinate); the spline's tridiagonal system (uniform knots) is solved by the
// Thomas recurrence in local memory.  fp64 divisions cost ~400 cycles each on this part: the kernel spent 30 us in
// ~200 of them per thread; the reciprocals of h, h^2 and 6 are formed once, and the modified super-diagonal of the
// recurrence (which depends on the horizon only) comes from a table made at solver creation.
#include "solver.h"

namespace idto {

namespace {
constexpr int kMaxKnots = 256;
}

__global__ void __launch_bounds__(64) k_mpc_advance(SolverConsts sc, SolverBufs bf, const double* __restrict__ elapsed,
                                                    const double* __restrict__ q0, const double* __restrict__ v0,
                                                    const double* __restrict__ selector, double* q_init, double* v_init,
                                                    double* q_nom) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nq = sc.nq, nv = sc.nv, N = sc.T;  // knots 0..N
  const bool live = idx < sc.B * nq;
  const int b = live ? idx / nq : 0, i = live ? idx % nq : 0;
  const double h = sc.dt;
  double* q = bf.st.q + size_t(b) * (N + 1) * nq + i;
  double y[kMaxKnots], M[kMaxKnots];
  for (int j = 0; j <= N; ++j) y[j] = q[size_t(j) * nq];
  // second derivatives M_j of the not-a-knot spline on uniform knots:
  //   M_{j-1} + 4 M_j + M_{j+1} = 6 (y_{j+1} - 2 y_j + y_{j-1}) / h^2,  j = 1..N-1
  //   M_0 = 2 M_1 - M_2,  M_N = 2 M_{N-1} - M_{N-2}   =>   6 M_1 = rhs_1,  6 M_{N-1} = rhs_{N-1}
  const double inv_h = 1.0 / h, six_inv_h2 = 6.0 / (h * h), sixth = 1.0 / 6.0;
  auto rhs = [&](int j) { return ((y[j + 1] - 2.0 * y[j]) + y[j - 1]) * six_inv_h2; };
  if (N >= 4) {
    M[1] = rhs(1) * sixth;
    M[N - 1] = rhs(N - 1) * sixth;
    // Thomas on j = 2..N-2 with the known neighbours M_1, M_{N-1} moved to the right-hand side
    // (cp: modified super-diagonal, M doubles as the modified right-hand side)
    auto d = [&](int j) { return rhs(j) - (j == 2 ? M[1] : 0.0) - (j == N - 2 ? M[N - 1] : 0.0); };
    const double mN1 = M[N - 1];
    // the modified super-diagonal cp[j] = 1 / (4 - cp[j-1]) depends on the horizon only: tabulated on the host at
    // solver creation (a chain of N dependent fp64 divisions otherwise)
    const double* __restrict__ cp = bf.spline_cp;
    M[2] = d(2) * 0.25;
    for (int j = 3; j <= N - 2; ++j) M[j] = (d(j) - M[j - 1]) * cp[j];
    M[N - 1] = mN1;
    for (int j = N - 3; j >= 2; --j) M[j] -= cp[j] * M[j + 1];
    M[0] = 2.0 * M[1] - M[2];
    M[N] = 2.0 * M[N - 1] - M[N - 2];
  } else if (N == 3) {  // four points: the not-a-knot spline is the cubic through them
    // M linear over the whole range: M_0 = 2 M_1 - M_2, M_3 = 2 M_2 - M_1, and the two interior equations
    //   6 M_1 = rhs_1 ... with M_2 unknown too: (2M_1 - M_2) + 4 M_1 + M_2 = 6 M_1;  M_1 + 4 M_2 + (2 M_2 - M_1) = 6 M_2
    M[1] = rhs(1) * sixth, M[2] = rhs(2) * sixth;
    M[0] = 2.0 * M[1] - M[2], M[3] = 2.0 * M[2] - M[1];
  } else if (N == 2) {  // three points: parabola
    M[0] = M[1] = M[2] = rhs(1) * sixth;
  } else {  // two points: line
    M[0] = M[1] = 0.0;
  }
  if (!live) return;  // (padding threads ran the recurrence of item (0, 0): harmless)
  const double tau0 = elapsed[b], tend = N * h;
  for (int j = 1; j <= N; ++j) {
    double tq = tau0 + j * h;
    tq = fmin(fmax(tq, 0.0), tend);
    int k = int(tq * inv_h);  // (the spline is C2: a knot attributed to the neighbouring piece changes nothing)
    k = k > N - 1 ? N - 1 : k;
    const double s = tq - k * h;
    const double bk = (y[k + 1] - y[k]) * inv_h - h * (2.0 * M[k] + M[k + 1]) * sixth;
    q[size_t(j) * nq] = y[k] + s * (bk + s * (0.5 * M[k] + s * ((M[k + 1] - M[k]) * (sixth * inv_h))));
  }
  const double qi0 = q0[size_t(b) * nq + i];
  q[0] = qi0;
  q_init[size_t(b) * nq + i] = qi0;
  if (selector) {
    double* qn = q_nom + size_t(b) * (N + 1) * nq + i;
    const double shift = selector[i] * (qi0 - qn[0]);
    for (int j = 0; j <= N; ++j) qn[size_t(j) * nq] += shift;
  }
  if (i < nv) v_init[size_t(b) * nv + i] = v0[size_t(b) * nv + i];
  if (i == 0) {
    for (int j = nq; j < nv; ++j) v_init[size_t(b) * nv + j] = v0[size_t(b) * nv + j];  // nv > nq never happens
    ProbCtl* ctl = bf.ctl + b;
    ctl->traj_dirty = 1, ctl->derivs_dirty = 1, ctl->pending = 0;  // state.h:333-350
  }
}

int launch_mpc_advance(const SolverConsts& sc, const SolverBufs& bf, const double* elapsed, const double* q0,
                       const double* v0, const double* selector, double* q_init, double* v_init, double* q_nom,
                       cudaStream_t stream) {
  if (sc.T + 1 > kMaxKnots) return IDTO_ERR_UNSUPPORTED;
  g_launch_counter += 1;
  k_mpc_advance<<<(sc.B * sc.nq + 63) / 64, 64, 0, stream>>>(sc, bf, elapsed, q0, v0, selector, q_init, v_init, q_nom);
  return IDTO_OK;
}

}  // namespace idto
