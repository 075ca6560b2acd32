// MPC shell, device side: what ModelPredictiveController::UpdateAbstractState does between two re-solves
// (examples/mpc_controller.cc:43-98, python_examples/mpc_utils.py:183-217), batched and without leaving HBM:
//   * the previous solution q_0..q_T is interpolated by a C2 cubic spline per coordinate
//     (PiecewisePolynomial::CubicWithContinuousSecondDerivatives = not-a-knot end conditions,
//     mpc_controller.cc:129-137) and the new guess is q_i = spline(elapsed + i dt), i = 1..T, clamped to the
//     spline's domain as PiecewisePolynomial::value does; q_0 = the measured q0 (mpc_controller.cc:56-58);
//   * the nominal trajectory moves with the initial condition for the selected coordinates:
//     q_nom_t += selector o (q0 - q_nom_0)   (mpc_controller.cc:62-69);
//   * ResetInitialConditions(q0, v0) (mpc_controller.cc:72) and every cache entry goes stale.
// One CTA per problem.  The spline's tridiagonal system (uniform knots) is solved by the Thomas recurrence, one
// thread per coordinate on shared-memory columns; loading the old solution, evaluating the spline at the T new
// knots and shifting the nominal trajectory are spread over all threads of the CTA (the first version ran
// everything in one thread per coordinate out of local memory: 16 us for 41 knots, all of it latency).
// fp64 divisions cost ~400 cycles each on this part: the reciprocals of h, h^2 and 6 are formed once, and the
// modified super-diagonal of the recurrence (which depends on the horizon only) comes from a table made at solver
// creation.
#include "solver.h"

namespace idto {

namespace {
constexpr int kMaxKnots = 256;
}

__global__ void __launch_bounds__(256) k_mpc_advance(SolverConsts sc, SolverBufs bf, const double* __restrict__ elapsed,
                                                     const double* __restrict__ q0, const double* __restrict__ v0,
                                                     const double* __restrict__ selector, double* q_init, double* v_init,
                                                     double* q_nom) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int nq = sc.nq, nv = sc.nv, N = sc.T;  // knots 0..N
  const double h = sc.dt;
  double* y = sm;                  // [N+1][nq] the previous solution
  double* M = y + (N + 1) * nq;    // [N+1][nq] second derivatives of its spline
  double* shift = M + (N + 1) * nq;  // [nq]
  double* q = bf.st.q + size_t(b) * (N + 1) * nq;
  for (int e = tid; e < (N + 1) * nq; e += nt) y[e] = q[e];
  __syncthreads();
  // second derivatives M_j of the not-a-knot spline on uniform knots:
  //   M_{j-1} + 4 M_j + M_{j+1} = 6 (y_{j+1} - 2 y_j + y_{j-1}) / h^2,  j = 1..N-1
  //   M_0 = 2 M_1 - M_2,  M_N = 2 M_{N-1} - M_{N-2}   =>   6 M_1 = rhs_1,  6 M_{N-1} = rhs_{N-1}
  const double inv_h = 1.0 / h, six_inv_h2 = 6.0 / (h * h), sixth = 1.0 / 6.0;
  if (tid < nq) {
    const int i = tid;
    auto Y = [&](int j) -> double { return y[j * nq + i]; };
    auto Mr = [&](int j) -> double& { return M[j * nq + i]; };
    auto rhs = [&](int j) { return ((Y(j + 1) - 2.0 * Y(j)) + Y(j - 1)) * six_inv_h2; };
    if (N >= 4) {
      const double M1 = rhs(1) * sixth, mN1 = rhs(N - 1) * sixth;
      Mr(1) = M1;
      // Thomas on j = 2..N-2 with the known neighbours M_1, M_{N-1} moved to the right-hand side
      // (cp: modified super-diagonal, M doubles as the modified right-hand side)
      auto d = [&](int j) { return rhs(j) - (j == 2 ? M1 : 0.0) - (j == N - 2 ? mN1 : 0.0); };
      // the modified super-diagonal cp[j] = 1 / (4 - cp[j-1]) depends on the horizon only: tabulated on the host at
      // solver creation (a chain of N dependent fp64 divisions otherwise)
      const double* __restrict__ cp = bf.spline_cp;
      double prev = d(2) * 0.25;
      Mr(2) = prev;
      for (int j = 3; j <= N - 2; ++j) {
        prev = (d(j) - prev) * cp[j];
        Mr(j) = prev;
      }
      Mr(N - 1) = mN1;
      double next = mN1;
      for (int j = N - 3; j >= 2; --j) {
        next = Mr(j) - cp[j] * (j == N - 3 ? Mr(N - 2) : next);
        Mr(j) = next;
      }
      Mr(0) = 2.0 * Mr(1) - Mr(2);
      Mr(N) = 2.0 * Mr(N - 1) - Mr(N - 2);
    } else if (N == 3) {  // four points: the not-a-knot spline is the cubic through them
      // M linear over the whole range: M_0 = 2 M_1 - M_2, M_3 = 2 M_2 - M_1, and the two interior equations
      //   (2M_1 - M_2) + 4 M_1 + M_2 = 6 M_1;  M_1 + 4 M_2 + (2 M_2 - M_1) = 6 M_2
      Mr(1) = rhs(1) * sixth, Mr(2) = rhs(2) * sixth;
      Mr(0) = 2.0 * Mr(1) - Mr(2), Mr(3) = 2.0 * Mr(2) - Mr(1);
    } else if (N == 2) {  // three points: parabola
      Mr(0) = Mr(1) = Mr(2) = rhs(1) * sixth;
    } else {  // two points: line
      Mr(0) = Mr(1) = 0.0;
    }
    const double qi0 = q0[size_t(b) * nq + i];
    if (selector) shift[i] = selector[i] * (qi0 - q_nom[size_t(b) * (N + 1) * nq + i]);
    q_init[size_t(b) * nq + i] = qi0;
  }
  __syncthreads();
  const double tau0 = elapsed[b], tend = N * h;
  for (int e = tid; e < (N + 1) * nq; e += nt) {
    const int j = e / nq, i = e - j * nq;
    if (j == 0) {
      q[e] = q0[size_t(b) * nq + i];  // the measured state (mpc_controller.cc:56-58)
      continue;
    }
    double tq = tau0 + j * h;
    tq = fmin(fmax(tq, 0.0), tend);
    int k = int(tq * inv_h);  // (the spline is C2: a knot attributed to the neighbouring piece changes nothing)
    k = k > N - 1 ? N - 1 : k;
    const double s = tq - k * h;
    const double yk = y[k * nq + i], yk1 = y[(k + 1) * nq + i], Mk = M[k * nq + i], Mk1 = M[(k + 1) * nq + i];
    const double bk = (yk1 - yk) * inv_h - h * (2.0 * Mk + Mk1) * sixth;
    q[e] = yk + s * (bk + s * (0.5 * Mk + s * ((Mk1 - Mk) * (sixth * inv_h))));
  }
  if (selector) {
    double* qn = q_nom + size_t(b) * (N + 1) * nq;
    for (int e = tid; e < (N + 1) * nq; e += nt) qn[e] += shift[e % nq];
  }
  for (int e = tid; e < nv; e += nt) v_init[size_t(b) * nv + e] = v0[size_t(b) * nv + e];
  if (tid == 0) {
    ProbCtl* ctl = bf.ctl + b;
    ctl->traj_dirty = 1, ctl->derivs_dirty = 1, ctl->pending = 0;  // state.h:333-350
  }
}

int launch_mpc_advance(const SolverConsts& sc, const SolverBufs& bf, const double* elapsed, const double* q0,
                       const double* v0, const double* selector, double* q_init, double* v_init, double* q_nom,
                       cudaStream_t stream) {
  if (sc.T + 1 > kMaxKnots) return IDTO_ERR_UNSUPPORTED;
  const size_t smem = (size_t(2) * (sc.T + 1) * sc.nq + sc.nq) * sizeof(double);
  static bool attr_set[kMaxDevices] = {};
  if (first_use_on_device(attr_set))
    cudaFuncSetAttribute(k_mpc_advance, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (smem > 200 * 1024) return IDTO_ERR_UNSUPPORTED;
  g_launch_counter += 1;
  k_mpc_advance<<<sc.B, 256, smem, stream>>>(sc, bf, elapsed, q0, v0, selector, q_init, v_init, q_nom);
  return IDTO_OK;
}

}  // namespace idto
