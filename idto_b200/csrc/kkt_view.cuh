// Time-major KKT matrix of one problem, read straight from the scaled Hessian / Jacobian bands
// (see kernels_solve.cu for the ordering of the unknowns).
#pragma once
#include "solver.h"

namespace idto {

struct KktView {
  const double *SA, *SB, *SC;  // scaled Hessian lower bands of problem b: [T+1][nq*nq] column-major
  const double *Jm, *Jt, *Jp;  // scaled Jacobian bands of problem b: [T][nu*nq] row-major (u, c)
  int nq, nu, T, eq;
};

// Entry (r, c) of the lower-band blocks of the time-major KKT matrix (see file header).
__device__ __forceinline__ double kkt_C(const KktView& V, int i, int r, int c) {
  const int nq = V.nq;
  if (r < nq && c < nq) return V.SC[size_t(i) * nq * nq + c * nq + r];
  if (r >= nq && c >= nq) return (i == 0 && r == c) ? 1.0 : 0.0;  // dummy lambda_{-1}
  if (i == 0) return 0.0;
  const int u = (r >= nq ? r : c) - nq, cc = (r >= nq ? c : r);
  return V.Jp[(size_t(i - 1) * V.nu + u) * nq + cc];
}
__device__ __forceinline__ double kkt_B(const KktView& V, int i, int r, int c) {  // block (i, i-1), i >= 1
  const int nq = V.nq;
  if (c >= nq) return 0.0;
  if (r < nq) return V.SB[size_t(i) * nq * nq + c * nq + r];
  return V.Jt[(size_t(i - 1) * V.nu + (r - nq)) * nq + c];
}
__device__ __forceinline__ double kkt_A(const KktView& V, int i, int r, int c) {  // block (i, i-2), i >= 2
  const int nq = V.nq;
  if (c >= nq) return 0.0;
  if (r < nq) return V.SA[size_t(i) * nq * nq + c * nq + r];
  return V.Jm[(size_t(i - 1) * V.nu + (r - nq)) * nq + c];
}


// Generic block (i, j), |i - j| <= 2, of the symmetric KKT matrix.
__device__ __forceinline__ double kkt_blk(const KktView& V, int i, int j, int r, int c) {
  if (j == i) return kkt_C(V, i, r, c);
  if (j == i - 1) return kkt_B(V, i, r, c);
  if (j == i - 2) return kkt_A(V, i, r, c);
  if (j == i + 1) return kkt_B(V, i + 1, c, r);
  return kkt_A(V, i + 2, c, r);  // j == i + 2
}

__device__ __forceinline__ KktView make_kkt_view(const SolverConsts& sc, const SolverBufs& bf, int b) {
  const int nblk = sc.T + 1, nq = sc.nq;
  KktView V;
  V.SA = bf.SA + size_t(b) * nblk * nq * nq, V.SB = bf.SB + size_t(b) * nblk * nq * nq;
  V.SC = bf.SC + size_t(b) * nblk * nq * nq;
  V.Jm = bf.Jm + size_t(b) * sc.T * sc.nu * nq, V.Jt = bf.Jt + size_t(b) * sc.T * sc.nu * nq;
  V.Jp = bf.Jp + size_t(b) * sc.T * sc.nu * nq;
  V.nq = nq, V.nu = sc.nu, V.T = sc.T, V.eq = sc.eq;
  return V;
}

}  // namespace idto
