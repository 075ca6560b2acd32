// Two-sided block elimination of the time-major KKT system, third generation: ONE warp factors the
// diagonal block, the right-hand sides are solved column-per-thread from registers.
//
// Same mathematics and the same results layout as k_kkt_tw2 (kernels_kkt2.cu; reference:
// PentaDiagonalFactorization, optimizer/penta_diagonal_solver.h:124-248, CalcLagrangeMultipliers cc:1371-1396,
// CalcDoglegPoint cc:2137-2140): CTA 0 of a 2-CTA cluster eliminates block rows 0..m top-down, CTA 1 rows
// N..m+1 bottom-up with
//     K_i = B_i - A_i Y_{i-2}        G_i = C_i - A_i Z_{i-2} - K_i Y_{i-1}
//     G_i [Y_i | Z_i | r_i] = [D_i - K_i Z_{i-1} | E_i | b_i - A_i r_{i-2} - K_i r_{i-1}]
// then an interface system for (x_m, x_{m+1}) and a concurrent back-substitution of both halves.
//
// What the second generation spent per block row (clock64, kb = 25, profiles/README.md): 11.3k cycles in the
// pivot loop + 5.2k in the back-substitution, because FOUR warps each eliminated the whole block redundantly
// (lane = row) with their share of the 2 kb + 1 right-hand-side columns riding along: every pivot step broadcast
// 4 x (trailing width) + 2 kb + 4 values through the SM's single shared-memory / shuffle crossbar, and the
// back-substitution was another 25 serial broadcast steps.  Here
//   * warp 0 alone runs the LU of G_i (lane = row, implicit partial pivoting, one REDUX per pivot search, the
//     reciprocal computed off the critical path — the proven inner loop of the second generation without any
//     right-hand-side cargo), recording the multipliers, the pivot order and U;
//   * meanwhile the other seven warps form everything that does not depend on that LU on the fp64 tensor
//     cores (DMMA m8n8k4): D_i - K_i Z_{i-1}, the r update, K_{i+1}, the Y_{i-1} / Z_{i-1} parts of G_{i+1}, and
//     prefetch the five blocks of row i+2;
//   * then 2 kb + 1 threads each take ONE right-hand-side column into registers in pivot order and run the
//     forward and backward substitutions as fully unrolled axpy sweeps (the L and U columns arrive as 16-byte
//     broadcast loads): no shuffles, no per-step synchronisation, 25-deep dependency chains instead of 25
//     broadcast round trips.
// Only the product K_i Y_{i-1} (all eight warps, one DMMA tile pair each) stays in front of the LU.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "kkt_view.cuh"
#include "reduce.cuh"
#include "solver.h"

#ifdef IDTO_KKT_TIMING
#define KT_DECL long long kt_t = clock64(), kt_acc[12] = {0}, kt_sub[8] = {0}, kt_s = 0;
#define KS0 kt_s = clock64();
#define KS(i) { const long long kt_n = clock64(); kt_sub[i] += kt_n - kt_s; kt_s = kt_n; }
#define KT(i) { const long long kt_n = clock64(); kt_acc[i] += kt_n - kt_t; kt_t = kt_n; }
#define KT_PRINT(b, dir)                                                                                    \
  if ((threadIdx.x == 0 || threadIdx.x == 32 || threadIdx.x == 128) && (b) == 0)                            \
    printf("kkt3 dir %d tid %d: prologue %lld s1 %lld bar1 %lld lu_or_help %lld bar2 %lld gather %lld rhs " \
           "%lld csync %lld iface_asm %lld iface_gj %lld csync2 %lld final %lld\n", dir, threadIdx.x,       \
           kt_acc[0], kt_acc[1], kt_acc[2], kt_acc[3], kt_acc[4], kt_acc[5], kt_acc[6], kt_acc[7],          \
           kt_acc[8], kt_acc[9], kt_acc[10], kt_acc[11]);                                                   \
  if ((threadIdx.x == 0 || threadIdx.x == 32 || threadIdx.x == 128) && (b) == 0)                            \
    printf("kkt3 sub dir %d tid %d: w0 load_g %lld lu %lld publish %lld | helpers issue+store %lld p2 %lld mz " \
           "%lld pre_row %lld hbar %lld commit %lld\n", dir, threadIdx.x, kt_sub[0], kt_sub[1], kt_sub[2],  \
           kt_sub[0], kt_sub[1], kt_sub[2], kt_sub[3], kt_sub[4], kt_sub[5]);
#else
#define KS0
#define KS(i)
#define KT_DECL
#define KT(i)
#define KT_PRINT(b, dir)
#endif

namespace idto {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kHelpers = kThreads - 32;  // warps 1..7

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void helper_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kHelpers) : "memory"); }
// (release: the stores of the calling lane before it are visible to whoever observes the completed phase)
__device__ __forceinline__ void lu_signal(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 1/x to ~1 ulp without the division subroutine: MUFU.RCP64H seed (>= 20 bits) + two Newton steps.
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// ---- warp-local LU of one KB x KB block --------------------------------------------------------------------
// g[column][lane = row].  Implicit partial pivoting: rows never move; at step c the pivot lane p_c is the
// unfinished row with the largest |g[.][c]| (judged on the exponent and 15 mantissa bits: partial pivoting up to a
// factor 1 + 2^-15).  The loop is rolled (straight-line code for all steps runs at instruction-fetch speed): after
// every step the live columns rotate one register to the left, so the pivot column is always g[0].
//   Lm[c][lane]  multiplier of row `lane` at step c (0 for finished rows and for the pivot row)
constexpr int kSeg = 4;

struct LuState {
  int p;         // pivot lane of the current step
  double m;      // multiplier of this lane's row for the current step
  bool done;     // this row has been a pivot row (or is a padding lane)
  bool fail;
  int ord;       // step at which this row was the pivot row (-1: not yet / padding lane)
  double myinv;  // reciprocal of this row's pivot
};

// Pivot search of step c on column values xc.  Besides the state it publishes, in PIVOT ORDER,
//   Uc[c][k] = U[k][c] / U[k][k] for the rows p_k chosen before step c (written by lane p_k itself),
//   invc[c] = 1 / pivot_c and piv[c] = p_c (written by the new pivot lane).
// Every lane executes every store (lanes with nothing to publish write to their own slot of `dump`): a conditional
// store becomes a divergent branch here, and a reconvergence point inside the pivot loop stops the compiler from
// interleaving the bulk shuffles of a step with the search of the next one (measured: 295 -> 435 cycles per step).
template <int LD>
__device__ __forceinline__ void lu_search(double xc, int c, int lane, double* Uc, double* invc, int* piv, double* dump,
                                          LuState& s) {
  const unsigned key =
      s.done ? 0u : (((unsigned(__double2hiint(xc)) & 0x7fffffe0u) + 32u) | unsigned(31 - lane));
  const unsigned mx = __reduce_max_sync(kFull, key);
  const double rc = fast_rcp(xc);
  const int p = 31 - int(mx & 31u);
  s.fail |= (mx < 64u) | (mx >= 0x7ff00020u);  // zero (or subnormal) pivot column, or inf / NaN in it
  const double inv = __shfl_sync(kFull, rc, p);
  const bool is_p = lane == p;
  double* udst = s.ord >= 0 ? Uc + c * LD + s.ord : dump + lane;
  *udst = xc * s.myinv;  // (off the critical path)
  s.m = (s.done || is_p) ? 0.0 : xc * inv;
  double* idst = is_p ? invc + c : dump + 32 + lane;
  *idst = inv;
  int* pdst = is_p ? piv + c : reinterpret_cast<int*>(dump + 64) + lane;
  *pdst = lane;
  s.done |= is_p;
  s.ord = is_p ? c : s.ord;
  s.myinv = is_p ? inv : s.myinv;
  s.p = p;
}

#ifndef IDTO_KKT3_SMEMLU
// The pivot row reaches the other rows through shuffles (2 SHFL per column).
template <int KB, int C0>
__device__ __forceinline__ void lu_segment(double (&g)[KB], int lane, double* Uc, double* invc, int* piv, double* Lm,
                                           double* prow, double* dump, uint64_t* bars, LuState& s) {
  constexpr int LD = (KB + 1) & ~1;
  constexpr int W = KB - 1 - C0;  // live columns right of the pivot column at the start of the segment
  constexpr int C1 = C0 + kSeg < KB ? C0 + kSeg : KB;
#pragma unroll 1
  for (int c = C0; c < C1; ++c) {
    const int p = s.p;
    const double m = s.m;
    Lm[c * 32 + lane] = m;
    double nxt = 0.0;
    if constexpr (W >= 1) nxt = fma(-m, __shfl_sync(kFull, g[1], p), g[1]);
    if (c + 1 < KB) lu_search<LD>(nxt, c + 1, lane, Uc, invc, piv, dump, s);  // overlaps the remaining updates of step c
#pragma unroll
    for (int j = 2; j <= W; ++j) g[j - 1] = fma(-m, __shfl_sync(kFull, g[j], p), g[j]);
    g[0] = nxt;
#ifndef IDTO_KKT3_NOPIPE
    // step c + 1 is published (pivot row, its U column and reciprocal; the multipliers of steps 0..c): the threads
    // that carry the right-hand-side columns may take it (every lane arrives: no branch in this loop)
    lu_signal(bars + (c + 1 < KB ? c + 1 : KB));
#endif
  }
  if constexpr (C1 < KB) lu_segment<KB, C1>(g, lane, Uc, invc, piv, Lm, prow, dump, bars, s);
}
#else
// The pivot row reaches the other rows through shared memory: the pivot lane stores its live columns (16 bytes per
// instruction) into prow[c & 1][.], everybody loads them back as broadcasts.
template <int KB, int C0>
__device__ __forceinline__ void lu_segment(double (&g)[KB], int lane, double* Uc, double* invc, int* piv, double* Lm,
                                           double* prow, double* dump, uint64_t* bars, LuState& s) {
  constexpr int LD = (KB + 1) & ~1;
  constexpr int W = KB - 1 - C0;
  constexpr int C1 = C0 + kSeg < KB ? C0 + kSeg : KB;
  constexpr int NP = (W + 1) / 2;  // 16-byte pairs covering g[1..W]
#pragma unroll 1
  for (int c = C0; c < C1; ++c) {
    const int p = s.p;
    const double m = s.m;
    Lm[c * 32 + lane] = m;
    double* buf = prow + (c & 1) * LD;
    if (lane == p) {
#pragma unroll
      for (int q = 0; q < NP; ++q)
        *reinterpret_cast<double2*>(buf + 2 * q) = make_double2(g[1 + 2 * q], 2 + 2 * q <= W ? g[2 + 2 * q] : 0.0);
    }
    __syncwarp();
    double nxt = 0.0;
    double2 v0 = make_double2(0.0, 0.0);
    if constexpr (W >= 1) {
      v0 = *reinterpret_cast<const double2*>(buf);
      nxt = fma(-m, v0.x, g[1]);
    }
    if (c + 1 < KB) lu_search<LD>(nxt, c + 1, lane, Uc, invc, piv, dump, s);
    if constexpr (W >= 2) g[1] = fma(-m, v0.y, g[2]);
#pragma unroll
    for (int q = 1; q < NP; ++q) {
      const double2 v = *reinterpret_cast<const double2*>(buf + 2 * q);
      g[2 * q] = fma(-m, v.x, g[1 + 2 * q]);
      if (2 + 2 * q <= W) g[1 + 2 * q] = fma(-m, v.y, g[2 + 2 * q]);
    }
    g[0] = nxt;
#ifndef IDTO_KKT3_NOPIPE
    lu_signal(bars + (c + 1 < KB ? c + 1 : KB));
#endif
  }
  if constexpr (C1 < KB) lu_segment<KB, C1>(g, lane, Uc, invc, piv, Lm, prow, dump, bars, s);
}
#endif

// ---- C = A * B with one output row per lane (FMA) -------------------------------------------------------------
// A: KB x KB, column-major with leading dimension LD (lane = row: conflict-free 8-byte loads); B: `ncols` columns
// of length KB at bcol(c), 16-byte aligned (the whole warp reads the same 16 bytes: broadcast).  The calling
// warp takes the columns cg, cg + ncg, ... (at most MAXC of them): MAXC independent accumulation chains.
template <int KB, int MAXC, class BCol, class Epi>
__device__ __forceinline__ void fma_prod(const double* A, int ncols, int cg, int ncg, int lane, BCol bcol, Epi epi) {
  constexpr int LD = (KB + 1) & ~1;
  const bool row = lane < KB;
  const int r = row ? lane : KB - 1;
  const double* bp[MAXC];
  double acc[MAXC];
#pragma unroll
  for (int j = 0; j < MAXC; ++j) {
    const int c = cg + j * ncg;
    bp[j] = bcol(c < ncols ? c : 0);
    acc[j] = 0.0;
  }
#pragma unroll
  for (int k = 0; k < KB; k += 2) {
    const double a0 = A[k * LD + r];
    const double a1 = k + 1 < KB ? A[(k + 1) * LD + r] : 0.0;
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      const double2 bv = *reinterpret_cast<const double2*>(bp[j] + k);
      acc[j] = fma(a0, bv.x, acc[j]);
      if (k + 1 < KB) acc[j] = fma(a1, bv.y, acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < MAXC; ++j) {
    const int c = cg + j * ncg;
    if (row && c < ncols) epi(r, c, acc[j]);
  }
}

// ---- C = A * B on the fp64 tensor cores (mma.sync m8n8k4, SASS DMMA) ----------------------------------------
// A: KB x KB, column-major with leading dimension LD, in shared memory.  B: `ncols` columns of length KB, column
// c at bcol(c).  Output tiles (8 columns x 16 rows: two DMMA row tiles share the B fragment) are dealt to the
// calling warps (w of nw); epi(row, col, value) receives every entry with row < KB, col < ncols.
template <int KB, class BCol, class Epi>
__device__ __forceinline__ void mma_prod(const double* A, int ncols, int w, int nw, int lane, BCol bcol, Epi epi) {
  constexpr int LD = (KB + 1) & ~1;
  constexpr int MT = (KB + 7) / 8, KT4 = (KB + 3) / 4, MT2 = (MT + 1) / 2;
  const int grp = lane >> 2, tig = lane & 3;
  const int ntn = (ncols + 7) / 8;
  for (int task = w; task < ntn * MT2; task += nw) {
    const int nt = task / MT2, mp = task - nt * MT2;
    const int r0 = (2 * mp) * 8 + grp, r1 = r0 + 8;
    const int nb = nt * 8 + grp;  // column this lane feeds
    const bool cok = nb < ncols;
    const double* bsrc = bcol(cok ? nb : 0);
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll
    for (int kt = 0; kt < KT4; ++kt) {
      const int kc = kt * 4 + tig;
      const bool kok = kc < KB;
      const double a0 = (kok && r0 < KB) ? A[kc * LD + r0] : 0.0;
      const double a1 = (kok && r1 < KB) ? A[kc * LD + r1] : 0.0;
      const double bfr = (kok && cok) ? bsrc[kc] : 0.0;
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c00), "+d"(c01)
                   : "d"(a0), "d"(bfr));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c10), "+d"(c11)
                   : "d"(a1), "d"(bfr));
    }
    const int t0 = nt * 8 + 2 * tig;
    if (r0 < KB) {
      if (t0 < ncols) epi(r0, t0, c00);
      if (t0 + 1 < ncols) epi(r0, t0 + 1, c01);
    }
    if (r1 < KB) {
      if (t0 < ncols) epi(r1, t0, c10);
      if (t0 + 1 < ncols) epi(r1, t0 + 1, c11);
    }
  }
}

}  // namespace

template <int KB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    k_kkt_v3(SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_fail;
  __shared__ int s_ord[64];
  __shared__ int s_piv[32];
  __shared__ __align__(8) uint64_t s_lubar[32];  // [c] "step c of the LU is published", one phase per block row
  const int b = blockIdx.x >> 1, dir = blockIdx.x & 1;
  if (!force && !bf.ctl[b].derivs_dirty) return;  // same decision in both CTAs of the cluster
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  constexpr int kb = KB, kk = KB * KB;
  constexpr int LD = (KB + 1) & ~1;  // even leading dimension: 16-byte aligned columns
  constexpr int kl = KB * LD;        // one padded column-major block
  constexpr int NRHS = 2 * KB + 1;   // [Y | Z | r]
  const int nblk = sc.T + 1, N = sc.T, nq = sc.nq;
  const int mid = N / 2;  // forward: rows 0..mid, backward: rows N..mid+1
  const int nsteps = dir == 0 ? mid + 1 : N - mid;
  const int sgn = dir == 0 ? 1 : -1, first = dir == 0 ? 0 : N;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool row = lane < kb;
  const int r = row ? lane : kb - 1;
  const KktView V = make_kkt_view(sc, bf, b);
  const double* gs = bf.gs + size_t(b) * sc.n;
  const double* h = bf.st.h + size_t(b) * sc.nh;
  // ---- shared memory -------------------------------------------------------------------------------
  double* Yb = sm;                  // [2][kl]  Y of the two previous rows (parity of the step index)
  double* Zb = Yb + 2 * kl;         // [2][kl]
  double* raw = Zb + 2 * kl;        // [2][5][kl] blocks of rows n, n+1: back-2, back-1, diag, front-1, front-2
  double* Kb = raw + 10 * kl;       // [2][kl]  K_n (parity of n)
  double* Mg = Kb + 2 * kl;         // [2][kl]  G_n (parity of n; the K Y_{n-1} part is subtracted in step n)
  double* My = Mg + 2 * kl;         // [kl]     D_n - K_n Z_{n-1}
  double* Mz = My + kl;             // [kl]     E_n
  double* Lc = Mz + kl;             // [kl]     L in pivot order: Lc[c][k] = multiplier of row p_k at step c (k > c)
  double* Uc = Lc + kl;             // [kl]     unit-diagonal U in pivot order: Uc[c][k] = U[k][c] / U[k][k]  (k < c)
  double* rb = Uc + kl;             // [2][LD]  r of the two previous rows
  double* rawb = rb + 2 * LD;       // [2][LD]  b_n
  double* rv = rawb + 2 * LD;       // [2][LD]  right-hand side of row n
  double* invc = rv + 2 * LD;       // [LD]     1 / pivot of step c
  double* prow = invc + LD;         // [2][LD]  pivot-row broadcast scratch of the LU warp
  double* dump = prow + 2 * LD;     // [96]     where lanes with nothing to publish store (lu_search)
  double* Lm = dump + 96;           // [KB][32] Lm[c][lane]: multiplier of row `lane` at step c
  constexpr int kSweepDoubles = 22 * kl + 9 * LD + 96 + KB * 32;
  double* FY = bf.FY + size_t(b) * nblk * kk;
  double* FZ = bf.FZ + size_t(b) * nblk * kk;
  double* Fr = bf.X + size_t(b) * nblk * kb;
  double* xq = bf.pH + size_t(b) * sc.n;
  double* lam = bf.lambda + size_t(b) * sc.nh;
  double* xint = bf.tmp2 + size_t(b) * sc.n;  // interface solution (x_mid, x_mid+1): 2*kb doubles (n >= 2*kb)
  for (int e = tid; e < kSweepDoubles; e += kThreads) sm[e] = 0.0;
  if (tid == 0) s_fail = 0;
  if (tid < 32) mbar_init(s_lubar + tid, 32);  // every lane of the LU warp arrives
  __syncthreads();
  KT_DECL

  // Blocks of the row visited at step n.  `issue` reads them into registers (independent loads, all in flight at
  // once), `commit` stores them to raw[n & 1] later: the loads of row n+2 fly while the helpers multiply.
  // The loads walk the SOURCE arrays (five nq x nq blocks of the scaled Hessian bands, five nu x nq blocks of the
  // scaled Jacobian bands: contiguous, coalesced, one add per address); commit scatters them into the five kb x kb
  // blocks of the KKT row (the blocks in front of the diagonal are transposes).  Entries no source covers are the
  // structural zeros of the time-major ordering: they stay zero from the initial fill.
  const int nqq = nq * nq, nuk = kb - nq, njq = nuk * nq;
  // slots of raw[.]: back-2, back-1, diag, front-1, front-2 along the direction of the sweep
  const int sl_m1 = dir == 0 ? 1 : 3, sl_p1 = dir == 0 ? 3 : 1, sl_m2 = dir == 0 ? 0 : 4, sl_p2 = dir == 0 ? 4 : 0;
  auto issue_row = [&](int n, int t0, int nt, auto& vh, auto& vj, double& bv) {
    constexpr int NH = std::extent<std::remove_reference_t<decltype(vh)>>::value;
    constexpr int NJ = std::extent<std::remove_reference_t<decltype(vj)>>::value;
    bv = 0.0;
    if (n >= nsteps) return;
    const int i = first + sgn * n;
    const bool vm1 = i >= 1, vm2 = i >= 2, vp1 = i + 1 <= N, vp2 = i + 2 <= N;
    const double* pC = V.SC + size_t(i) * nqq;
    const double* pB = V.SB + size_t(i) * nqq;          // block (i, i-1)
    const double* pB1 = V.SB + size_t(i + 1) * nqq;     // block (i+1, i): transposed into (i, i+1)
    const double* pA = V.SA + size_t(i) * nqq;          // block (i, i-2)
    const double* pA2 = V.SA + size_t(i + 2) * nqq;     // block (i+2, i)
#pragma unroll
    for (int k = 0; k < NH; ++k) {
      const int e = t0 + k * nt;
      const bool ok = e < nqq;
      vh[k][0] = ok ? pC[e] : 0.0;
      vh[k][1] = ok && vm1 ? pB[e] : 0.0;
      vh[k][2] = ok && vp1 ? pB1[e] : 0.0;
      vh[k][3] = ok && vm2 ? pA[e] : 0.0;
      vh[k][4] = ok && vp2 ? pA2[e] : 0.0;
    }
    const double* jP = V.Jp + (ptrdiff_t(i) - 1) * njq;   // rows (nq+u) / columns (nq+u) of the diagonal block
    const double* jT0 = V.Jt + (ptrdiff_t(i) - 1) * njq;  // rows nq+u of block (i, i-1)
    const double* jT1 = V.Jt + ptrdiff_t(i) * njq;        // columns nq+u of block (i, i+1)
    const double* jM0 = V.Jm + (ptrdiff_t(i) - 1) * njq;  // rows nq+u of block (i, i-2)
    const double* jM1 = V.Jm + (ptrdiff_t(i) + 1) * njq;  // columns nq+u of block (i, i+2)
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      const int e = t0 + k * nt;
      const bool ok = e < njq;
      vj[k][0] = ok && vm1 ? jP[e] : 0.0;
      vj[k][1] = ok && vm1 ? jT0[e] : 0.0;
      vj[k][2] = ok && vp1 ? jT1[e] : 0.0;
      vj[k][3] = ok && vm2 ? jM0[e] : 0.0;
      vj[k][4] = ok && vp2 ? jM1[e] : 0.0;
    }
    if (t0 < kb) bv = t0 < nq ? -gs[i * nq + t0] : (i >= 1 ? -h[(i - 1) * sc.nu + (t0 - nq)] : 0.0);
  };
  // (dn, dt: where element k of the calling thread goes in a block and in its transpose; fixed for the whole sweep)
  auto commit_row = [&](int n, int t0, int nt, const auto& vh, const auto& vj, double bv, const auto& hdn,
                        const auto& hdt, const auto& jdn, const auto& jdt) {
    constexpr int NH = std::extent<std::remove_reference_t<decltype(vh)>>::value;
    constexpr int NJ = std::extent<std::remove_reference_t<decltype(vj)>>::value;
    if (n >= nsteps) return;
    const int i = first + sgn * n;
    double* dst = raw + (n & 1) * 5 * kl;
#pragma unroll
    for (int k = 0; k < NH; ++k) {
      if (t0 + k * nt < nqq) {
        dst[2 * kl + hdn[k]] = vh[k][0];
        dst[sl_m1 * kl + hdn[k]] = vh[k][1];
        dst[sl_p1 * kl + hdt[k]] = vh[k][2];
        dst[sl_m2 * kl + hdn[k]] = vh[k][3];
        dst[sl_p2 * kl + hdt[k]] = vh[k][4];
      }
    }
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      if (t0 + k * nt < njq) {
        dst[2 * kl + jdn[k]] = vj[k][0];
        dst[2 * kl + jdt[k]] = vj[k][0];
        dst[sl_m1 * kl + jdn[k]] = vj[k][1];
        dst[sl_p1 * kl + jdt[k]] = vj[k][2];
        dst[sl_m2 * kl + jdn[k]] = vj[k][3];
        dst[sl_p2 * kl + jdt[k]] = vj[k][4];
      }
    }
    if (t0 < nuk) dst[2 * kl + (nq + t0) * LD + nq + t0] = i == 0 ? 1.0 : 0.0;  // dummy lambda_{-1}
    if (t0 < kb) rawb[(n & 1) * LD + t0] = bv;
  };
  auto make_maps = [&](int t0, int nt, auto& hdn, auto& hdt, auto& jdn, auto& jdt) {
    constexpr int NH = std::extent<std::remove_reference_t<decltype(hdn)>>::value;
    constexpr int NJ = std::extent<std::remove_reference_t<decltype(jdn)>>::value;
#pragma unroll
    for (int k = 0; k < NH; ++k) {
      const int e = t0 + k * nt, c = e / nq, rr = e - c * nq;
      hdn[k] = c * LD + rr, hdt[k] = rr * LD + c;
    }
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      const int e = t0 + k * nt, u = e / nq, cc = e - u * nq;  // Jacobian bands are row-major (u, cc)
      jdn[k] = cc * LD + nq + u, jdt[k] = (nq + u) * LD + cc;
    }
  };

  // Products with the results of row n-2 (they do not depend on row n-1): K_n, the first part of G_n and of
  // the right-hand side r_n, by the calling warps (w of nw).  Y_{n-2} lives in Yb[n & 1].
  auto pre_row = [&](int n, int w, int nw) {
    if (n >= nsteps) return;
    const int par = n & 1;
    const double* rw = raw + par * 5 * kl;
    const double* Yp = Yb + par * kl;
    const double* Zp = Zb + par * kl;
    const double* rp = rb + par * LD;
    auto bcol = [&](int c) { return c < kb ? Yp + c * LD : (c < 2 * kb ? Zp + (c - kb) * LD : rp); };
    auto epi = [&](int rr, int c, double acc) {
      if (c < kb)
        Kb[par * kl + c * LD + rr] = rw[kl + c * LD + rr] - acc;
      else if (c < 2 * kb)
        Mg[par * kl + (c - kb) * LD + rr] = rw[2 * kl + (c - kb) * LD + rr] - acc;
      else
        rv[par * LD + rr] = rawb[par * LD + rr] - acc;
    };
#ifndef IDTO_KKT3_FMA
    mma_prod<KB>(rw, NRHS, w, nw, lane, bcol, epi);
#else
    if (nw == kWarps)
      fma_prod<KB, (NRHS + kWarps - 1) / kWarps>(rw, NRHS, w, nw, lane, bcol, epi);
    else
      fma_prod<KB, (NRHS + kWarps - 2) / (kWarps - 1)>(rw, NRHS, w, nw, lane, bcol, epi);
#endif
  };
  // Y_n, Z_n, r_n of step n (in shared memory since the end of that step) -> HBM for the back-substitution
  auto store_row = [&](int n, int t0, int nt) {
    const int i = first + sgn * n, par = n & 1;
    for (int e = t0; e < kk; e += nt) {
      const int c = e / kb, rr = e - c * kb;
      FY[size_t(i) * kk + e] = Yb[par * kl + c * LD + rr];
      FZ[size_t(i) * kk + e] = Zb[par * kl + c * LD + rr];
    }
    for (int e = t0; e < kb; e += nt) Fr[size_t(i) * kb + e] = rb[par * LD + e];
  };

  {
    constexpr int NHA = (kk + kThreads - 1) / kThreads, NJA = (kk / 4 + kThreads - 1) / kThreads;  // nu nq <= kb^2 / 4
    int hdn[NHA], hdt[NHA], jdn[NJA], jdt[NJA];
    make_maps(tid, kThreads, hdn, hdt, jdn, jdt);
    double h0[NHA][5], j0[NJA][5], h1[NHA][5], j1[NJA][5], b0, b1;
    issue_row(0, tid, kThreads, h0, j0, b0);
    issue_row(1, tid, kThreads, h1, j1, b1);
    commit_row(0, tid, kThreads, h0, j0, b0, hdn, hdt, jdn, jdt);
    commit_row(1, tid, kThreads, h1, j1, b1, hdn, hdt, jdn, jdt);
  }
  __syncthreads();
  pre_row(0, wid, kWarps);
  __syncthreads();
  KT(0)

  // the threads that prefetch the rows during the sweep, and where their elements go
#ifndef IDTO_KKT3_NOPIPE
  constexpr int kLoadT = kThreads - 96;  // warps 3..7
#else
  constexpr int kLoadT = kHelpers;       // warps 1..7
#endif
  constexpr int NHL = (kk + kLoadT - 1) / kLoadT, NJL = (kk / 4 + kLoadT - 1) / kLoadT;
  const int lt = tid - (kThreads - kLoadT);
  int hdn[NHL], hdt[NHL], jdn[NJL], jdt[NJL];
  make_maps(lt < 0 ? 0 : lt, kLoadT, hdn, hdt, jdn, jdt);

  for (int n = 0; n < nsteps; ++n) {
    const int cur = n & 1, prv = cur ^ 1;
    // ---- S1 (all warps): G_n -= K_n Y_{n-1} ------------------------------------------------------------
    {
      const double* Yp = Yb + prv * kl;
      double* G = Mg + cur * kl;
#ifndef IDTO_KKT3_FMA
      mma_prod<KB>(
          Kb + cur * kl, kb, wid, kWarps, lane, [&](int c) { return Yp + c * LD; },
          [&](int rr, int c, double acc) { G[c * LD + rr] -= acc; });
#else
      // (one output row per lane, 4 independent FMA chains per thread)
      fma_prod<KB, (KB + kWarps - 1) / kWarps>(
          Kb + cur * kl, kb, wid, kWarps, lane, [&](int c) { return Yp + c * LD; },
          [&](int rr, int c, double acc) { G[c * LD + rr] -= acc; });
#endif
    }
    KT(1)
    __syncthreads();
    KT(2)
#ifndef IDTO_KKT3_NOPIPE
    if (wid == 0) {
      // ---- S2, warp 0: LU of G_n; every step is signalled to the right-hand-side threads as it is published ----
      double g[KB];
      KS0
#pragma unroll
      for (int c = 0; c < KB; ++c) g[c] = row ? Mg[cur * kl + c * LD + r] : 0.0;
      KS(0)
      LuState s;
      s.done = lane >= KB, s.fail = false, s.ord = -1, s.myinv = 0.0;
      lu_search<LD>(g[0], 0, lane, Uc, invc, s_piv, dump, s);
      lu_signal(s_lubar);
      lu_segment<KB, 0>(g, lane, Uc, invc, s_piv, Lm, prow, dump, s_lubar, s);
      KS(1)
      if (s.fail && lane == 0) s_fail = 1;
      KS(2)
    } else {
      // ---- S2, warps 1..7 ---------------------------------------------------------------------------------
      // all of them: the right-hand sides D_n - K_n Z_{n-1}, E_n, r_n;  then warps 1, 2: one right-hand-side column
      // per thread, forward substitution in step with the LU (step c needs the pivot row p_c and the multipliers
      // of row p_c at the steps before it: y_c = b[p_c] - sum_{k<c} Lm[k][p_c] y_k) and the backward substitution
      // once U is complete;  warps 3..7: prefetch of row n+2, Y/Z/r of row n-1 to HBM, the products of row n+1.
      constexpr int kPre = kThreads - 96;                    // threads of warps 3..7
      const int hw = wid - 1;
      const int pt = tid - 96;
      double vh[NHL][5], vj[NJL][5], bvn = 0.0;
      KS0
      const double* rw = raw + cur * 5 * kl;
      const double* Zp = Zb + prv * kl;
      const double* rp = rb + prv * LD;
      {
        auto bcol = [&](int c) { return c < kb ? Zp + c * LD : rp; };
        auto epi = [&](int rr, int c, double acc) {
          if (c < kb)
            My[c * LD + rr] = rw[3 * kl + c * LD + rr] - acc;
          else
            rv[cur * LD + rr] -= acc;
        };
#ifndef IDTO_KKT3_FMA
        mma_prod<KB>(Kb + cur * kl, kb + 1, hw, kWarps - 1, lane, bcol, epi);
#else
        fma_prod<KB, (KB + 1 + kWarps - 2) / (kWarps - 1)>(Kb + cur * kl, kb + 1, hw, kWarps - 1, lane, bcol, epi);
#endif
      }
      KS(1)
      for (int e = tid - 32; e < kl; e += kHelpers) Mz[e] = rw[4 * kl + e];  // E_n (raw[cur] is overwritten below)
      KS(2)
      helper_barrier();  // the right-hand sides are complete; nobody reads raw[cur] any more
      KS(3)
      if (wid >= 3) {
        issue_row(n + 2, pt, kPre, vh, vj, bvn);  // loads fly during the products below
        if (n >= 1) store_row(n - 1, pt, kPre);
        KS(0)
        pre_row(n + 1, wid - 3, kWarps - 3);
        KS(4)
        commit_row(n + 2, pt, kPre, vh, vj, bvn, hdn, hdt, jdn, jdt);
      } else {
        const int t = tid - 32;  // flat column: Y (t < kb), Z (t < 2 kb), r (t == 2 kb)
        if (t < NRHS) {
          const double* src = t < kb ? My + t * LD : (t < 2 * kb ? Mz + (t - kb) * LD : rv + cur * LD);
          const unsigned phase = unsigned(n & 1);
          double y[KB];
#pragma unroll
          for (int c = 0; c < KB; ++c) {
            mbar_wait(s_lubar + c, phase);
            const int p = s_piv[c];
            const double* lp = Lm + p;
            double a0 = src[p], a1 = 0.0;
#pragma unroll
            for (int k = 0; k < c; ++k) {
              if (k & 1)
                a1 = fma(-lp[k * 32], y[k], a1);
              else
                a0 = fma(-lp[k * 32], y[k], a0);
            }
            y[c] = a0 + a1;
          }
          // backward: unit-diagonal U~ = diag(U)^-1 U, x = U~^-1 diag(U)^-1 y
#pragma unroll
          for (int k = 0; k < KB; ++k) y[k] *= invc[k];
#pragma unroll
          for (int c = KB - 1; c >= 1; --c) {
            const double xc = y[c];
#pragma unroll
            for (int k = 0; k < c; k += 2) {
              const double2 u = *reinterpret_cast<const double2*>(Uc + c * LD + k);
              y[k] = fma(-u.x, xc, y[k]);
              if (k + 1 < c) y[k + 1] = fma(-u.y, xc, y[k + 1]);
            }
          }
          double* dst = t < kb ? Yb + cur * kl + t * LD : (t < 2 * kb ? Zb + cur * kl + (t - kb) * LD : rb + cur * LD);
#pragma unroll
          for (int k = 0; k < KB; ++k) dst[k] = y[k];
        }
        KS(5)
      }
    }
    KT(3)
#else
    if (wid == 0) {
      // ---- S2, warp 0: LU of G_n ------------------------------------------------------------------------
      double g[KB];
      KS0
#pragma unroll
      for (int c = 0; c < KB; ++c) g[c] = row ? Mg[cur * kl + c * LD + r] : 0.0;
      KS(0)
      LuState s;
      s.done = lane >= KB, s.fail = false, s.ord = -1, s.myinv = 0.0;
      lu_search<LD>(g[0], 0, lane, Uc, invc, s_piv, dump, s);
      lu_segment<KB, 0>(g, lane, Uc, invc, s_piv, Lm, prow, dump, s_lubar, s);
#ifdef IDTO_KKT3_SERIAL
      asm volatile("bar.sync 2, 256;" ::: "memory");  // experiment: the helpers start after the LU
#endif
      KS(1)
      if (s.fail && lane == 0) s_fail = 1;
      KS(2)
    } else {
      // ---- S2, warps 1..7: everything that does not need the LU -----------------------------------------
      const int ht = tid - 32, hw = wid - 1;
      double vh[NHL][5], vj[NJL][5], bvn;
#ifdef IDTO_KKT3_SERIAL
      asm volatile("bar.sync 2, 256;" ::: "memory");
#endif
      KS0
      issue_row(n + 2, ht, kHelpers, vh, vj, bvn);  // loads fly during the products below
      if (n >= 1) store_row(n - 1, ht, kHelpers);
      KS(0)
      const double* rw = raw + cur * 5 * kl;
      const double* Zp = Zb + prv * kl;
      const double* rp = rb + prv * LD;
      // D_n - K_n Z_{n-1} and the r update: K_n [Z_{n-1} | r_{n-1}]
      {
        auto bcol = [&](int c) { return c < kb ? Zp + c * LD : rp; };
        auto epi = [&](int rr, int c, double acc) {
          if (c < kb)
            My[c * LD + rr] = rw[3 * kl + c * LD + rr] - acc;
          else
            rv[cur * LD + rr] -= acc;
        };
#ifndef IDTO_KKT3_FMA
        mma_prod<KB>(Kb + cur * kl, kb + 1, hw, kWarps - 1, lane, bcol, epi);
#else
        fma_prod<KB, (KB + 1 + kWarps - 2) / (kWarps - 1)>(Kb + cur * kl, kb + 1, hw, kWarps - 1, lane, bcol, epi);
#endif
      }
      KS(1)
      for (int e = ht; e < kl; e += kHelpers) Mz[e] = rw[4 * kl + e];  // E_n (raw[cur] is overwritten below)
      KS(2)
      pre_row(n + 1, hw, kWarps - 1);
      KS(3)
      helper_barrier();  // every helper is done reading raw[cur]
      KS(4)
      commit_row(n + 2, ht, kHelpers, vh, vj, bvn, hdn, hdt, jdn, jdt);
      KS(5)
    }
    KT(3)
    __syncthreads();
    KT(4)
    // ---- L in pivot order (all threads): the right-hand-side threads then read it with 16-byte broadcast loads ----
    for (int e = tid; e < kk; e += kThreads) {
      const int c = e / kb, k = e - c * kb;
      Lc[c * LD + k] = k > c ? Lm[c * 32 + s_piv[k]] : 0.0;
    }
    KT(5)
    __syncthreads();
    // ---- S3: one right-hand-side column per thread (warps 1, 2, ...) -------------------------------------------
    {
      const int t = tid - 32;  // flat column: Y (t < kb), Z (t < 2 kb), r (t == 2 kb)
      if (t >= 0 && t < NRHS) {
        const double* src = t < kb ? My + t * LD : (t < 2 * kb ? Mz + (t - kb) * LD : rv + cur * LD);
        double x[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) x[k] = src[s_piv[k]];
        // forward: L y = P b (axpy form: column c of L updates the rows chosen later)
#pragma unroll
        for (int c = 0; c < KB - 1; ++c) {
          const double xc = x[c];
#pragma unroll
          for (int k = (c + 1) & ~1; k < KB; k += 2) {  // (an even c revisits k = c with a zero multiplier)
            const double2 l = *reinterpret_cast<const double2*>(Lc + c * LD + k);
            x[k] = fma(-l.x, xc, x[k]);
            if (k + 1 < KB) x[k + 1] = fma(-l.y, xc, x[k + 1]);
          }
        }
        // backward: unit-diagonal U~ = diag(U)^-1 U, x = U~^-1 diag(U)^-1 y
#pragma unroll
        for (int k = 0; k < KB; ++k) x[k] *= invc[k];
#pragma unroll
        for (int c = KB - 1; c >= 1; --c) {
          const double xc = x[c];
#pragma unroll
          for (int k = 0; k < c; k += 2) {
            const double2 u = *reinterpret_cast<const double2*>(Uc + c * LD + k);
            x[k] = fma(-u.x, xc, x[k]);
            if (k + 1 < c) x[k + 1] = fma(-u.y, xc, x[k + 1]);
          }
        }
        double* dst = t < kb ? Yb + cur * kl + t * LD : (t < 2 * kb ? Zb + cur * kl + (t - kb) * LD : rb + cur * LD);
#pragma unroll
        for (int k = 0; k < KB; ++k) dst[k] = x[k];
      }
    }
    KT(6)
#endif
    __syncthreads();
  }
  store_row(nsteps - 1, tid, kThreads);
  if (tid == 0 && s_fail) atomicExch(bf.status, IDTO_ERR_FACTORIZATION);
  __threadfence();
  cluster.sync();  // both chains (and their Y, Z, r in HBM) are complete
  KT(7)

  // ---- interface system for u = (x_mid, x_mid+1), solved by CTA 0 ----------------------------------------
  //   (I - Z_m Z'_{m+2}) x_m + (Y_m - Z_m Y'_{m+2}) x_{m+1} = r_m - Z_m r'_{m+2}
  //   (Y'_{m+1} - Z'_{m+1} Y_{m-1}) x_m + (I - Z'_{m+1} Z_{m-1}) x_{m+1} = r'_{m+1} - Z'_{m+1} r_{m-1}
  // (primes: the bottom-up chain; a missing neighbour contributes zero blocks)
  constexpr int n2 = 2 * KB, W2 = 2 * KB + 1;
  if (dir == 0) {
    double* Q = sm;  // n2 x W2 column-major; the sweep buffers are dead
    __syncthreads();
    const int m0 = mid;
    const bool has_m2 = m0 + 2 <= N, has_mm1 = m0 - 1 >= 0;
    double* S = sm + ((n2 * W2 + 1) & ~1);  // [8][kk]: Y, Z of rows m-1, m, m+1, m+2
    double* Sr = S + 8 * kk;                // [4][kb]
    {
      constexpr int NE = (8 * kk + kThreads - 1) / kThreads;  // static trip count: all loads issue before the stores
      double tmp[NE];
#pragma unroll
      for (int k = 0; k < NE; ++k) {
        const int e = tid + k * kThreads, blk = e / kk, o = e - blk * kk, rowi = m0 - 1 + (blk >> 1);
        const bool ok = e < 8 * kk && (rowi != m0 - 1 || has_mm1) && (rowi != m0 + 2 || has_m2);
        tmp[k] = ok ? __ldcg(((blk & 1) ? FZ : FY) + size_t(rowi) * kk + o) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < NE; ++k)
        if (tid + k * kThreads < 8 * kk) S[tid + k * kThreads] = tmp[k];
    }
    for (int e = tid; e < 4 * kb; e += kThreads) {
      const int rowi = m0 - 1 + e / kb;
      const bool ok = (rowi != m0 - 1 || has_mm1) && (rowi != m0 + 2 || has_m2);
      Sr[e] = ok ? __ldcg(Fr + size_t(rowi) * kb + e % kb) : 0.0;
    }
    __syncthreads();
    const double *Ym1 = S, *Zm1 = S + kk, *Ym = S + 2 * kk, *Zm = S + 3 * kk, *Yp1 = S + 4 * kk, *Zp1 = S + 5 * kk,
                 *Yp2 = S + 6 * kk, *Zp2 = S + 7 * kk;
    for (int e = tid; e < n2 * W2; e += kThreads) {
      const int c = e / n2, rr = e % n2;
      const int br = rr / kb, r1 = rr % kb;
      const double* Zl = br == 0 ? Zm : Zp1;
      double val;
      if (c < n2) {
        const int bc = c / kb, cc = c % kb;
        const double* Rt;
        if (br == 0 && bc == 0) val = (r1 == cc) ? 1.0 : 0.0, Rt = Zp2;          // I - Z_m Z'_{m+2}
        else if (br == 0) val = Ym[cc * kb + r1], Rt = Yp2;                      // Y_m - Z_m Y'_{m+2}
        else if (bc == 0) val = Yp1[cc * kb + r1], Rt = Ym1;                     // Y'_{m+1} - Z'_{m+1} Y_{m-1}
        else val = (r1 == cc) ? 1.0 : 0.0, Rt = Zm1;                             // I - Z'_{m+1} Z_{m-1}
        double v2 = 0.0;
#pragma unroll
        for (int j = 0; j < KB; j += 2) {
          val -= Zl[j * kb + r1] * Rt[cc * kb + j];
          if (j + 1 < KB) v2 -= Zl[(j + 1) * kb + r1] * Rt[cc * kb + j + 1];
        }
        val += v2;
      } else {
        const double* rt = br == 0 ? Sr + 3 * kb : Sr;  // r'_{m+2} or r_{m-1}
        val = br == 0 ? Sr[kb + r1] : Sr[2 * kb + r1];   // r_m or r'_{m+1}
        double v2 = 0.0;
#pragma unroll
        for (int j = 0; j < KB; j += 2) {
          val -= Zl[j * kb + r1] * rt[j];
          if (j + 1 < KB) v2 -= Zl[(j + 1) * kb + r1] * rt[j + 1];
        }
        val += v2;
      }
      Q[e] = val;
    }
    __syncthreads();
    KT(8)
    // Gauss-Jordan with implicit partial pivoting on the n2 x (n2+1) system, ONE barrier per step
    {
      constexpr int NCG = kThreads / 64;
      const int rr = tid & 63, cg = tid >> 6;
      const bool rowok = rr < n2;
      bool d0 = lane >= n2, d1 = lane + 32 >= n2;
      bool bad = false;
      for (int c = 0; c < n2; ++c) {
        const double a0 = d0 ? 0.0 : Q[c * n2 + lane], a1 = d1 ? 0.0 : Q[c * n2 + lane + 32];
        const unsigned k0 = d0 ? 0u : (((unsigned(__double2hiint(a0)) & 0x7fffffc0u) + 64u) | unsigned(63 - lane));
        const unsigned k1 = d1 ? 0u : (((unsigned(__double2hiint(a1)) & 0x7fffffc0u) + 64u) | unsigned(31 - lane));
        const unsigned mx = __reduce_max_sync(kFull, k0 > k1 ? k0 : k1);
        const int p = 63 - int(mx & 63u);
        bad |= (mx < 128u) | (mx >= 0x7ff00040u);
        d0 |= p == lane, d1 |= p == lane + 32;
        const double inv = fast_rcp(Q[c * n2 + p]);
        if (tid == 0) s_ord[p] = c;
        if (rowok && rr != p) {
          const double m = Q[c * n2 + rr] * inv;
#pragma unroll 4
          for (int j = c + 1 + cg; j < W2; j += NCG) Q[j * n2 + rr] = fma(-m, Q[j * n2 + p], Q[j * n2 + rr]);
        }
        __syncthreads();
      }
      if (bad && tid == 0) atomicExch(bf.status, IDTO_ERR_FACTORIZATION);
      if (tid < n2) xint[s_ord[tid]] = Q[n2 * n2 + tid] / Q[s_ord[tid] * n2 + tid];
    }
    __threadfence();
    KT(9)
  }
  cluster.sync();  // the interface solution is visible to both CTAs
  KT(10)

  // ---- back-substitution of each half ------------------------------------------------------------------
  //   top half:    x_i = r_i - Y_i x_{i+1} - Z_i x_{i+2},  i = mid-1 .. 0
  //   bottom half: x_i = r_i - Y_i x_{i-1} - Z_i x_{i-2},  i = mid+2 .. N
  __syncthreads();
  // Warp 0 owns the recurrence (lane = row); all threads stream Y_i, Z_i, r_i from L2 into a ring of kRing
  // shared-memory slots, kRing - 1 rows ahead of it (an L2 round trip is ~800 cycles, a row of the recurrence ~300:
  // with the 3-slot ring of the second generation every row waited for its data, 1.25k cycles per row).
  constexpr int kSlot = 2 * kk + LD;
  constexpr int kRing = 6;
  double* ring = sm;                  // [kRing][Y | Z | r]
  double* xs = sm + kRing * kSlot;    // [2][LD] the two most recent solution blocks
  const int istart = dir == 0 ? mid - 1 : mid + 2;
  const int nrows = dir == 0 ? mid : N - mid - 1;
  auto fetch = [&](int it, int t0, int nt) {
    if (it < nrows) {
      const int i = istart - sgn * it;
      double* dst = ring + (it % kRing) * kSlot;
      for (int e = t0; e < kk; e += nt) {
        cp_async8(dst + e, FY + size_t(i) * kk + e);
        cp_async8(dst + kk + e, FZ + size_t(i) * kk + e);
      }
      for (int e = t0; e < kb; e += nt) cp_async8(dst + 2 * kk + e, Fr + size_t(i) * kb + e);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per call, also when empty
  };
  auto emit = [&](int i, double out) {
    if (row) {
      if (lane < nq)
        xq[i * nq + lane] = out;
      else if (i >= 1)
        lam[(i - 1) * sc.nu + (lane - nq)] = out;
    }
  };
#pragma unroll
  for (int k = 0; k < kRing - 1; ++k) fetch(k, tid, kThreads);
  if (wid == 0) {
    const double u0 = row ? __ldcg(xint + lane) : 0.0, u1 = row ? __ldcg(xint + kb + lane) : 0.0;
    if (lane < LD) xs[lane] = dir == 0 ? u0 : u1, xs[LD + lane] = dir == 0 ? u1 : u0;
    emit(dir == 0 ? mid : mid + 1, dir == 0 ? u0 : u1);
  }
  asm volatile("cp.async.wait_group %0;" ::"n"(kRing - 2) : "memory");  // row 0 has landed (the others may be in flight)
  __syncthreads();
  for (int it = 0; it < nrows; ++it) {
    if (wid == 0) {
      const double* Ys = ring + (it % kRing) * kSlot;
      const double* Zs = Ys + kk;
      const double* x1 = xs + (it & 1) * LD;
      const double* x2 = xs + ((it & 1) ^ 1) * LD;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int j = 0; j < KB; j += 2) {
        const double2 v1 = *reinterpret_cast<const double2*>(x1 + j);
        const double2 v2 = *reinterpret_cast<const double2*>(x2 + j);
        a0 = fma(Ys[j * kb + r], v1.x, a0), a2 = fma(Zs[j * kb + r], v2.x, a2);
        if (j + 1 < KB) a1 = fma(Ys[(j + 1) * kb + r], v1.y, a1), a3 = fma(Zs[(j + 1) * kb + r], v2.y, a3);
      }
      const double out = row ? (Ys[2 * kk + r] - (a0 + a1)) - (a2 + a3) : 0.0;
      emit(istart - sgn * it, out);
      __syncwarp();
      if (lane < LD) xs[((it & 1) ^ 1) * LD + lane] = out;  // becomes x1 of the next row; the old x1 becomes x2
    }
    fetch(it + kRing - 1, tid, kThreads);  // its slot, (it - 1) % kRing, was consumed in iteration it - 1
    asm volatile("cp.async.wait_group %0;" ::"n"(kRing - 2) : "memory");  // row it + 1 has landed
    __syncthreads();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  KT(11)
  KT_PRINT(b, dir)
}

template <int KB>
static void launch_v3_kb(const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  constexpr int LD = (KB + 1) & ~1, kl = KB * LD;
  const int sweep = 22 * kl + 9 * LD + 96 + KB * 32;
  const int tail = std::max(((2 * KB * (2 * KB + 1) + 1) & ~1) + 8 * KB * KB + 4 * KB, 6 * (2 * KB * KB + LD) + 2 * LD);
  const int smem = std::max(sweep, tail) * 8;
  static bool attr_set[kMaxDevices] = {};
  if (first_use_on_device(attr_set))
    cudaFuncSetAttribute(k_kkt_v3<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);  // (+ static: the sum is capped at 227 KB)
  k_kkt_v3<KB><<<2 * sc.B, kThreads, smem, stream>>>(sc, bf, force ? 1 : 0);
}

// Block sizes instantiated: the shipped models with and without equality constraints (every instance unrolls two
// triangular solves of its size).  Other sizes fall back to the earlier generations.
bool launch_kkt_v3(int kb, const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  static const int gen = [] {
    const char* e = std::getenv("IDTO_KKT_GEN");
    return e ? std::atoi(e) : 3;
  }();
  if (gen != 3) return false;
  switch (kb) {
#define IDTO_V3_CASE(K) \
  case K: launch_v3_kb<K>(sc, bf, force, stream); return true;
    IDTO_V3_CASE(2) IDTO_V3_CASE(3) IDTO_V3_CASE(4) IDTO_V3_CASE(5) IDTO_V3_CASE(8) IDTO_V3_CASE(19)
    IDTO_V3_CASE(23) IDTO_V3_CASE(25) IDTO_V3_CASE(29)
#undef IDTO_V3_CASE
    default: return false;
  }
}

}  // namespace idto
