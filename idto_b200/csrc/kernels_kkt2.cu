// Two-sided block elimination of the time-major KKT system, second generation: register-resident,
// barrier-free LU inside a warp.
//
// Same mathematics as k_kkt_twisted (kernels_solve.cu; reference: PentaDiagonalFactorization,
// optimizer/penta_diagonal_solver.h:124-248, CalcLagrangeMultipliers cc:1371-1396, CalcDoglegPoint
// cc:2137-2140): CTA 0 of a 2-CTA cluster eliminates block rows 0..m top-down, CTA 1 rows N..m+1
// bottom-up, with the block-Thomas recurrence
//     K_i = B_i - A_i Y_{i-2}        G_i = C_i - A_i Z_{i-2} - K_i Y_{i-1}
//     G_i [Y_i | Z_i | r_i] = [D_i - K_i Z_{i-1} | E_i | b_i - A_i r_{i-2} - K_i r_{i-1}]
// (A, B = "back" blocks, D, E = "front" blocks of the sweep direction), an interface system for
// (x_m, x_{m+1}) and a concurrent back-substitution of both halves.
//
// What changed is how one block row is processed.  Measured on the first generation (B200, clock64):
// 53 % of the sweep was the pivot loop at ~1080 cycles per pivot step — an fp64 division (413 cycles
// dependent latency), four shared-memory round trips and a CTA barrier per step.  Here
//   * each of NLU "LU warps" keeps the WHOLE diagonal block G_i (lane = row, one register per column)
//     plus its share of the right-hand-side columns in registers and eliminates it redundantly: a
//     pivot step is warp-local (one REDUX on a key made of the high word of |x| and the lane index finds
//     the pivot row; the pivot row is broadcast through a few bytes of shared memory) — no CTA barrier;
//   * pivoting is implicit (rows are never swapped; the pivot order is carried per lane), which makes
//     a column update half an LDS.128 + 1 DFMA;
//   * every lane computes the reciprocal of its own candidate (MUFU.RCP64H + 2 Newton steps, ~65
//     cycles) while the search runs, so the division leaves the critical path;
//   * the other warps of the CTA prepare the next block row meanwhile (prefetch of its five blocks from
//     L2/HBM and the products with Y_{i-2}, Z_{i-2}, which do not depend on the row being eliminated).
// The pivot row is the one with the largest |x| judged on the exponent and the first 15 mantissa bits:
// partial pivoting up to a factor 1 + 2^-15, i.e. the same growth bound.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "kkt_view.cuh"
#include "reduce.cuh"
#include "solver.h"

#ifdef IDTO_KKT_TIMING
#define KT_DECL long long kt_t = clock64(), kt_acc[12] = {0};
#define KT(i) { const long long kt_n = clock64(); kt_acc[i] += kt_n - kt_t; kt_t = kt_n; }
#define KT_PRINT(b, dir)                                                                                   \
  if ((threadIdx.x == 0 || threadIdx.x == 128) && (b) == 0)                                                \
    printf("kkt2 dir %d tid %d: prologue %lld phaseA %lld load %lld lu %lld backsub %lld store %lld wait " \
           "%lld csync %lld iface_asm %lld iface_gj %lld csync2 %lld final %lld\n", dir, threadIdx.x,      \
           kt_acc[0], kt_acc[1], kt_acc[2], kt_acc[3], kt_acc[4], kt_acc[5], kt_acc[6], kt_acc[7],         \
           kt_acc[8], kt_acc[9], kt_acc[10], kt_acc[11]);
#else
#define KT_DECL
#define KT(i)
#define KT_PRINT(b, dir)
#endif

namespace idto {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kLuWarps = 4;   // warps that eliminate (each holds G and 1/kLuWarps of the right-hand sides)
constexpr int kThreads = 256;

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
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

// ---- warp-local LU ------------------------------------------------------------------------------------
// The KB x KB block is held as g[column][lane = row]; x[.][lane] are NC right-hand-side columns.
// Straight-line code for all KB pivot steps is ~10k instructions and runs at instruction-fetch speed
// (measured: 940 cycles per step), so the pivot loop is ROLLED: after every step the columns are rotated
// one register to the left (folded into the update: g[j-1] = g[j] - m * pivot_row[j]), which keeps the
// pivot column in g[0] for every step.  Finished columns (the U factor) go to shared memory `ub`.  The
// loop is cut into segments of kSeg steps so that the number of columns updated per step follows the
// shrinking trailing matrix, and it is software-pipelined: the search of step c+1 starts as soon as its
// column is updated and overlaps the remaining updates of step c.
constexpr int kSeg = 4;

struct LuState {
  int p;         // pivot lane of the current step
  double m;      // multiplier of this lane's row for the current step (0 for finished rows)
  bool done;     // this row has been a pivot row
  bool fail;
  int ord;       // step at which this row was the pivot row
  int src;       // lane k: pivot lane of step k
  double myinv;  // reciprocal of this row's pivot
};

__device__ __forceinline__ void lu_search(double xc, int c, int lane, double* ub, LuState& s) {
  // one REDUX finds the pivot row: magnitude (exponent + 15 mantissa bits of |x|) in the high 27 bits of the
  // key, 31 - lane in the low 5 (ties go to the lowest row); finished and padding rows carry key 0
  const unsigned key =
      s.done ? 0u : (((unsigned(__double2hiint(xc)) & 0x7fffffe0u) + 32u) | unsigned(31 - lane));
  const unsigned mx = __reduce_max_sync(kFull, key);
  const double rc = fast_rcp(xc);
  const int p = 31 - int(mx & 31u);
  s.fail |= (mx < 64u) | (mx >= 0x7ff00020u);  // zero (or subnormal) pivot column, or inf / NaN in it
  const double inv = __shfl_sync(kFull, rc, p);
  const bool is_p = lane == p;
  s.m = (s.done || is_p) ? 0.0 : xc * inv;
  if (is_p) s.done = true, s.ord = c, s.myinv = inv;
  if (lane == c) s.src = p;
  s.p = p;
  ub[c * 32 + lane] = xc;  // U[., c] in the rows chosen before step c (final); unused elsewhere
}

template <int KB, int NC, int C0>
__device__ __forceinline__ void lu_segment(double (&g)[KB], double (&x)[NC], int lane, double* ub, double* pr,
                                           LuState& s) {
  constexpr int W = KB - 1 - C0;  // live columns right of the pivot column at the start of the segment
  constexpr int C1 = C0 + kSeg < KB ? C0 + kSeg : KB;
  // The pivot row reaches the other rows on two data paths: the trailing columns of G through shuffles
  // (2 SHFL + 2 MOV per column), the right-hand-side columns through shared memory (lane p stores 16 bytes per
  // instruction, everybody loads them back as a broadcast).  Measured per pivot step, kb = 25: everything through
  // shuffles 480 cycles, everything through shared memory 468, this split 440.
  constexpr int NB = NC;
#pragma unroll 1
  for (int c = C0; c < C1; ++c) {
    const int p = s.p;
    const double m = s.m;
    double nxt = 0.0;
    if constexpr (W >= 1) nxt = fma(-m, __shfl_sync(kFull, g[1], p), g[1]);
    double* b = pr + (c & 1) * ((NB + 1) & ~1);
    if (lane == p) {
#pragma unroll
      for (int q = 0; q < NB; q += 2)
        *reinterpret_cast<double2*>(b + q) = make_double2(x[q], q + 1 < NB ? x[q + 1] : 0.0);
    }
    __syncwarp();
    if (c + 1 < KB) lu_search(nxt, c + 1, lane, ub, s);
#pragma unroll
    for (int j = 2; j <= W; ++j) g[j - 1] = fma(-m, __shfl_sync(kFull, g[j], p), g[j]);
#pragma unroll
    for (int q = 0; q < NB; q += 2) {
      const double2 v = *reinterpret_cast<const double2*>(b + q);
      x[q] = fma(-m, v.x, x[q]);
      if (q + 1 < NB) x[q + 1] = fma(-m, v.y, x[q + 1]);
    }
    g[0] = nxt;
  }
  if constexpr (C1 < KB) lu_segment<KB, NC, C1>(g, x, lane, ub, pr, s);
}

// LU with implicit partial pivoting of G (in place) applied to the right-hand sides x.  `ub` [KB][32] and
// `pr` [2][KB + NC] are this warp's scratch.
template <int KB, int NC>
__device__ __forceinline__ void warp_lu_eliminate(double (&g)[KB], double (&x)[NC], double* ub, double* pr, int lane,
                                                  LuState& s) {
  s.done = lane >= KB, s.fail = false, s.ord = KB - 1, s.src = lane, s.myinv = 0.0;
  lu_search(g[0], 0, lane, ub, s);
  lu_segment<KB, NC, 0>(g, x, lane, ub, pr, s);
}
// Row normalisation and back-substitution.  On return lane l holds, in x[j], component s.ord of column j of
// G^-1 X (the row chosen at step k holds x_k).
template <int KB, int NC>
__device__ __forceinline__ void warp_lu_backsub(double (&x)[NC], const double* ub, double* pr, int lane,
                                                const LuState& s) {
#pragma unroll
  for (int j = 0; j < NC; ++j) x[j] *= s.myinv;  // unit-diagonal U: every row scaled by 1 / its pivot
  __syncwarp();
  // the row chosen at step k (lane p_k) holds x_k once the steps > k are done
#pragma unroll 1
  for (int k = KB - 1; k >= 1; --k) {
    const int pk = __shfl_sync(kFull, s.src, k);
    const double u = (s.ord < k) ? ub[k * 32 + lane] * s.myinv : 0.0;  // rows chosen before step k carry U[., k]
    double* b = pr + (k & 1) * ((NC + 1) & ~1);
    if (lane == pk) {
#pragma unroll
      for (int q = 0; q < NC; q += 2)
        *reinterpret_cast<double2*>(b + q) = make_double2(x[q], q + 1 < NC ? x[q + 1] : 0.0);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NC; q += 2) {
      const double2 v = *reinterpret_cast<const double2*>(b + q);
      x[q] = fma(-u, v.x, x[q]);
      if (q + 1 < NC) x[q + 1] = fma(-u, v.y, x[q + 1]);
    }
  }
}

}  // namespace

template <int KB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    k_kkt_tw2(SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_fail;
  __shared__ int s_ord[64];
  const int b = blockIdx.x >> 1, dir = blockIdx.x & 1;
  if (!force && !bf.ctl[b].derivs_dirty) return;  // same decision in both CTAs of the cluster
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  constexpr int kb = KB, kk = KB * KB;
  constexpr int LD = (KB + 1) & ~1;   // even leading dimension: 16-byte aligned columns for LDS.128
  constexpr int kl = KB * LD;         // one padded column-major block
  // the 2*kb + 1 right-hand-side columns [Y | Z | r] are dealt round-robin to the LU warps
  constexpr int NC = (2 * KB + 1 + kLuWarps - 1) / kLuWarps;
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
  double* Kb = raw + 10 * kl;       // [kl]     K_i
  double* Mg = Kb + kl;             // [2][kl]  G_i (parity: written one row ahead)
  double* My = Mg + 2 * kl;         // [kl]     D_i - K_i Z_{i-1}
  double* Mz = My + kl;             // [kl]     E_i
  double* rb = Mz + kl;             // [2][LD]  r of the two previous rows
  double* rawb = rb + 2 * LD;       // [2][LD]  b_i
  double* rv = rawb + 2 * LD;       // [2][LD]  right-hand side of row i
  double* ubuf = rv + 2 * LD;       // [kLuWarps][KB][32] finished U columns of each LU warp
  double* prow = ubuf + kLuWarps * KB * 32;  // [kLuWarps][2][kPr] pivot-row broadcast scratch
  constexpr int kPr = (KB + NC + 1) & ~1;
  double* FY = bf.FY + size_t(b) * nblk * kk;
  double* FZ = bf.FZ + size_t(b) * nblk * kk;
  double* Fr = bf.X + size_t(b) * nblk * kb;
  double* xq = bf.pH + size_t(b) * sc.n;
  double* lam = bf.lambda + size_t(b) * sc.nh;
  double* xint = bf.tmp2 + size_t(b) * sc.n;  // interface solution (x_mid, x_mid+1): 2*kb doubles (n >= 2*kb)
  for (int e = tid; e < 19 * kl + 6 * LD; e += kThreads) sm[e] = 0.0;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  KT_DECL

  // Blocks of the row visited at step n -> raw[n & 1] (zeros where the matrix has no such block).  All
  // loads of a thread are issued before the first store (the element loop has a static trip count).
  auto load_row = [&](int n, int t0, auto nt_tag) {
    constexpr int NT = decltype(nt_tag)::value;
    constexpr int NE = (kk + NT - 1) / NT;
    if (n >= nsteps) return;
    const int i = first + sgn * n;
    double* dst = raw + (n & 1) * 5 * kl;
    const bool hb1 = n >= 1, hb2 = n >= 2;
    const bool hf1 = dir == 0 ? (i + 1 <= N) : (i - 1 >= 0);
    const bool hf2 = dir == 0 ? (i + 2 <= N) : (i - 2 >= 0);
    double v[NE][5];
#pragma unroll
    for (int k = 0; k < NE; ++k) {
      const int e = t0 + k * NT;
      const int c = e / kb, rr = e - c * kb;
      const bool ok = e < kk;
      v[k][0] = ok && hb2 ? kkt_blk(V, i, i - 2 * sgn, rr, c) : 0.0;
      v[k][1] = ok && hb1 ? kkt_blk(V, i, i - sgn, rr, c) : 0.0;
      v[k][2] = ok ? kkt_C(V, i, rr, c) : 0.0;
      v[k][3] = ok && hf1 ? kkt_blk(V, i, i + sgn, rr, c) : 0.0;
      v[k][4] = ok && hf2 ? kkt_blk(V, i, i + 2 * sgn, rr, c) : 0.0;
    }
    double bv = 0.0;
    if (t0 < kb) bv = t0 < nq ? -gs[i * nq + t0] : (i >= 1 ? -h[(i - 1) * sc.nu + (t0 - nq)] : 0.0);
#pragma unroll
    for (int k = 0; k < NE; ++k) {
      const int e = t0 + k * NT;
      const int c = e / kb, rr = e - c * kb, o = c * LD + rr;
      if (e < kk) {
#pragma unroll
        for (int q = 0; q < 5; ++q) dst[q * kl + o] = v[k][q];
      }
    }
    if (t0 < kb) rawb[(n & 1) * LD + t0] = bv;
  };
  // Products with the results of row n-2 (they do not depend on row n-1): K, G partial, r partial of
  // step n, by NG warps (gw = 0..NG-1); every warp runs its columns as independent FMA chains.
  // Y_{n-2} lives in Yb[n & 1].
  auto pre_row = [&](int n, int gw, auto ng_tag) {
    constexpr int NG = decltype(ng_tag)::value;
    constexpr int NTK = (KB + 1 + NG - 1) / NG;  // columns (incl. the right-hand side) per warp
    if (n >= nsteps) return;
    const double* rw = raw + (n & 1) * 5 * kl;
    const double* Yp = Yb + (n & 1) * kl;
    const double* Zp = Zb + (n & 1) * kl;
    const double* rp = rb + (n & 1) * LD;
    double ar[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) ar[j] = rw[j * LD + r];
    double ka[NTK], ga[NTK];
    const double *ys[NTK], *zs[NTK];
#pragma unroll
    for (int k = 0; k < NTK; ++k) {
      const int c = gw + NG * k;
      ka[k] = 0.0, ga[k] = 0.0;
      ys[k] = c < kb ? Yp + c * LD : rp;  // c == kb: the right-hand side (one chain); c > kb: idle
      zs[k] = c < kb ? Zp + c * LD : rp;
    }
#pragma unroll
    for (int j = 0; j < KB; j += 2) {
#pragma unroll
      for (int k = 0; k < NTK; ++k) {
        const double2 yv = *reinterpret_cast<const double2*>(ys[k] + j);
        const double2 zv = *reinterpret_cast<const double2*>(zs[k] + j);
        ka[k] = fma(ar[j], yv.x, ka[k]), ga[k] = fma(ar[j], zv.x, ga[k]);
        if (j + 1 < KB) ka[k] = fma(ar[j + 1], yv.y, ka[k]), ga[k] = fma(ar[j + 1], zv.y, ga[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < NTK; ++k) {
      const int c = gw + NG * k;
      if (row && c < kb) {
        Kb[c * LD + r] = rw[kl + c * LD + r] - ka[k];
        Mg[(n & 1) * kl + c * LD + r] = rw[2 * kl + c * LD + r] - ga[k];
      } else if (row && c == kb) {
        rv[(n & 1) * LD + r] = rawb[(n & 1) * LD + r] - ka[k];
      }
    }
  };
  using AllT = std::integral_constant<int, kThreads>;
  using HelpT = std::integral_constant<int, kThreads - 32 * kLuWarps>;
  using AllW = std::integral_constant<int, kThreads / 32>;
  using HelpW = std::integral_constant<int, kThreads / 32 - kLuWarps>;

  load_row(0, tid, AllT{});
  load_row(1, tid, AllT{});
  __syncthreads();
  pre_row(0, wid, AllW{});
  __syncthreads();
  KT(0)

  for (int n = 0; n < nsteps; ++n) {
    const int i = first + sgn * n, cur = n & 1, prv = cur ^ 1;
    // ---- phase A (all warps): products with the results of row n-1 ----------------------------------
    //   G -= K Y_{n-1};  Yrhs = D - K Z_{n-1};  r -= K r_{n-1};  E is copied so that raw[cur] is free
    {
      const double* rw = raw + cur * 5 * kl;
      const double* Yp = Yb + prv * kl;
      const double* Zp = Zb + prv * kl;
      const double* rp = rb + prv * LD;
      // out = K_i [Y_{n-1} | Z_{n-1} | r_{n-1}]  (kb x (2 kb + 1)) on the fp64 tensor cores: DMMA m8n8k4 tiles,
      // one 8-column tile of the right-hand side per warp, the K fragments of all (row tile, k tile) pairs in
      // registers.  The scalar version was bound by the shared-memory pipe (one broadcast load per multiply-add
      // column step: 2.7k cycles per block row); fragments are full-width loads.
      constexpr int NW = kThreads / 32;
      constexpr int MT = (KB + 7) / 8, KT4 = (KB + 3) / 4, NTN = (2 * KB + 1 + 7) / 8;
      const int grp = lane >> 2, tig = lane & 3;
      double afr[MT][KT4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int kt = 0; kt < KT4; ++kt) {
          const int rr = mt * 8 + grp, kc = kt * 4 + tig;
          afr[mt][kt] = (rr < kb && kc < kb) ? Kb[kc * LD + rr] : 0.0;
        }
      for (int nt = wid; nt < NTN; nt += NW) {
        double c0[MT], c1[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) c0[mt] = 0.0, c1[mt] = 0.0;
        const int nb = nt * 8 + grp;  // column of the flat right-hand side this lane feeds
        const double* bsrc = nb < kb ? Yp + nb * LD : (nb < 2 * kb ? Zp + (nb - kb) * LD : rp);
#pragma unroll
        for (int kt = 0; kt < KT4; ++kt) {
          const int kr = kt * 4 + tig;
          const double bfr = (kr < kb && nb <= 2 * kb) ? bsrc[kr] : 0.0;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[mt]), "+d"(c1[mt])
                         : "d"(afr[mt][kt]), "d"(bfr));
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int rr = mt * 8 + grp;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int t = nt * 8 + 2 * tig + h;
            const double acc = h ? c1[mt] : c0[mt];
            if (rr < kb) {
              if (t < kb) {
                Mg[cur * kl + t * LD + rr] -= acc;
              } else if (t < 2 * kb) {
                const int c = t - kb;
                My[c * LD + rr] = rw[3 * kl + c * LD + rr] - acc;
                Mz[c * LD + rr] = rw[4 * kl + c * LD + rr];
              } else if (t == 2 * kb) {
                rv[cur * LD + rr] -= acc;
              }
            }
          }
        }
      }
    }
    __syncthreads();
    KT(1)
    // ---- phase B -------------------------------------------------------------------------------------
    if (wid < kLuWarps) {
      // LU warps: G and the own right-hand-side columns (w, w + NLU, ...) into registers
      double g[KB], x[NC];
#pragma unroll
      for (int c = 0; c < KB; ++c) g[c] = row ? Mg[cur * kl + c * LD + r] : 0.0;
#pragma unroll
      for (int jj = 0; jj < NC; ++jj) {
        const int t = wid + kLuWarps * jj;  // flat column: Y (t < kb), Z (t < 2 kb), r (t == 2 kb), padding
        const double* src = t < kb ? My + t * LD : (t < 2 * kb ? Mz + (t - kb) * LD : rv + cur * LD);
        x[jj] = (row && t <= 2 * kb) ? src[r] : 0.0;
      }
      KT(2)
      LuState lus;
      warp_lu_eliminate<KB, NC>(g, x, ubuf + wid * KB * 32, prow + wid * 2 * kPr, lane, lus);
      KT(3)
      warp_lu_backsub<KB, NC>(x, ubuf + wid * KB * 32, prow + wid * 2 * kPr, lane, lus);
      KT(4)
      const int ro = lus.ord;  // this lane holds row `ro` of the solution
      if (lus.fail && lane == 0) s_fail = 1;
      // Y_i, Z_i, r_i: to shared memory for the next two rows and to HBM for the back-substitution
#pragma unroll
      for (int jj = 0; jj < NC; ++jj) {
        const int t = wid + kLuWarps * jj;
        if (row) {
          if (t < kb) {
            Yb[cur * kl + t * LD + ro] = x[jj];
            FY[size_t(i) * kk + t * kb + ro] = x[jj];
          } else if (t < 2 * kb) {
            Zb[cur * kl + (t - kb) * LD + ro] = x[jj];
            FZ[size_t(i) * kk + (t - kb) * kb + ro] = x[jj];
          } else if (t == 2 * kb) {
            rb[cur * LD + ro] = x[jj];
            Fr[size_t(i) * kb + ro] = x[jj];
          }
        }
      }
      KT(5)
    } else {
      // helper warps: prefetch the blocks of row n+2 (raw[cur] is free since phase A) and prepare row
      // n+1 from the results of row n-1
      KT(2)
      pre_row(n + 1, wid - kLuWarps, HelpW{});
      KT(3)
      load_row(n + 2, tid - 32 * kLuWarps, HelpT{});
      KT(5)
    }
    __syncthreads();
    KT(6)
  }
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
    // stage Y, Z, r of the four rows around the interface (independent, coalesced loads: one L2 round trip; the
    // element-wise version chased 25 dependent global loads per entry: 22k cycles)
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
      Sr[e] = ok ? Fr[size_t(rowi) * kb + e % kb] : 0.0;
    }
    __syncthreads();
    const double *Ym1 = S, *Zm1 = S + kk, *Ym = S + 2 * kk, *Zm = S + 3 * kk, *Yp1 = S + 4 * kk, *Zp1 = S + 5 * kk,
                 *Yp2 = S + 6 * kk, *Zp2 = S + 7 * kk;
    for (int e = tid; e < n2 * W2; e += kThreads) {
      const int c = e / n2, rr = e % n2;
      const int br = rr / kb, r1 = rr % kb;
      // row block 0: x_m row uses Z_m against the bottom chain's row m+2; row block 1: Z'_{m+1} against row m-1
      const double* Zl = br == 0 ? Zm : Zp1;
      double val;
      if (c < n2) {
        const int bc = c / kb, cc = c % kb;
        const double* Rt;  // right factor (column cc) and the leading term
        if (br == 0 && bc == 0) val = (r1 == cc) ? 1.0 : 0.0, Rt = Zp2;          // I - Z_m Z'_{m+2}
        else if (br == 0) val = Ym[cc * kb + r1], Rt = Yp2;                      // Y_m - Z_m Y'_{m+2}
        else if (bc == 0) val = Yp1[cc * kb + r1], Rt = Ym1;                     // Y'_{m+1} - Z'_{m+1} Y_{m-1}
        else val = (r1 == cc) ? 1.0 : 0.0, Rt = Zm1;                             // I - Z'_{m+1} Z_{m-1}
        double v2 = 0.0;  // two chains; fully unrolled so that the loads run ahead of the multiply-adds
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
    // Gauss-Jordan with implicit partial pivoting on the n2 x (n2+1) system, ONE barrier per step: every
    // warp finds the pivot row redundantly (two candidate rows per lane, one REDUX on a key that carries
    // the row index), thread (row, column group) updates its entries; the pivot row is never modified and
    // is normalised at the end.
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
      // row r was the pivot row of step s_ord[r]: x[s_ord[r]] = rhs[r] / pivot
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
  // Warp 0 owns the recurrence (lane = row); the other warps stream Y_i, Z_i, r_i from L2 into a ring of
  // three shared-memory slots, two rows ahead of it.
  __syncthreads();
  constexpr int kSlot = 2 * kk + LD;
  double* ring = sm;              // [3][Y | Z | r]
  double* xs = sm + 3 * kSlot;    // [2][LD] the two most recent solution blocks
  const int istart = dir == 0 ? mid - 1 : mid + 2, iend = dir == 0 ? -1 : N + 1;
  const int nrows = dir == 0 ? mid : N - mid - 1;
  // cp.async (LDGSTS): the loader threads do not wait for their loads, the ring slot of row it+2 fills while
  // rows it and it+1 are consumed (with plain loads every row paid an L2 round trip before its barrier)
  auto fetch = [&](int it, int t0, int nt) {
    if (it < nrows) {
      const int i = istart - sgn * it;
      double* dst = ring + (it % 3) * kSlot;
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
  fetch(0, tid, kThreads);
  fetch(1, tid, kThreads);
  if (wid == 0) {
    // x1 = most recent block, x2 = the one before: (x_mid, x_mid+1) for the top half, reversed below
    const double u0 = row ? xint[lane] : 0.0, u1 = row ? xint[kb + lane] : 0.0;
    if (lane < LD) xs[lane] = dir == 0 ? u0 : u1, xs[LD + lane] = dir == 0 ? u1 : u0;
    emit(dir == 0 ? mid : mid + 1, dir == 0 ? u0 : u1);
  }
  asm volatile("cp.async.wait_group 1;" ::: "memory");  // row 0 has landed (row 1 may be in flight)
  __syncthreads();
  for (int it = 0; it < nrows; ++it) {
    if (wid == 0) {
      const double* Ys = ring + (it % 3) * kSlot;
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
    fetch(it + 2, tid, kThreads);                          // slot (it + 2) % 3 was consumed in iteration it - 1
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // row it + 1 has landed
    __syncthreads();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  KT(11)
  KT_PRINT(b, dir)
}

template <int KB>
static void launch_tw2_kb(const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  constexpr int LD = (KB + 1) & ~1, kl = KB * LD;
  const int sweep = 19 * kl + 6 * LD + kLuWarps * KB * 32 + kLuWarps * 2 * (KB + (2 * KB + 1 + kLuWarps - 1) / kLuWarps + 1), tail = std::max(2 * KB * (2 * KB + 1) + 2 + 8 * KB * KB + 4 * KB, 3 * (2 * KB * KB + LD) + 2 * LD);
  const int smem = std::max(sweep, tail) * 8;
  static bool attr_set[kMaxDevices] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(k_kkt_tw2<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  }
  k_kkt_tw2<KB><<<2 * sc.B, kThreads, smem, stream>>>(sc, bf, force ? 1 : 0);
}

// Block sizes instantiated for the register-resident sweep (every unrolled instance is ~10k instructions):
// the shipped models with and without equality constraints.  Other sizes use k_kkt_twisted.
bool launch_kkt_tw2(int kb, const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  static const bool off = std::getenv("IDTO_KKT_V1") != nullptr;
  if (off) return false;
  switch (kb) {
#define IDTO_TW2_CASE(K) \
  case K: launch_tw2_kb<K>(sc, bf, force, stream); return true;
    IDTO_TW2_CASE(2) IDTO_TW2_CASE(3) IDTO_TW2_CASE(4) IDTO_TW2_CASE(5) IDTO_TW2_CASE(8) IDTO_TW2_CASE(19)
    IDTO_TW2_CASE(23) IDTO_TW2_CASE(25) IDTO_TW2_CASE(29)
#undef IDTO_TW2_CASE
    default: return false;
  }
}

}  // namespace idto
