// C ABI of the idto_b200 CUDA layer (include/idto_b200.h) and the host-side driver that enqueues
// the kernel family.  There is NO CPU fallback: without a CUDA device every entry point that needs
// one returns IDTO_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "dynamics.cuh"
#include "solver.h"

namespace idto {

std::atomic<long> g_launch_counter{0};
bool use_chain_kernels(const DevModel& dm) {
  static const bool force_group = [] {
    const char* e = std::getenv("IDTO_DYNAMICS");
    return e && std::string(e) == "group";
  }();
  return !force_group && chain_supported(dm);
}
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

struct DevAlloc {
  std::vector<void*> ptrs;
  template <typename T>
  cudaError_t get(T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    e = cudaMemset(q, 0, std::max<size_t>(count, 1) * sizeof(T));
    ptrs.push_back(q);
    *p = static_cast<T*>(q);
    return e;
  }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
  }
};

}  // namespace idto

using namespace idto;

struct idto_model_s {
  DevModel dm;
  DevAlloc mem;
  int device;
  std::vector<int> unactuated, quat_starts;
  std::vector<int> jtype, qs, vs;
  int nq, nv;
};

struct ProfEvent {
  cudaEvent_t a, b;
};

struct idto_solver_s {
  idto_model_t model;
  SolverConsts sc;
  SolverBufs bf;
  DevAlloc mem;
  cudaStream_t stream = nullptr;
  int device;
  idto_params params;
  // mutable per-batch problem data
  double *q_init, *v_init, *q_nom, *v_nom;
  double* mpc_in = nullptr;  // staging of idto_mpc_advance inputs: [B] elapsed | [B][nq] q0 | [B][nv] v0 | [nq] selector
  int* iters_dev = nullptr;    // [B] ProbCtl::iters packed for the asynchronous D2H of idto_resolve_async
  double* delta_in = nullptr;  // [B] staging of idto_set_delta (bf.red holds the cached dogleg scalars: not scratch)
  long launches0 = 0;
  bool profile = false;
  std::map<std::string, std::vector<ProfEvent>> prof;
  std::vector<ProbCtl> ctl_host;
  int* status_host = nullptr;  // pinned mirror of bf.status: the D2H of idto_synchronize stays asynchronous
  size_t stats_cap = 0;
  // sub-batches on their own streams: the latency-bound per-problem kernels (KKT sweep, TR scalars)
  // of one sub-batch overlap with the throughput-bound kernels (ID partials, assembly) of another
  int nsub = 1;
  std::vector<cudaStream_t> sub_streams;
  std::vector<cudaEvent_t> sub_done;
  cudaEvent_t ev_start = nullptr;
  bool main_dirty = false, subs_dirty = false;
  // CUDA graph of one idto_resolve_async call (stream captured once per argument set, replayed afterwards): an
  // MPC loop issues the same ~25 copies / launches / event operations every step
  struct GraphKey {
    int iters = -1;
    size_t stats_cap = 0;
    const void* p[14] = {};  // host pointers of the call (10 of the re-solve, 4 of a leading MPC advance)
    bool operator==(const GraphKey& o) const {
      return iters == o.iters && stats_cap == o.stats_cap && std::memcmp(p, o.p, sizeof(p)) == 0;
    }
  };
  GraphKey gkey;
  cudaGraphExec_t gexec = nullptr;
  long glaunches = 0;
  int gcalls = 0;  // calls with the current key (the first one runs eagerly: first-use attribute calls, allocations)
  bool use_graph = true;
};

namespace {

int check_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    set_last_error("no CUDA device available; idto_b200 has no CPU fallback");
    return IDTO_ERR_NO_DEVICE;
  }
  return IDTO_OK;
}

bool is_diag(const double* Mx, int n) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i)
      if (i != j && Mx[size_t(j) * n + i] != 0.0) return false;
  return true;
}

struct Prof {
  idto_solver_s* s;
  const char* name;
  ProfEvent ev{};
  Prof(idto_solver_s* s_, const char* n) : s(s_), name(n) {
    if (s->profile) {
      cudaEventCreate(&ev.a), cudaEventCreate(&ev.b);
      cudaEventRecord(ev.a, s->stream);
    }
  }
  ~Prof() {
    if (s->profile) {
      cudaEventRecord(ev.b, s->stream);
      s->prof[name].push_back(ev);
    }
  }
};

// Stage pipeline; each stage is gated on the per-problem dirty flags unless force is set.
void enqueue_trajectory(idto_solver_s* s, bool scratch, bool force) {
  Prof p(s, scratch ? "trajectory_scratch" : "trajectory");
  if (s->bf.act_base && !scratch)  // debug trace: evaluate every problem so that the whole trace is rewritten
    cudaMemsetAsync(s->bf.act_base, 0, size_t(s->sc.B) * s->sc.T * std::max(s->model->dm.np, 1) * sizeof(int), s->stream),
        force = true;
  launch_traj(s->model->dm, s->sc, s->bf, scratch, force, s->stream);
  launch_tau(s->model->dm, s->sc, s->bf, scratch, force, s->stream);
}
void enqueue_derivatives(idto_solver_s* s, bool force) {
  Prof p(s, "id_partials");
  if (s->bf.act_fd) {  // debug trace: 0xff.. = -1 = "pair not visited by this evaluation"; pruned models report the
    // pairs of their compacted list only, so their trace starts from 0
    cudaMemsetAsync(s->bf.act_fd, s->model->dm.prune ? 0 : 0xff,
                    size_t(s->sc.B) * s->sc.T * s->sc.nq * 4 * std::max(s->model->dm.np, 1) * sizeof(int), s->stream);
    force = true;
  }
  launch_partials(s->model->dm, s->sc, s->bf, force, s->stream);
}
void enqueue_assembly(idto_solver_s* s, bool force, bool with_gm = true) {
  {
    Prof p(s, "assemble");
    launch_assemble(s->model->dm, s->sc, s->bf, force, s->stream);
  }
  {
    Prof p(s, "factor");
    launch_factor(s->sc, s->bf, force, s->stream);
  }
  {
    Prof p(s, "lagrange");
    launch_lagrange(s->model->dm, s->sc, s->bf, force, s->stream, with_gm);
  }
}
// Dogleg point and the model terms of the trust ratio; then the scalar tail of the iteration.
void enqueue_dogleg(const SolverConsts& sc, const SolverBufs& bf, cudaStream_t st, bool gm_done) {
  (void)gm_done;
  launch_dogleg(sc, bf, st);
}
void enqueue_trust(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool commit, cudaStream_t st) {
  if (trust_final_enabled())
    launch_trust_final(dm, sc, bf, commit, st);
  else
    launch_trust_update(dm, sc, bf, commit, st);
}
void enqueue_iteration(idto_solver_s* s) {
  const bool split = true;
  enqueue_trajectory(s, false, false);
  enqueue_derivatives(s, false);
  enqueue_assembly(s, false, split);
  if (s->sc.check_convergence) launch_conv_check(s->sc, s->bf, s->stream);
  {
    Prof p(s, "dogleg");
    enqueue_dogleg(s->sc, s->bf, s->stream, split);
  }
  enqueue_trajectory(s, true, true);
  {
    Prof p(s, "trust_update");
    enqueue_trust(s->model->dm, s->sc, s->bf, true, s->stream);
  }
}

// View of the solver restricted to problems [b0, b0+nb): every buffer is batch-major.
void make_view(const idto_solver_s* s, int b0, int nb, SolverConsts* scv, SolverBufs* bfv) {
  const SolverConsts& c = s->sc;
  *scv = c;
  scv->B = nb;
  SolverBufs v = s->bf;
  const size_t T = c.T, T1 = c.T + 1, nq = c.nq, nv = c.nv, n = c.n, nh = std::max(c.nh, 1), o = size_t(b0);
  const size_t kb = nq + c.nu, nuq = size_t(std::max(c.nu, 1)) * nq;
  for (TrajBuf* tb : {&v.st, &v.sc}) {
    tb->q += o * T1 * nq, tb->v += o * T1 * nv, tb->a += o * T * nv, tb->tau += o * T * nv;
    tb->Nplus += o * T1 * nv * nq, tb->cost += o, tb->h += o * nh;
    if (tb->near) tb->near += o * T * near_stride(s->model->dm.nact);
  }
  v.q_init += o * nq, v.v_init += o * nv, v.q_nom += o * T1 * nq, v.v_nom += o * T1 * nv;
  v.dqm += o * T * nv * nq, v.dqt += o * T * nv * nq, v.dqp += o * T * nv * nq;
  v.g += o * n, v.D += o * n, v.gs += o * n, v.gm += o * n, v.lambda += o * nh, v.merit += o;
  for (double** p : {&v.HA, &v.HB, &v.HC, &v.SA, &v.SB, &v.SC}) *p += o * T1 * nq * nq;
  v.Jm += o * T * nuq, v.Jt += o * T * nuq, v.Jp += o * T * nuq;
  v.FY += o * T1 * kb * kb, v.FZ += o * T1 * kb * kb, v.X += o * T1 * kb, v.rhs += o * nh;
  if (v.S) v.S += o * n * (3 * nq);
  if (v.crw) v.crw += cr_workspace_doubles(b0, c.T, int(nq) + (c.eq ? c.nu : 0));
  v.pH += o * n, v.dq += o * n, v.dqH += o * n, v.tmp1 += o * n, v.tmp2 += o * n, v.red += o * 8, v.part += o * T1 * 4, v.cnt += o;
  if (v.stash) v.stash += o * T * size_t(s->model->dm.nb) * 48;
  v.ctl += o;
  if (v.stats) v.stats += o * size_t(v.stats_cap) * IDTO_NUM_STATS;
  *bfv = v;
}

// ---- sub-batch streams ------------------------------------------------------------------------------
// With nsub > 1 the batch is cut into contiguous sub-batches, each with its own stream.  A sub-batch's
// whole iteration sequence lives on its stream, so consecutive re-solves of different sub-batches
// pipeline: the sequential KKT sweep of one sub-batch (one CTA per problem, latency bound) overlaps with
// the ID-partials / assembly kernels of the others.  The caller's stream only forks (after host copies)
// and joins (before anything the host reads).
struct Part {
  SolverConsts sc;
  SolverBufs bf;
  cudaStream_t st;
  int b0;
};
bool multi(const idto_solver_s* s) { return s->nsub > 1 && !s->profile && !s->bf.act_base; }
// Device-visible alias of a pinned (page-locked, mapped) host buffer; nullptr for pageable memory, which has to go
// through cudaMemcpyAsync.  IDTO_NO_ZEROCOPY=1 keeps every transfer on the copy engines.
template <class T>
T* mapped_alias(T* host) {
  static const bool off = std::getenv("IDTO_NO_ZEROCOPY") != nullptr;
  if (!host || off) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
  return static_cast<T*>(a.devicePointer);
}
std::vector<Part> make_parts(const idto_solver_s* s) {
  std::vector<Part> parts(s->nsub);
  for (int i = 0; i < s->nsub; ++i) {
    const int b0 = int(size_t(s->sc.B) * i / s->nsub), b1 = int(size_t(s->sc.B) * (i + 1) / s->nsub);
    make_view(s, b0, b1 - b0, &parts[i].sc, &parts[i].bf);
    parts[i].st = s->sub_streams[i];
    parts[i].b0 = b0;
  }
  return parts;
}
// Before enqueuing on the caller's stream: join the sub-streams.
void use_main(idto_solver_s* s) {
  if (s->subs_dirty) {
    for (int i = 0; i < s->nsub; ++i) {
      cudaEventRecord(s->sub_done[i], s->sub_streams[i]);
      cudaStreamWaitEvent(s->stream, s->sub_done[i], 0);
    }
    s->subs_dirty = false;
  }
  s->main_dirty = true;
}
// Before enqueuing on the sub-streams: they wait for what is already on the caller's stream.
void use_subs(idto_solver_s* s) {
  if (s->main_dirty) {
    cudaEventRecord(s->ev_start, s->stream);
    for (int i = 0; i < s->nsub; ++i) cudaStreamWaitEvent(s->sub_streams[i], s->ev_start, 0);
    s->main_dirty = false;
  }
  s->subs_dirty = true;
}

__global__ void k_set_ctl(ProbCtl* ctl, int B, int what, double Delta0);
__global__ void k_set_prev_cost(ProbCtl* ctl, const double* cost, int B);

// `count` trust-region iterations on the sub-streams, stage-major so that earlier sub-batches get the
// SMs first at every stage (their KKT sweeps then start while later sub-batches still differentiate).
void enqueue_iterations_parts(idto_solver_s* s, const std::vector<Part>& P, int count) {
  const DevModel& dm = s->model->dm;
  for (int k = 0; k < count; ++k) {
    for (const Part& p : P) {
      launch_traj(dm, p.sc, p.bf, false, false, p.st);
      launch_tau(dm, p.sc, p.bf, false, false, p.st);
    }
    for (const Part& p : P) launch_partials(dm, p.sc, p.bf, false, p.st);
    for (const Part& p : P) launch_assemble(dm, p.sc, p.bf, false, p.st);
    const bool split = true;
    for (const Part& p : P) launch_lagrange(dm, p.sc, p.bf, false, p.st, split);
    for (const Part& p : P) {
      if (s->sc.check_convergence) launch_conv_check(p.sc, p.bf, p.st);
      enqueue_dogleg(p.sc, p.bf, p.st, split);
      launch_traj(dm, p.sc, p.bf, true, true, p.st);
      launch_tau(dm, p.sc, p.bf, true, true, p.st);
      enqueue_trust(dm, p.sc, p.bf, true, p.st);
    }
  }
}

__global__ void k_set_ctl(ProbCtl* ctl, int B, int what, double Delta0) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (what == 0) {  // creation
    ctl[b].Delta = Delta0, ctl[b].traj_dirty = 1, ctl[b].derivs_dirty = 1, ctl[b].active = 1;
    ctl[b].iters = 0, ctl[b].reason = 0, ctl[b].pending = 0, ctl[b].tr_active = 0;
  } else if (what == 1) {  // q / problem changed: invalidate all caches (state.h:333-350)
    ctl[b].traj_dirty = 1, ctl[b].derivs_dirty = 1, ctl[b].pending = 0;
  } else if (what == 2) {  // solve start
    ctl[b].active = 1, ctl[b].iters = 0, ctl[b].reason = 0, ctl[b].pending = 0;
  }
}
__global__ void k_set_prev_cost(ProbCtl* ctl, const double* cost, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) ctl[b].prev_cost = cost[b];
}
__global__ void k_set_delta(ProbCtl* ctl, const double* d, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) ctl[b].Delta = d[b];
}
__global__ void k_pack_iters(const ProbCtl* ctl, int* out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = ctl[b].iters;
}
// Results of a re-solve written straight into the caller's PINNED host buffers (mapped into the device's address
// space): one launch per sub-batch instead of four or five copy-engine transfers of ~200 KB, each with its own
// fixed cost (48 -> ~28 us per step of 64 quadruped problems).  Outputs that are not pinned go through cudaMemcpyAsync.
__global__ void k_export(const double* __restrict__ q, const double* __restrict__ v, const double* __restrict__ tau,
                         const double* __restrict__ stats, const ProbCtl* __restrict__ ctl, double* __restrict__ qh,
                         double* __restrict__ vh, double* __restrict__ tauh, double* __restrict__ sth,
                         int* __restrict__ ith, size_t nq_tot, size_t nv_tot, size_t ntau_tot, int B, int iters,
                         size_t stats_cap) {
  const size_t stride = size_t(gridDim.x) * blockDim.x, t0 = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (qh)
    for (size_t i = t0; i < nq_tot; i += stride) qh[i] = q[i];
  if (vh)
    for (size_t i = t0; i < nv_tot; i += stride) vh[i] = v[i];
  if (tauh)
    for (size_t i = t0; i < ntau_tot; i += stride) tauh[i] = tau[i];
  if (sth) {
    const size_t rowlen = size_t(iters) * IDTO_NUM_STATS;
    for (size_t i = t0; i < size_t(B) * rowlen; i += stride) {
      const size_t b = i / rowlen, r = i - b * rowlen;
      sth[i] = stats[b * stats_cap * IDTO_NUM_STATS + r];
    }
  }
  if (ith)
    for (size_t i = t0; i < size_t(B); i += stride) ith[i] = ctl[i].iters;
}
__global__ void k_fill(double* p, size_t n, double v) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) p[i] = v;
}
// p[b*stride + i] = v for i < n, one CTA per b
__global__ void k_fill_strided(double* p, size_t n, size_t stride, double v) {
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) p[blockIdx.x * stride + i] = v;
}
// Constant part of N+ (identity pattern for 1-dof / planar joints and floating translations).
__global__ void k_init_nplus(DevModel dm, double* Np, int B, int T) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * (T + 1)) return;
  double* N = Np + size_t(idx) * dm.nv * dm.nq;
  const int* jtype = dm.itab + dm.o_jtype;
  const int* qs = dm.itab + dm.o_qs;
  const int* vs = dm.itab + dm.o_vs;
  for (int k = 0; k < dm.nb; ++k) {
    if (jtype[k] == IDTO_JOINT_QUAT_FLOATING) {
      for (int i = 0; i < 3; ++i) N[size_t(qs[k] + 4 + i) * dm.nv + vs[k] + 3 + i] = 1.0;
    } else {
      const int n = jtype[k] == IDTO_JOINT_PLANAR ? 3 : 1;
      for (int i = 0; i < n; ++i) N[size_t(qs[k] + i) * dm.nv + vs[k] + i] = 1.0;
    }
  }
}

int ensure_stats(idto_solver_s* s, size_t cap) {
  if (cap <= s->stats_cap) return IDTO_OK;
  double* p = nullptr;
  IDTO_CUDA_CHECK(cudaMalloc(&p, cap * s->sc.B * IDTO_NUM_STATS * sizeof(double)));
  s->mem.ptrs.push_back(p);
  s->bf.stats = p;
  s->bf.stats_cap = int(cap);
  s->stats_cap = cap;
  return IDTO_OK;
}

int check_status(idto_solver_s* s) {
  use_main(s);
  if (!s->status_host) IDTO_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&s->status_host), sizeof(int), cudaHostAllocDefault));
  IDTO_CUDA_CHECK(cudaMemcpyAsync(s->status_host, s->bf.status, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  IDTO_CUDA_CHECK(cudaStreamSynchronize(s->stream));
  const int st = *s->status_host;
  if (st != 0) {
    if (st == IDTO_ERR_CONTACT_OVERFLOW)
      set_last_error("more than " + std::to_string(s->model->dm.nact) +
                     " contact pairs within the activation distance in one inverse-dynamics evaluation (raise "
                     "IDTO_MAX_ACTIVE_PAIRS before creating the model)");
    else if (st == IDTO_ERR_UNSUPPORTED)
      set_last_error("no KKT sweep kernel for this block size");
    else
      set_last_error("penta-diagonal factorisation failed (singular diagonal block)");
    int zero = 0;
    cudaMemcpyAsync(s->bf.status, &zero, sizeof(int), cudaMemcpyHostToDevice, s->stream);
    return st;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error(std::string("kernel launch: ") + cudaGetErrorString(e));
    return IDTO_ERR_CUDA;
  }
  return IDTO_OK;
}

}  // namespace

extern "C" {

const char* idto_last_error(void) { return g_last_error.c_str(); }

int idto_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int idto_set_device(int device) {
  if (int rc = check_device()) return rc;
  if (device < 0 || device >= idto_device_count()) return IDTO_ERR_INVALID_ARG;
  IDTO_CUDA_CHECK(cudaSetDevice(device));
  return IDTO_OK;
}

void idto_params_default(idto_params* p) {  // optimizer/solver_parameters.h:64-167
  std::memset(p, 0, sizeof(*p));
  p->max_iterations = 100, p->gradients_method = IDTO_GRAD_FORWARD, p->normalize_quaternions = 0;
  p->contact_stiffness = 100, p->dissipation_velocity = 0.1, p->stiction_velocity = 0.05;
  p->friction_coefficient = 0.5, p->smoothing_factor = 0.1, p->scaling = 1;
  p->scaling_method = IDTO_SCALING_DOUBLE_SQRT, p->equality_constraints = 1;
  p->Delta0 = 1e-1, p->Delta_max = 1e5, p->check_convergence = 0, p->linear_solver = IDTO_LINSOLVE_TWISTED;
}

int idto_model_create(const idto_model_desc* d, idto_model_t* out) {
  if (!d || !out) return IDTO_ERR_INVALID_ARG;
  if (int rc = check_device()) return rc;
  if (d->nbodies < 1 || d->nbodies > kMaxGroup) {
    set_last_error("model has " + std::to_string(d->nbodies) + " moving bodies; supported: 1.." +
                   std::to_string(kMaxGroup));
    return IDTO_ERR_UNSUPPORTED;
  }
  const int nb = d->nbodies, ng = d->ngeoms, np = d->npairs;
  for (int i = 0; i < ng; ++i)
    if (d->geom_type[i] < IDTO_GEOM_SPHERE || d->geom_type[i] > IDTO_GEOM_HALF_SPACE) {
      set_last_error("unknown geometry type (sphere, box, capsule, cylinder, half space are supported)");
      return IDTO_ERR_UNSUPPORTED;
    }
  for (int i = 0; i < np; ++i)
    if (d->geom_type[d->pair_geomA[i]] != IDTO_GEOM_SPHERE && d->geom_type[d->pair_geomB[i]] != IDTO_GEOM_SPHERE) {
      set_last_error("contact pairs without a sphere on one side have no closed-form signed distance here");
      return IDTO_ERR_UNSUPPORTED;
    }
  auto* m = new idto_model_s();
  cudaGetDevice(&m->device);
  int G = 2;
  while (G < nb) G *= 2;
  const int nbp = G, npp = std::max(2, (np + 1) / 2 * 2), ngp = std::max(ng, 1);
  // levels / children
  std::vector<int> level(nb), nchild(nb, 0), child(size_t(kMaxChildren) * nbp, 0);
  int nlevels = 0;
  for (int k = 0; k < nb; ++k) {
    const int p = d->parent[k];
    if (p >= k) {
      delete m;
      set_last_error("bodies must be topologically ordered");
      return IDTO_ERR_INVALID_ARG;
    }
    level[k] = p < 0 ? 0 : level[p] + 1;
    nlevels = std::max(nlevels, level[k] + 1);
    if (p >= 0) {
      if (nchild[p] >= kMaxChildren) {
        delete m;
        set_last_error("more than kMaxChildren children on one body");
        return IDTO_ERR_UNSUPPORTED;
      }
      child[size_t(nchild[p]) * nbp + p] = k;
      nchild[p]++;
    }
  }
  // int table
  std::vector<int> it;
  auto push_i = [&](const int* src, int n, int padded) {
    const int off = int(it.size());
    for (int i = 0; i < padded; ++i) it.push_back(i < n ? src[i] : 0);
    return off;
  };
  DevModel& dm = m->dm;
  dm.nb = nb, dm.nbp = nbp, dm.nq = d->nq, dm.nv = d->nv, dm.ng = ngp, dm.np = np, dm.npp = npp;
  dm.prune = np > kMaxActivePairs ? 1 : 0;
  dm.nact = npp;  // (pruned models: set below, once the column split is known)
  {
    double chain = 0.0, off = 0.0;
    for (int k = 0; k < nb; ++k) {
      const double* p = d->X_PF + size_t(k) * 12 + 9;
      chain += std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    }
    for (int gi = 0; gi < ng; ++gi) {
      const double* p = d->X_BG + size_t(gi) * 12 + 9;
      off = std::max(off, std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));
    }
    dm.reach = chain + off;
  }
  dm.nlevels = nlevels, dm.group = G;
  dm.gx = d->gravity[0], dm.gy = d->gravity[1], dm.gz = d->gravity[2];
  std::vector<int> parent_p(nbp, -1);
  for (int k = 0; k < nb; ++k) parent_p[k] = d->parent[k];
  dm.o_parent = push_i(parent_p.data(), nbp, nbp);
  dm.o_jtype = push_i(d->joint_type, nb, nbp);
  dm.o_qs = push_i(d->q_start, nb, nbp);
  dm.o_vs = push_i(d->v_start, nb, nbp);
  dm.o_level = push_i(level.data(), nb, nbp);
  dm.o_nchild = push_i(nchild.data(), nb, nbp);
  dm.o_child = push_i(child.data(), kMaxChildren * nbp, kMaxChildren * nbp);
  std::vector<int> flags(nbp, 0), qowner(d->nq, 0);
  for (int k = 0; k < nb; ++k) {
    const double* R = d->R_MB + 9 * k;
    bool ident = true;
    for (int e = 0; e < 9; ++e) ident &= (R[e] == ((e % 4 == 0) ? 1.0 : 0.0));
    flags[k] = ident ? 1 : 0;
    const int jt = d->joint_type[k];
    const int nqb = jt == IDTO_JOINT_QUAT_FLOATING ? 7 : (jt == IDTO_JOINT_PLANAR ? 3 : 1);
    for (int i = 0; i < nqb; ++i) qowner[d->q_start[k] + i] = k;
    if (jt == IDTO_JOINT_QUAT_FLOATING) m->quat_starts.push_back(d->q_start[k]);
  }
  dm.o_flags = push_i(flags.data(), nbp, nbp);
  dm.o_qowner = push_i(qowner.data(), d->nq, d->nq);
  dm.o_gbody = push_i(d->geom_body, ng, ngp);
  dm.o_gtype = push_i(d->geom_type, ng, ngp);
  dm.o_pA = push_i(d->pair_geomA, np, npp);
  dm.o_pB = push_i(d->pair_geomB, np, npp);
  // ---- chain decomposition: a body's first child continues its chain, the others start new chains ----
  {
    std::vector<int> bchain(nb, -1), chain_top;
    int nchains = 0;
    for (int k = 0; k < nb; ++k) {  // bodies are in depth-first order: parents first
      const int p = d->parent[k];
      const bool continues = p >= 0 && child[size_t(0) * nbp + p] == k;  // first child of its parent
      if (continues) {
        bchain[k] = bchain[p];
      } else {
        bchain[k] = nchains++;
        chain_top.push_back(k);
      }
    }
    int CG = 1;
    while (CG < nchains) CG *= 2;
    dm.cgroup = CG, dm.nchains = nchains;
    dm.cg_res = 4;  // measured best on B200 (residues 4 and 12 spread the stride-3 leg bodies over all banks)
    if (const char* e = std::getenv("IDTO_CHAIN_RES")) dm.cg_res = std::atoi(e) & 15;
    std::vector<int> levbody(size_t(kMaxLevels) * CG, -1), levcross(kMaxLevels, 0), plane(nb, -1);
    if (nlevels > kMaxLevels || CG > 32) {
      delete m;
      set_last_error("tree too deep / too many chains for the chain-lane kernels");
      return IDTO_ERR_UNSUPPORTED;
    }
    for (int k = 0; k < nb; ++k) {
      levbody[size_t(level[k]) * CG + bchain[k]] = k;
      const int p = d->parent[k];
      if (p >= 0) {
        plane[k] = bchain[p];
        if (bchain[p] != bchain[k]) levcross[level[k]] = 1;
      }
    }
    std::vector<int> gslot(nb, -1), gdyn(ngp, -1);
    int ngb = 0, ngd = 0;
    for (int gi = 0; gi < ng; ++gi) {
      const int bdy = d->geom_body[gi];
      if (bdy >= 0) {
        gdyn[gi] = ngd++;
        if (gslot[bdy] < 0) gslot[bdy] = ngb++;
      }
    }
    dm.ngb = ngb, dm.ngd = ngd;
    dm.chain_ok = (CG <= 8 && nlevels <= kMaxLevels) ? 1 : 0;
    for (int k = 0; k < nb; ++k)  // multi-dof joints must hang off the world (R_WF == R_PF in the backward pass)
      if ((d->joint_type[k] == IDTO_JOINT_PLANAR || d->joint_type[k] == IDTO_JOINT_QUAT_FLOATING) && d->parent[k] >= 0)
        dm.chain_ok = 0;
    // column split (see DevModel::npath)
    std::vector<int> pathcols, fullcols;
    {
      const bool enable = dm.chain_ok && !(std::getenv("IDTO_PATH_COLS") && std::atoi(std::getenv("IDTO_PATH_COLS")) == 0);
      for (int i = 0; i < d->nq; ++i) {
        const int k = qowner[i];
        bool ok = enable && (d->joint_type[k] == IDTO_JOINT_REVOLUTE || d->joint_type[k] == IDTO_JOINT_PRISMATIC);
        std::vector<char> in_sub(nb, 0);
        int ndown = 0;
        for (int bdy = k; ok;) {  // the subtree must be a chain of 1-dof joints, at most 4 bodies long
          in_sub[bdy] = 1;
          ++ndown;
          const int jt = d->joint_type[bdy];
          if (jt != IDTO_JOINT_REVOLUTE && jt != IDTO_JOINT_PRISMATIC) ok = false;
          if (nchild[bdy] > 1 || ndown > 4) ok = false;
          if (nchild[bdy] == 0) break;
          bdy = child[size_t(0) * nbp + bdy];
        }
        int nup = 0;
        for (int a = d->parent[k]; a >= 0; a = d->parent[a]) ++nup;
        if (nup > 7) ok = false;
        for (int ip = 0; ip < np && ok; ++ip) {  // contact partners of subtree bodies must be world-anchored
          const int bA = d->geom_body[d->pair_geomA[ip]], bB = d->geom_body[d->pair_geomB[ip]];
          const bool sA = bA >= 0 && in_sub[bA], sB = bB >= 0 && in_sub[bB];
          if ((sA && bB >= 0) || (sB && bA >= 0)) ok = false;
        }
        (ok ? pathcols : fullcols).push_back(i);
      }
    }
    dm.npath = int(pathcols.size()), dm.nfull = int(fullcols.size());
    dm.o_pathcols = push_i(pathcols.data(), dm.npath, std::max(dm.npath, 1));
    dm.o_fullcols = push_i(fullcols.data(), dm.nfull, std::max(dm.nfull, 1));
    dm.o_levbody = push_i(levbody.data(), kMaxLevels * CG, kMaxLevels * CG);
    dm.o_levcross = push_i(levcross.data(), kMaxLevels, kMaxLevels);
    dm.o_plane = push_i(plane.data(), nb, nbp);
    dm.o_gslot = push_i(gslot.data(), nb, nbp);
    dm.o_gdyn = push_i(gdyn.data(), ngp, ngp);
    dm.o_bchain = push_i(bchain.data(), nb, nbp);
  }
  while (it.size() % 4) it.push_back(0);
  // double table (SoA: field-major, body-minor)
  std::vector<double> dt;
  auto push_soa = [&](const double* src, int fields, int n, int stride) {
    const int off = int(dt.size());
    for (int f = 0; f < fields; ++f)
      for (int i = 0; i < stride; ++i) dt.push_back(i < n ? src[size_t(i) * fields + f] : 0.0);
    return off;
  };
  dm.o_XPF = push_soa(d->X_PF, 12, nb, nbp);
  dm.o_RMB = push_soa(d->R_MB, 9, nb, nbp);
  dm.o_axis = push_soa(d->axis, 3, nb, nbp);
  dm.o_mass = push_soa(d->mass, 1, nb, nbp);
  dm.o_com = push_soa(d->com, 3, nb, nbp);
  dm.o_inertia = push_soa(d->inertia, 6, nb, nbp);
  dm.o_damping = push_soa(d->damping, 1, d->nv, d->nv);
  dm.o_gdims = push_soa(d->geom_dims, 3, ng, ngp);
  dm.o_XBG = push_soa(d->X_BG, 12, ng, ngp);
  dm.o_XWGs = push_soa(d->X_BG, 12, ng, ngp);  // world-anchored geometries: X_WG == X_BG (body = world)
  while (dt.size() % 2) dt.push_back(0.0);
  dm.itab_bytes = int(it.size() * sizeof(int));
  dm.dtab_bytes = int(dt.size() * sizeof(double));
  if (dm.prune) {  // capacity of the per-evaluation active list (and of the near lists): even, at most npp
    if (const char* e = std::getenv("IDTO_MAX_ACTIVE_PAIRS"))
      dm.nact = std::min(npp, (std::max(8, std::atoi(e)) + 1) / 2 * 2);
    else
      dm.nact = chain_fit_pair_slots(dm, d->nv);
  }
  int* di = nullptr;
  double* dd = nullptr;
  if (m->mem.get(&di, it.size()) != cudaSuccess || m->mem.get(&dd, dt.size()) != cudaSuccess ||
      cudaMemcpy(di, it.data(), dm.itab_bytes, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(dd, dt.data(), dm.dtab_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_last_error("cudaMalloc/cudaMemcpy failed for the model tables");
    m->mem.release();
    delete m;
    return IDTO_ERR_CUDA;
  }
  dm.itab = di, dm.dtab = dd;
  // cc:63-72: unactuated dofs (fully actuated if the actuation matrix is empty)
  bool any = false;
  for (int i = 0; i < d->nv; ++i) any |= d->actuated[i] != 0;
  if (any)
    for (int i = 0; i < d->nv; ++i)
      if (!d->actuated[i]) m->unactuated.push_back(i);
  m->nq = d->nq, m->nv = d->nv;
  m->jtype.assign(d->joint_type, d->joint_type + nb);
  m->qs.assign(d->q_start, d->q_start + nb);
  m->vs.assign(d->v_start, d->v_start + nb);
  *out = m;
  return IDTO_OK;
}

int idto_model_destroy(idto_model_t m) {
  if (!m) return IDTO_ERR_INVALID_ARG;
  m->mem.release();
  delete m;
  return IDTO_OK;
}
int idto_model_num_unactuated(idto_model_t m) { return m ? int(m->unactuated.size()) : IDTO_ERR_INVALID_ARG; }
int idto_model_unactuated_dofs(idto_model_t m, int* out) {
  if (!m || !out) return IDTO_ERR_INVALID_ARG;
  std::copy(m->unactuated.begin(), m->unactuated.end(), out);
  return IDTO_OK;
}

int idto_solver_create(idto_model_t m, const idto_problem_desc* pd, const idto_params* p, int batch,
                       idto_solver_t* out) {
  if (!m || !pd || !p || !out || batch < 1 || pd->num_steps < 1 || !(pd->time_step > 0)) return IDTO_ERR_INVALID_ARG;
  if (int rc = check_device()) return rc;
  const int nq = m->nq, nv = m->nv, T = pd->num_steps, B = batch;
  const bool dense_w = !is_diag(pd->Qq, nq) || !is_diag(pd->Qv, nv) || !is_diag(pd->Qf_q, nq) ||
                       !is_diag(pd->Qf_v, nv) || !is_diag(pd->R, nv);
  if (dense_w && !use_chain_kernels(m->dm)) {
    set_last_error("dense (non-diagonal) cost weights need the chain-lane inverse-dynamics kernels");
    return IDTO_ERR_UNSUPPORTED;
  }
  if (p->linear_solver < IDTO_LINSOLVE_THOMAS || p->linear_solver > IDTO_LINSOLVE_CYCLIC_REDUCTION) return IDTO_ERR_INVALID_ARG;
  if (p->gradients_method < IDTO_GRAD_FORWARD || p->gradients_method > IDTO_GRAD_CENTRAL4) {
    set_last_error("gradients_method must be forward, central or central4 (autodiff needs Drake scalars)");
    return IDTO_ERR_UNSUPPORTED;
  }
  {
    // the KKT sweeps and the row-parallel mat-vecs keep one block row per warp (lane = row): block size <= 32
    const int nu = int(m->unactuated.size());
    const int kb = nq + ((p->equality_constraints && nu > 0) ? nu : 0);
    if (nq > 32 || kb > 32) {
      set_last_error("KKT block size nq + num_unactuated = " + std::to_string(kb) + " (nq = " + std::to_string(nq) +
                     "): the penta-diagonal kernels support block sizes up to 32");
      return IDTO_ERR_UNSUPPORTED;
    }
  }
  {
    const bool chain = use_chain_kernels(m->dm);
    if (m->dm.prune && !chain) {
      set_last_error("models with more than " + std::to_string(kMaxActivePairs) +
                     " candidate contact pairs need the chain-lane inverse-dynamics kernels");
      return IDTO_ERR_UNSUPPORTED;
    }
    const int need = chain ? chain_min_smem_bytes(m->dm, nv, p->gradients_method) : partials_smem_bytes(m->dm, nq);
    if (need > 226 * 1024) {
      set_last_error("model needs " + std::to_string(need / 1024) + " KB of shared memory per inverse-dynamics CTA (" +
                     std::to_string(m->dm.np) + " candidate contact pairs, " + std::to_string(m->dm.nact) +
                     " pair slots per evaluation): more than an SM has");
      return IDTO_ERR_UNSUPPORTED;
    }
  }
  auto* s = new idto_solver_s();
  s->model = m, s->params = *p;
  cudaSetDevice(m->device);  // the workspace lives on the device that holds the baked tables
  s->device = m->device;
  SolverConsts& sc = s->sc;
  sc.B = B, sc.T = T, sc.nq = nq, sc.nv = nv, sc.nu = int(m->unactuated.size());
  sc.n = (T + 1) * nq, sc.nh = sc.nu * T, sc.dt = pd->time_step;
  sc.method = p->gradients_method, sc.scaling = p->scaling, sc.scaling_method = p->scaling_method;
  sc.eq = p->equality_constraints && sc.nh > 0, sc.normalize_quat = p->normalize_quaternions;
  sc.check_convergence = p->check_convergence;
  sc.linear_solver = p->linear_solver;
  sc.k = p->contact_stiffness, sc.sigma = p->smoothing_factor, sc.vd = p->dissipation_velocity;
  sc.vs = p->stiction_velocity, sc.mu = p->friction_coefficient;
  const double eps = std::sqrt(std::numeric_limits<double>::epsilon());
  sc.threshold = -sc.sigma * std::log(std::exp(eps / (sc.sigma * sc.k)) - 1.0);  // cc:268-269
  sc.Delta_max = p->Delta_max;
  sc.tol[0] = p->tol_rel_cost_reduction, sc.tol[1] = p->tol_abs_cost_reduction;
  sc.tol[2] = p->tol_rel_gradient_along_dq, sc.tol[3] = p->tol_abs_gradient_along_dq;
  sc.tol[4] = p->tol_rel_state_change, sc.tol[5] = p->tol_abs_state_change;
  sc.nquat = int(m->quat_starts.size());

  DevAlloc& A = s->mem;
  SolverBufs& bf = s->bf;
  bool ok = true;
  auto alloc = [&](double** ptr, size_t n) { ok = ok && (A.get(ptr, n) == cudaSuccess); };
  double *dQq, *dQv, *dQfq, *dQfv, *dR;
  alloc(&dQq, nq), alloc(&dQv, nv), alloc(&dQfq, nq), alloc(&dQfv, nv), alloc(&dR, nv);
  int *dun = nullptr, *dqs = nullptr;
  ok = ok && A.get(&dun, m->unactuated.size()) == cudaSuccess && A.get(&dqs, m->quat_starts.size()) == cudaSuccess;
  const size_t nTq = size_t(B) * (T + 1) * nq, nTv = size_t(B) * (T + 1) * nv, nA = size_t(B) * T * nv;
  const size_t nN = size_t(B) * (T + 1) * nv * nq, nP = size_t(B) * T * nv * nq, nH = size_t(B) * (T + 1) * nq * nq;
  const size_t nh = std::max(sc.nh, 1), nvar = size_t(B) * sc.n;
  for (TrajBuf* tb : {&bf.st, &bf.sc}) {
    alloc(&tb->q, nTq), alloc(&tb->v, nTv), alloc(&tb->a, nA), alloc(&tb->tau, nA), alloc(&tb->Nplus, nN);
    alloc(&tb->cost, B), alloc(&tb->h, size_t(B) * nh);
    tb->near = nullptr;
    if (m->dm.prune) ok = ok && A.get(&tb->near, size_t(B) * T * near_stride(m->dm.nact)) == cudaSuccess;
  }
  alloc(&s->q_init, size_t(B) * nq), alloc(&s->v_init, size_t(B) * nv), alloc(&s->q_nom, nTq), alloc(&s->v_nom, nTv);
  alloc(&bf.dqm, nP), alloc(&bf.dqt, nP), alloc(&bf.dqp, nP);
  alloc(&bf.g, nvar), alloc(&bf.D, nvar), alloc(&bf.gs, nvar), alloc(&bf.gm, nvar);
  alloc(&bf.lambda, size_t(B) * nh), alloc(&bf.merit, B);
  alloc(&bf.HA, nH), alloc(&bf.HB, nH), alloc(&bf.HC, nH), alloc(&bf.SA, nH), alloc(&bf.SB, nH), alloc(&bf.SC, nH);
  const size_t nJ = size_t(B) * T * std::max(sc.nu, 1) * nq;
  alloc(&bf.Jm, nJ), alloc(&bf.Jt, nJ), alloc(&bf.Jp, nJ);
  const size_t kbm = size_t(nq) + sc.nu;  // KKT block size
  bf.FK = bf.FG = bf.S = nullptr;
  if (p->linear_solver == IDTO_LINSOLVE_DENSE_LDLT) alloc(&bf.S, size_t(B) * sc.n * (3 * nq));  // band storage of H~
  bf.crw = nullptr;
  if (p->linear_solver == IDTO_LINSOLVE_CYCLIC_REDUCTION)
    alloc(&bf.crw, cr_workspace_doubles(B, sc.T, nq + (sc.eq ? sc.nu : 0)));
  alloc(&bf.FY, size_t(B) * (T + 1) * kbm * kbm), alloc(&bf.FZ, size_t(B) * (T + 1) * kbm * kbm);
  alloc(&bf.X, size_t(B) * (T + 1) * kbm);
  alloc(&bf.rhs, size_t(B) * nh);
  alloc(&bf.pH, nvar), alloc(&bf.dq, nvar), alloc(&bf.dqH, nvar), alloc(&bf.tmp1, nvar), alloc(&bf.tmp2, nvar);
  alloc(&bf.red, size_t(B) * 8);
  alloc(&bf.part, size_t(B) * (T + 1) * 4);
  bf.stash = nullptr;
  bf.stash_half = size_t(B) * T * m->dm.nb * 48;
  if (m->dm.npath > 0) alloc(&bf.stash, 2 * bf.stash_half);
  alloc(&s->mpc_in, size_t(B) * (1 + nq + nv) + nq);
  alloc(&s->delta_in, B);
  ok = ok && A.get(&s->iters_dev, B) == cudaSuccess;
  ok = ok && A.get(&bf.cnt, B) == cudaSuccess;
  ok = ok && A.get(&bf.ctl, B) == cudaSuccess && A.get(&bf.status, 1) == cudaSuccess;
  if (!ok) {
    set_last_error("cudaMalloc failed while creating the solver workspace");
    A.release();
    delete s;
    return IDTO_ERR_CUDA;
  }
  bf.q_init = s->q_init, bf.v_init = s->v_init, bf.q_nom = s->q_nom, bf.v_nom = s->v_nom;
  bf.stats = nullptr, bf.stats_cap = 0;
  bf.act_base = bf.act_fd = nullptr;
  {
    std::vector<double> cp(T + 1, 0.0);
    if (T >= 2) cp[2] = 0.25;
    for (int j = 3; j <= T - 2; ++j) cp[j] = 1.0 / (4.0 - cp[j - 1]);
    double* dcp = nullptr;
    if (A.get(&dcp, cp.size()) == cudaSuccess) cudaMemcpy(dcp, cp.data(), cp.size() * sizeof(double), cudaMemcpyHostToDevice);
    bf.spline_cp = dcp;
  }
  sc.Qq = dQq, sc.Qv = dQv, sc.Qfq = dQfq, sc.Qfv = dQfv, sc.R = dR, sc.unact = dun, sc.quat_starts = dqs;
  sc.status = bf.status;
  auto up_diag = [&](double* dst, const double* Mx, int n) {
    std::vector<double> dg(n);
    for (int i = 0; i < n; ++i) dg[i] = Mx[size_t(i) * n + i];
    cudaMemcpy(dst, dg.data(), n * sizeof(double), cudaMemcpyHostToDevice);
  };
  up_diag(dQq, pd->Qq, nq), up_diag(dQv, pd->Qv, nv), up_diag(dQfq, pd->Qf_q, nq), up_diag(dQfv, pd->Qf_v, nv);
  up_diag(dR, pd->R, nv);
  sc.dense_w = dense_w ? 1 : 0;
  sc.QqM = sc.QvM = sc.QfqM = sc.QfvM = sc.RM = nullptr;
  if (dense_w) {
    auto up_full = [&](const double** dst, const double* Mx, int n) {
      double* p = nullptr;
      if (A.get(&p, size_t(n) * n) == cudaSuccess) cudaMemcpy(p, Mx, size_t(n) * n * sizeof(double), cudaMemcpyHostToDevice);
      *dst = p;
    };
    up_full(&sc.QqM, pd->Qq, nq), up_full(&sc.QvM, pd->Qv, nv), up_full(&sc.QfqM, pd->Qf_q, nq);
    up_full(&sc.QfvM, pd->Qf_v, nv), up_full(&sc.RM, pd->R, nv);
  }
  cudaMemcpy(dun, m->unactuated.data(), m->unactuated.size() * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(dqs, m->quat_starts.data(), m->quat_starts.size() * sizeof(int), cudaMemcpyHostToDevice);
  // broadcast the template problem to every batch element
  for (int b = 0; b < B; ++b) {
    cudaMemcpy(s->q_init + size_t(b) * nq, pd->q_init, nq * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(s->v_init + size_t(b) * nv, pd->v_init, nv * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(s->q_nom + size_t(b) * (T + 1) * nq, pd->q_nom, size_t(T + 1) * nq * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(s->v_nom + size_t(b) * (T + 1) * nv, pd->v_nom, size_t(T + 1) * nv * sizeof(double), cudaMemcpyHostToDevice);
  }
  // preset entries (inverse_dynamics_partials.h:35-42; state.h:68)
  const size_t blk = size_t(nv) * nq;
  k_fill_strided<<<B, 128>>>(bf.dqm, blk, size_t(T) * blk, std::numeric_limits<double>::quiet_NaN());
  k_fill<<<64, 256>>>(bf.D, nvar, 1.0);
  k_init_nplus<<<(B * (T + 1) + 127) / 128, 128>>>(m->dm, bf.st.Nplus, B, T);
  k_init_nplus<<<(B * (T + 1) + 127) / 128, 128>>>(m->dm, bf.sc.Nplus, B, T);
  k_set_ctl<<<(B + 127) / 128, 128>>>(bf.ctl, B, 0, p->Delta0);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    set_last_error("solver initialisation kernels failed");
    A.release();
    delete s;
    return IDTO_ERR_CUDA;
  }
  s->ctl_host.resize(B);
  if (const char* e = std::getenv("IDTO_GRAPH")) s->use_graph = std::atoi(e) != 0;
  s->launches0 = g_launch_counter;
  cudaEventCreateWithFlags(&s->ev_start, cudaEventDisableTiming);
  {
    int nsub = B >= 32 ? 2 : 1;  // measured on B200 (64 problems): e2e 64.4k / 65.7k / 64.0k / 63.0k iters/s for 1 / 2 / 3 / 4
    if (const char* e = std::getenv("IDTO_SUBSTREAMS")) nsub = std::max(1, std::atoi(e));
    idto_solver_set_substreams(s, nsub);
  }
  *out = s;
  return IDTO_OK;
}

int idto_solver_destroy(idto_solver_t s) {
  if (!s) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  use_main(s);
  cudaDeviceSynchronize();
  if (s->gexec) cudaGraphExecDestroy(s->gexec);
  for (auto st : s->sub_streams) cudaStreamDestroy(st);
  for (auto ev : s->sub_done) cudaEventDestroy(ev);
  if (s->ev_start) cudaEventDestroy(s->ev_start);
  if (s->status_host) cudaFreeHost(s->status_host);
  for (auto& kv : s->prof)
    for (auto& e : kv.second) cudaEventDestroy(e.a), cudaEventDestroy(e.b);
  s->mem.release();
  delete s;
  return IDTO_OK;
}

int idto_solver_set_substreams(idto_solver_t s, int n) {
  if (!s || n < 1 || n > 16) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  if (n > s->sc.B) n = s->sc.B;
  use_main(s);
  cudaStreamSynchronize(s->stream);
  if (s->gexec) cudaGraphExecDestroy(s->gexec);  // the graph holds the old streams' fork / join structure
  s->gexec = nullptr, s->gcalls = 0, s->gkey = idto_solver_s::GraphKey();
  for (auto st : s->sub_streams) cudaStreamDestroy(st);
  for (auto ev : s->sub_done) cudaEventDestroy(ev);
  s->sub_streams.clear(), s->sub_done.clear();
  s->nsub = n;
  if (n > 1)
    for (int i = 0; i < n; ++i) {
      cudaStream_t st;
      cudaEvent_t ev;
      IDTO_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      IDTO_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      s->sub_streams.push_back(st), s->sub_done.push_back(ev);
    }
  return IDTO_OK;
}

int idto_solver_set_stream(idto_solver_t s, void* stream) {
  if (!s) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  s->stream = static_cast<cudaStream_t>(stream);
  if (s->gexec) cudaGraphExecDestroy(s->gexec);
  s->gexec = nullptr, s->gcalls = 0, s->gkey = idto_solver_s::GraphKey();
  return IDTO_OK;
}

int idto_set_q(idto_solver_t s, const double* q) {
  if (!s || !q) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  use_main(s);
  IDTO_CUDA_CHECK(cudaMemcpyAsync(s->bf.st.q, q, size_t(s->sc.B) * s->sc.n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  k_set_ctl<<<(s->sc.B + 127) / 128, 128, 0, s->stream>>>(s->bf.ctl, s->sc.B, 1, 0.0);
  return IDTO_OK;
}
int idto_invalidate(idto_solver_t s) {
  if (!s) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  if (multi(s)) {
    use_subs(s);
    for (const Part& p : make_parts(s)) k_set_ctl<<<(p.sc.B + 127) / 128, 128, 0, p.st>>>(p.bf.ctl, p.sc.B, 1, 0.0);
    return IDTO_OK;
  }
  use_main(s);
  k_set_ctl<<<(s->sc.B + 127) / 128, 128, 0, s->stream>>>(s->bf.ctl, s->sc.B, 1, 0.0);
  return IDTO_OK;
}
int idto_reset_initial_conditions(idto_solver_t s, const double* q0, const double* v0) {
  if (!s || !q0 || !v0) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  use_main(s);
  IDTO_CUDA_CHECK(cudaMemcpyAsync(s->q_init, q0, size_t(s->sc.B) * s->sc.nq * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  IDTO_CUDA_CHECK(cudaMemcpyAsync(s->v_init, v0, size_t(s->sc.B) * s->sc.nv * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  k_set_ctl<<<(s->sc.B + 127) / 128, 128, 0, s->stream>>>(s->bf.ctl, s->sc.B, 1, 0.0);
  return IDTO_OK;
}
int idto_update_nominal_trajectory(idto_solver_t s, const double* qn, const double* vn) {
  if (!s || !qn || !vn) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  use_main(s);
  const size_t T1 = s->sc.T + 1;
  IDTO_CUDA_CHECK(cudaMemcpyAsync(s->q_nom, qn, size_t(s->sc.B) * T1 * s->sc.nq * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  IDTO_CUDA_CHECK(cudaMemcpyAsync(s->v_nom, vn, size_t(s->sc.B) * T1 * s->sc.nv * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  k_set_ctl<<<(s->sc.B + 127) / 128, 128, 0, s->stream>>>(s->bf.ctl, s->sc.B, 1, 0.0);
  return IDTO_OK;
}
int idto_set_delta(idto_solver_t s, const double* delta) {
  if (!s || !delta) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  use_main(s);
  IDTO_CUDA_CHECK(cudaMemcpyAsync(s->delta_in, delta, s->sc.B * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  k_set_delta<<<(s->sc.B + 127) / 128, 128, 0, s->stream>>>(s->bf.ctl, s->delta_in, s->sc.B);
  return IDTO_OK;
}
int idto_get_delta(idto_solver_t s, double* delta) { return idto_get(s, "delta", delta); }

int idto_eval_trajectory(idto_solver_t s) {
  if (!s) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  use_main(s);
  enqueue_trajectory(s, false, false);
  return check_status(s);
}
int idto_eval_derivatives(idto_solver_t s) {
  if (!s) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  use_main(s);
  enqueue_trajectory(s, false, false);
  enqueue_derivatives(s, false);
  return check_status(s);
}
int idto_eval_assembly(idto_solver_t s) {
  if (!s) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  use_main(s);
  enqueue_trajectory(s, false, false);
  enqueue_derivatives(s, false);
  enqueue_assembly(s, false);
  launch_clear_dirty(s->sc, s->bf, s->stream);
  return check_status(s);
}
int idto_eval_dogleg(idto_solver_t s) {
  if (!s) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  if (int rc = idto_eval_assembly(s)) return rc;
  k_set_ctl<<<(s->sc.B + 127) / 128, 128, 0, s->stream>>>(s->bf.ctl, s->sc.B, 2, 0.0);
  enqueue_dogleg(s->sc, s->bf, s->stream, true);
  return check_status(s);
}
int idto_eval_trust_ratio(idto_solver_t s) {
  if (!s) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  if (int rc = idto_eval_dogleg(s)) return rc;
  enqueue_trajectory(s, true, true);
  enqueue_trust(s->model->dm, s->sc, s->bf, false, s->stream);
  return check_status(s);
}

long idto_field_size(idto_solver_t s, const char* field) {
  if (!s || !field) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  const SolverConsts& c = s->sc;
  const std::string f(field);
  const long T = c.T, nq = c.nq, nv = c.nv;
  if (f == "q" || f == "q_nom") return (T + 1) * nq;
  if (f == "v" || f == "v_nom") return (T + 1) * nv;
  if (f == "a" || f == "tau") return T * nv;
  if (f == "Nplus") return (T + 1) * nv * nq;
  if (f == "cost" || f == "merit" || f == "dq_active" || f == "rho" || f == "delta") return 1;
  if (f == "h" || f == "lambda") return c.nh;
  if (f == "dtau_dqm" || f == "dtau_dqt" || f == "dtau_dqp") return T * nv * nq;
  if (f == "g" || f == "D" || f == "gs" || f == "gm" || f == "dq" || f == "dqH") return c.n;
  if (f == "H_A" || f == "H_B" || f == "H_C" || f == "Hs_A" || f == "Hs_B" || f == "Hs_C") return (T + 1) * nq * nq;
  if (f == "J") return long(c.nh) * c.n;
  if (f == "dvt_dqt" || f == "dvt_dqm") return (T + 1) * nv * nq;
  if (f == "pair_active") return s->bf.act_base ? T * long(s->model->dm.np) : IDTO_ERR_INVALID_ARG;
  if (f == "pair_active_fd") return s->bf.act_fd ? T * nq * 4 * long(s->model->dm.np) : IDTO_ERR_INVALID_ARG;
  {  // debugging views of the KKT sweep (block size kb = nq + nu when equality constraints are on)
    const long kb = nq + (c.eq ? c.nu : 0);
    if (f == "dbg_FY" || f == "dbg_FZ") return (T + 1) * kb * kb;
    if (f == "dbg_Fr") return (T + 1) * kb;
  }
  return IDTO_ERR_INVALID_ARG;
}

int idto_get(idto_solver_t s, const char* field, double* out) {
  if (!s || !field || !out) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  const long sz = idto_field_size(s, field);
  if (sz < 0) return IDTO_ERR_INVALID_ARG;
  const SolverConsts& c = s->sc;
  const SolverBufs& bf = s->bf;
  const std::string f(field);
  use_main(s);
  IDTO_CUDA_CHECK(cudaStreamSynchronize(s->stream));
  const double* src = nullptr;
  if (f == "q") src = bf.st.q;
  else if (f == "q_nom") src = bf.q_nom;
  else if (f == "v_nom") src = bf.v_nom;
  else if (f == "v") src = bf.st.v;
  else if (f == "a") src = bf.st.a;
  else if (f == "tau") src = bf.st.tau;
  else if (f == "Nplus") src = bf.st.Nplus;
  else if (f == "cost") src = bf.st.cost;
  else if (f == "h") src = bf.st.h;
  else if (f == "dtau_dqm") src = bf.dqm;
  else if (f == "dtau_dqt") src = bf.dqt;
  else if (f == "dtau_dqp") src = bf.dqp;
  else if (f == "g") src = bf.g;
  else if (f == "H_A") src = bf.HA;
  else if (f == "H_B") src = bf.HB;
  else if (f == "H_C") src = bf.HC;
  else if (f == "D") src = bf.D;
  else if (f == "Hs_A") src = bf.SA;
  else if (f == "Hs_B") src = bf.SB;
  else if (f == "Hs_C") src = bf.SC;
  else if (f == "gs") src = bf.gs;
  else if (f == "lambda") src = bf.lambda;
  else if (f == "merit") src = bf.merit;
  else if (f == "gm") src = bf.gm;
  else if (f == "dq") src = bf.dq;
  else if (f == "dqH") src = bf.dqH;
  else if (f == "dbg_FY") src = bf.FY;
  else if (f == "dbg_FZ") src = bf.FZ;
  else if (f == "dbg_Fr") src = bf.X;
  if (src) {
    IDTO_CUDA_CHECK(cudaMemcpy(out, src, size_t(sz) * c.B * sizeof(double), cudaMemcpyDeviceToHost));
    return IDTO_OK;
  }
  if (f == "dvt_dqt" || f == "dvt_dqm") {  // VelocityPartials (velocity_partials.h:19-39, cc:962-973): +-N+_t / dt
    IDTO_CUDA_CHECK(cudaMemcpy(out, bf.st.Nplus, size_t(sz) * c.B * sizeof(double), cudaMemcpyDeviceToHost));
    const bool qm = f == "dvt_dqm";
    const size_t blk = size_t(c.nv) * c.nq;
    for (int b = 0; b < c.B; ++b)
      for (int t = 0; t <= c.T; ++t) {
        double* N = out + (size_t(b) * (c.T + 1) + t) * blk;
        for (size_t e = 0; e < blk; ++e)
          N[e] = qm ? (t == 0 ? std::numeric_limits<double>::quiet_NaN() : -N[e] / c.dt) : N[e] / c.dt;
      }
    return IDTO_OK;
  }
  if (f == "pair_active" || f == "pair_active_fd") {
    std::vector<int> tmp(size_t(sz) * c.B);
    IDTO_CUDA_CHECK(cudaMemcpy(tmp.data(), f == "pair_active" ? bf.act_base : bf.act_fd, tmp.size() * sizeof(int),
                               cudaMemcpyDeviceToHost));
    for (size_t e = 0; e < tmp.size(); ++e) out[e] = double(tmp[e]);
    return IDTO_OK;
  }
  if (f == "J") {  // expand the three bands into the reference's dense (nu*T) x n, column-major
    const size_t nJ = size_t(c.B) * c.T * std::max(c.nu, 1) * c.nq;
    std::vector<double> jm(nJ), jt(nJ), jp(nJ);
    IDTO_CUDA_CHECK(cudaMemcpy(jm.data(), bf.Jm, nJ * sizeof(double), cudaMemcpyDeviceToHost));
    IDTO_CUDA_CHECK(cudaMemcpy(jt.data(), bf.Jt, nJ * sizeof(double), cudaMemcpyDeviceToHost));
    IDTO_CUDA_CHECK(cudaMemcpy(jp.data(), bf.Jp, nJ * sizeof(double), cudaMemcpyDeviceToHost));
    std::fill(out, out + size_t(sz) * c.B, 0.0);
    for (int b = 0; b < c.B; ++b)
      for (int t = 0; t < c.T; ++t)
        for (int u = 0; u < c.nu; ++u) {
          const size_t jb = ((size_t(b) * c.T + t) * c.nu + u) * c.nq;
          const size_t row = size_t(t) * c.nu + u;
          double* J = out + size_t(b) * sz;
          for (int cc = 0; cc < c.nq; ++cc) {
            J[(size_t(t + 1) * c.nq + cc) * c.nh + row] = jp[jb + cc];
            if (t > 0) J[(size_t(t) * c.nq + cc) * c.nh + row] = jt[jb + cc];
            if (t > 1) J[(size_t(t - 1) * c.nq + cc) * c.nh + row] = jm[jb + cc];
          }
        }
    return IDTO_OK;
  }
  IDTO_CUDA_CHECK(cudaMemcpy(s->ctl_host.data(), bf.ctl, c.B * sizeof(ProbCtl), cudaMemcpyDeviceToHost));
  for (int b = 0; b < c.B; ++b) {
    const ProbCtl& k = s->ctl_host[b];
    out[b] = f == "rho" ? k.rho : (f == "delta" ? k.Delta : double(k.tr_active));
  }
  return IDTO_OK;
}

static int solve_enqueue(idto_solver_t s, int max_iterations) {
  if (int rc = ensure_stats(s, size_t(max_iterations))) return rc;
  if (multi(s)) {
    use_subs(s);
    const std::vector<Part> P = make_parts(s);
    const DevModel& dm = s->model->dm;
    for (const Part& p : P) {
      k_set_ctl<<<(p.sc.B + 127) / 128, 128, 0, p.st>>>(p.bf.ctl, p.sc.B, 2, 0.0);
      if (s->sc.check_convergence) {  // previous_cost = EvalCost(state) (cc:2494)
        launch_traj(dm, p.sc, p.bf, false, false, p.st);
        launch_tau(dm, p.sc, p.bf, false, false, p.st);
        k_set_prev_cost<<<(p.sc.B + 127) / 128, 128, 0, p.st>>>(p.bf.ctl, p.bf.st.cost, p.sc.B);
      }
    }
    enqueue_iterations_parts(s, P, max_iterations);
    if (s->sc.check_convergence)
      for (const Part& p : P) {  // the check of the last accepted step (cc:2604, 2673)
        launch_traj(dm, p.sc, p.bf, false, false, p.st);
        launch_tau(dm, p.sc, p.bf, false, false, p.st);
        launch_partials(dm, p.sc, p.bf, false, p.st);
        launch_assemble(dm, p.sc, p.bf, false, p.st);
        launch_lagrange(dm, p.sc, p.bf, false, p.st);
        launch_conv_check(p.sc, p.bf, p.st);
        launch_clear_dirty(p.sc, p.bf, p.st);
      }
    return IDTO_OK;
  }
  use_main(s);
  k_set_ctl<<<(s->sc.B + 127) / 128, 128, 0, s->stream>>>(s->bf.ctl, s->sc.B, 2, 0.0);
  if (s->sc.check_convergence) {
    // previous_cost = EvalCost(state) (cc:2494)
    enqueue_trajectory(s, false, false);
    k_set_prev_cost<<<(s->sc.B + 127) / 128, 128, 0, s->stream>>>(s->bf.ctl, s->bf.st.cost, s->sc.B);
  }
  for (int k = 0; k < max_iterations; ++k) enqueue_iteration(s);
  if (s->sc.check_convergence) {
    // the check of the last accepted step needs the new state's merit gradient (cc:2604, 2673)
    enqueue_trajectory(s, false, false);
    enqueue_derivatives(s, false);
    enqueue_assembly(s, false);
    launch_conv_check(s->sc, s->bf, s->stream);
    launch_clear_dirty(s->sc, s->bf, s->stream);
  }
  return IDTO_OK;
}

static int solve_collect(idto_solver_t s, int max_iterations, int* iters_out, int* reason_out, double* stats_out) {
  const int B = s->sc.B;
  IDTO_CUDA_CHECK(cudaMemcpy(s->ctl_host.data(), s->bf.ctl, B * sizeof(ProbCtl), cudaMemcpyDeviceToHost));
  for (int b = 0; b < B; ++b) {
    if (iters_out) iters_out[b] = s->ctl_host[b].iters;
    if (reason_out) reason_out[b] = s->ctl_host[b].reason;
  }
  if (stats_out) {
    for (int b = 0; b < B; ++b)
      IDTO_CUDA_CHECK(cudaMemcpy(stats_out + size_t(b) * max_iterations * IDTO_NUM_STATS,
                                 s->bf.stats + size_t(b) * s->bf.stats_cap * IDTO_NUM_STATS,
                                 size_t(max_iterations) * IDTO_NUM_STATS * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return IDTO_OK;
}

int idto_solve(idto_solver_t s, int max_iterations, int* iters_out, int* reason_out, double* stats_out) {
  if (!s || max_iterations < 0) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  if (int rc = solve_enqueue(s, max_iterations)) return rc;
  if (int rc = check_status(s)) return rc;
  return solve_collect(s, max_iterations, iters_out, reason_out, stats_out);
}

// host pointers of an MPC advance that leads a re-solve (idto_mpc_resolve_async); null `elapsed`: none
struct MpcArgs {
  const double *elapsed = nullptr, *q0 = nullptr, *v0 = nullptr, *selector = nullptr;
};
static int mpc_advance_enqueue(idto_solver_t s, const MpcArgs& m);
static int resolve_async_enqueue(idto_solver_t s, int max_iterations, const double* q_guess, const double* q_init,
                                 const double* v_init, const double* q_nom, const double* v_nom, double* q_out,
                                 double* v_out, double* tau_out, int* iters_out, double* stats_out);
static int advance_and_resolve_enqueue(idto_solver_t s, const MpcArgs& m, int max_iterations, const double* q_guess,
                                       const double* q_init, const double* v_init, const double* q_nom,
                                       const double* v_nom, double* q_out, double* v_out, double* tau_out,
                                       int* iters_out, double* stats_out) {
  if (m.elapsed)
    if (int rc = mpc_advance_enqueue(s, m)) return rc;
  return resolve_async_enqueue(s, max_iterations, q_guess, q_init, v_init, q_nom, v_nom, q_out, v_out, tau_out,
                               iters_out, stats_out);
}

static void drop_graph(idto_solver_t s) {
  if (s->gexec) cudaGraphExecDestroy(s->gexec);
  s->gexec = nullptr, s->gcalls = 0, s->gkey = idto_solver_s::GraphKey();
}

static int resolve_impl(idto_solver_t s, const MpcArgs& mpc, int max_iterations, const double* q_guess,
                        const double* q_init, const double* v_init, const double* q_nom, const double* v_nom,
                        double* q_out, double* v_out, double* tau_out, int* iters_out, double* stats_out) {
  if (!s || max_iterations < 0) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  if (!s->use_graph || s->profile || s->bf.act_base)
    return advance_and_resolve_enqueue(s, mpc, max_iterations, q_guess, q_init, v_init, q_nom, v_nom, q_out, v_out, tau_out,
                                 iters_out, stats_out);
  if (int rc = ensure_stats(s, size_t(max_iterations))) return rc;  // (allocates: not inside a capture)
  idto_solver_s::GraphKey key;
  key.iters = max_iterations, key.stats_cap = s->stats_cap;
  const void* ptrs[14] = {q_guess, q_init, v_init, q_nom, v_nom, q_out, v_out, tau_out, iters_out, stats_out,
                          mpc.elapsed, mpc.q0, mpc.v0, mpc.selector};
  std::memcpy(key.p, ptrs, sizeof(ptrs));
  if (!(key == s->gkey)) {
    drop_graph(s);
    s->gkey = key;
  }
  if (s->gexec) {  // replay
    use_main(s);
    IDTO_CUDA_CHECK(cudaGraphLaunch(s->gexec, s->stream));
    g_launch_counter += s->glaunches;
    return IDTO_OK;
  }
  if (s->gcalls++ == 0)  // first call with these arguments: eager
    return advance_and_resolve_enqueue(s, mpc, max_iterations, q_guess, q_init, v_init, q_nom, v_nom, q_out, v_out, tau_out,
                                 iters_out, stats_out);
  // second call: capture the same enqueue sequence (the sub-batch streams fork from and join into the caller's
  // stream through events, which puts them into the capture), instantiate, launch
  use_main(s);
  cudaStream_t cs = s->stream;
  cudaStream_t own = nullptr;
  if (cs == nullptr) {  // the legacy default stream cannot be captured: capture on a private one, replay on `stream`
    IDTO_CUDA_CHECK(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
    s->stream = own;
  }
  const long l0 = g_launch_counter;
  cudaGraph_t graph = nullptr;
  int rc = IDTO_OK;
  if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    rc = IDTO_ERR_CUDA;
  } else {
    s->main_dirty = true;
    rc = advance_and_resolve_enqueue(s, mpc, max_iterations, q_guess, q_init, v_init, q_nom, v_nom, q_out, v_out, tau_out,
                               iters_out, stats_out);
    use_main(s);  // joins the sub-batch streams
    if (cudaStreamEndCapture(s->stream, &graph) != cudaSuccess || !graph) rc = rc ? rc : IDTO_ERR_CUDA;
  }
  s->stream = cs;
  if (own) cudaStreamDestroy(own);
  s->glaunches = g_launch_counter - l0;
  if (rc == IDTO_OK && cudaGraphInstantiate(&s->gexec, graph, 0) != cudaSuccess) rc = IDTO_ERR_CUDA;
  if (graph) cudaGraphDestroy(graph);
  if (rc != IDTO_OK) {  // capture is not available here: stay eager from now on
    cudaGetLastError();
    drop_graph(s);
    s->use_graph = false;
    g_launch_counter = l0;
    return advance_and_resolve_enqueue(s, mpc, max_iterations, q_guess, q_init, v_init, q_nom, v_nom, q_out, v_out, tau_out,
                                 iters_out, stats_out);
  }
  IDTO_CUDA_CHECK(cudaGraphLaunch(s->gexec, s->stream));
  s->main_dirty = true;
  return IDTO_OK;
}

int idto_resolve_async(idto_solver_t s, int max_iterations, const double* q_guess, const double* q_init,
                       const double* v_init, const double* q_nom, const double* v_nom, double* q_out, double* v_out,
                       double* tau_out, int* iters_out, double* stats_out) {
  return resolve_impl(s, MpcArgs(), max_iterations, q_guess, q_init, v_init, q_nom, v_nom, q_out, v_out, tau_out,
                      iters_out, stats_out);
}

// One MPC re-plan in one call: idto_mpc_advance followed by idto_resolve_async without new inputs
// (ModelPredictiveController::UpdateAbstractState, examples/mpc_controller.cc:43-85, is exactly this sequence).
// With pinned inputs the advance kernel is part of the captured graph of the re-solve; pageable inputs are copied
// eagerly first (a copy from pageable memory does not belong into a graph that is replayed).
int idto_mpc_resolve_async(idto_solver_t s, const double* elapsed, const double* q0, const double* v0,
                           const double* q_nom_selector, int max_iterations, double* q_out, double* v_out,
                           double* tau_out, int* iters_out, double* stats_out) {
  if (!s || !elapsed || !q0 || !v0 || max_iterations < 0) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);
  const bool pinned = mapped_alias(elapsed) && mapped_alias(q0) && mapped_alias(v0) &&
                      (!q_nom_selector || mapped_alias(q_nom_selector));
  if (!pinned) {
    if (int rc = idto_mpc_advance(s, elapsed, q0, v0, q_nom_selector)) return rc;
    return resolve_impl(s, MpcArgs(), max_iterations, nullptr, nullptr, nullptr, nullptr, nullptr, q_out, v_out,
                        tau_out, iters_out, stats_out);
  }
  MpcArgs m;
  m.elapsed = elapsed, m.q0 = q0, m.v0 = v0, m.selector = q_nom_selector;
  return resolve_impl(s, m, max_iterations, nullptr, nullptr, nullptr, nullptr, nullptr, q_out, v_out, tau_out,
                      iters_out, stats_out);
}

static int resolve_async_enqueue(idto_solver_t s, int max_iterations, const double* q_guess, const double* q_init,
                                 const double* v_init, const double* q_nom, const double* v_nom, double* q_out,
                                 double* v_out, double* tau_out, int* iters_out, double* stats_out) {
  const SolverConsts& c = s->sc;
  const size_t T1 = c.T + 1, B = c.B;
  if (multi(s)) {
    // every sub-batch moves its own slice of the host buffers on its own stream: copies of one
    // sub-batch overlap with the kernels of the others, and nothing touches the caller's stream
    if (int rc = ensure_stats(s, size_t(max_iterations))) return rc;
    use_subs(s);
    const bool dirty = q_guess || q_init || v_init || q_nom || v_nom;
    for (const Part& p : make_parts(s)) {
      const size_t o = size_t(p.b0), nb = size_t(p.sc.B);
      if (q_guess) IDTO_CUDA_CHECK(cudaMemcpyAsync(p.bf.st.q, q_guess + o * T1 * c.nq, nb * T1 * c.nq * 8, cudaMemcpyHostToDevice, p.st));
      if (q_init) IDTO_CUDA_CHECK(cudaMemcpyAsync(const_cast<double*>(p.bf.q_init), q_init + o * c.nq, nb * c.nq * 8, cudaMemcpyHostToDevice, p.st));
      if (v_init) IDTO_CUDA_CHECK(cudaMemcpyAsync(const_cast<double*>(p.bf.v_init), v_init + o * c.nv, nb * c.nv * 8, cudaMemcpyHostToDevice, p.st));
      if (q_nom) IDTO_CUDA_CHECK(cudaMemcpyAsync(const_cast<double*>(p.bf.q_nom), q_nom + o * T1 * c.nq, nb * T1 * c.nq * 8, cudaMemcpyHostToDevice, p.st));
      if (v_nom) IDTO_CUDA_CHECK(cudaMemcpyAsync(const_cast<double*>(p.bf.v_nom), v_nom + o * T1 * c.nv, nb * T1 * c.nv * 8, cudaMemcpyHostToDevice, p.st));
      if (dirty) k_set_ctl<<<(p.sc.B + 127) / 128, 128, 0, p.st>>>(p.bf.ctl, p.sc.B, 1, 0.0);
    }
    if (int rc = solve_enqueue(s, max_iterations)) return rc;
    // pinned host buffers are written by one kernel per sub-batch, pageable ones by the copy engines
    double *qm = mapped_alias(q_out), *vm = mapped_alias(v_out), *tm = mapped_alias(tau_out),
           *sm = max_iterations > 0 ? mapped_alias(stats_out) : nullptr;
    int* im = mapped_alias(iters_out);
    for (const Part& p : make_parts(s)) {
      const size_t o = size_t(p.b0), nb = size_t(p.sc.B);
      // solution = {q, EvalV, EvalTau} (cc:2636-2638): current after any iteration (its first stage evaluates a
      // stale trajectory, an accepted step adopts the scratch one), so only a 0-iteration call evaluates here
      if (max_iterations == 0) {
        launch_traj(s->model->dm, p.sc, p.bf, false, false, p.st);
        launch_tau(s->model->dm, p.sc, p.bf, false, false, p.st);
      }
      const size_t srow = size_t(max_iterations) * IDTO_NUM_STATS;
      if (qm || vm || tm || sm || im) {
        const size_t tot = nb * (T1 * c.nq + T1 * c.nv + size_t(c.T) * c.nv);
        g_launch_counter += 1;
        k_export<<<int(std::min<size_t>(296, (tot + 511) / 512)), 256, 0, p.st>>>(
            p.bf.st.q, p.bf.st.v, p.bf.st.tau, p.bf.stats, p.bf.ctl, qm ? qm + o * T1 * c.nq : nullptr,
            vm ? vm + o * T1 * c.nv : nullptr, tm ? tm + o * c.T * c.nv : nullptr, sm ? sm + o * srow : nullptr,
            im ? im + o : nullptr, nb * T1 * c.nq, nb * T1 * c.nv, nb * c.T * c.nv, p.sc.B, max_iterations,
            s->bf.stats_cap);
      }
      if (q_out && !qm) IDTO_CUDA_CHECK(cudaMemcpyAsync(q_out + o * T1 * c.nq, p.bf.st.q, nb * T1 * c.nq * 8, cudaMemcpyDeviceToHost, p.st));
      if (v_out && !vm) IDTO_CUDA_CHECK(cudaMemcpyAsync(v_out + o * T1 * c.nv, p.bf.st.v, nb * T1 * c.nv * 8, cudaMemcpyDeviceToHost, p.st));
      if (tau_out && !tm) IDTO_CUDA_CHECK(cudaMemcpyAsync(tau_out + o * c.T * c.nv, p.bf.st.tau, nb * c.T * c.nv * 8, cudaMemcpyDeviceToHost, p.st));
      if (stats_out && !sm)
        IDTO_CUDA_CHECK(cudaMemcpy2DAsync(stats_out + o * max_iterations * IDTO_NUM_STATS,
                                          size_t(max_iterations) * IDTO_NUM_STATS * 8, p.bf.stats,
                                          s->bf.stats_cap * IDTO_NUM_STATS * 8,
                                          size_t(max_iterations) * IDTO_NUM_STATS * 8, nb, cudaMemcpyDeviceToHost, p.st));
      if (iters_out && !im) {  // rows of stats_out beyond iters_out[b] are not written by this call (early convergence)
        k_pack_iters<<<(p.sc.B + 127) / 128, 128, 0, p.st>>>(p.bf.ctl, s->iters_dev + o, p.sc.B);
        IDTO_CUDA_CHECK(cudaMemcpyAsync(iters_out + o, s->iters_dev + o, nb * sizeof(int), cudaMemcpyDeviceToHost, p.st));
      }
    }
    return IDTO_OK;
  }
  use_main(s);
  if (q_guess) IDTO_CUDA_CHECK(cudaMemcpyAsync(s->bf.st.q, q_guess, B * T1 * c.nq * 8, cudaMemcpyHostToDevice, s->stream));
  if (q_init) IDTO_CUDA_CHECK(cudaMemcpyAsync(s->q_init, q_init, B * c.nq * 8, cudaMemcpyHostToDevice, s->stream));
  if (v_init) IDTO_CUDA_CHECK(cudaMemcpyAsync(s->v_init, v_init, B * c.nv * 8, cudaMemcpyHostToDevice, s->stream));
  if (q_nom) IDTO_CUDA_CHECK(cudaMemcpyAsync(s->q_nom, q_nom, B * T1 * c.nq * 8, cudaMemcpyHostToDevice, s->stream));
  if (v_nom) IDTO_CUDA_CHECK(cudaMemcpyAsync(s->v_nom, v_nom, B * T1 * c.nv * 8, cudaMemcpyHostToDevice, s->stream));
  if (q_guess || q_init || v_init || q_nom || v_nom)
    k_set_ctl<<<(c.B + 127) / 128, 128, 0, s->stream>>>(s->bf.ctl, c.B, 1, 0.0);
  if (int rc = solve_enqueue(s, max_iterations)) return rc;
  // solution = {q, EvalV, EvalTau} (cc:2636-2638); see the sub-stream branch
  if (max_iterations == 0) enqueue_trajectory(s, false, false);
  double *qm = mapped_alias(q_out), *vm = mapped_alias(v_out), *tm = mapped_alias(tau_out),
         *sm = max_iterations > 0 ? mapped_alias(stats_out) : nullptr;
  int* im = mapped_alias(iters_out);
  if (qm || vm || tm || sm || im) {
    const size_t tot = B * (T1 * c.nq + T1 * c.nv + size_t(c.T) * c.nv);
    g_launch_counter += 1;
    k_export<<<int(std::min<size_t>(296, (tot + 511) / 512)), 256, 0, s->stream>>>(
        s->bf.st.q, s->bf.st.v, s->bf.st.tau, s->bf.stats, s->bf.ctl, qm, vm, tm, sm, im, B * T1 * c.nq,
        B * T1 * c.nv, B * c.T * c.nv, c.B, max_iterations, s->bf.stats_cap);
  }
  if (q_out && !qm) IDTO_CUDA_CHECK(cudaMemcpyAsync(q_out, s->bf.st.q, B * T1 * c.nq * 8, cudaMemcpyDeviceToHost, s->stream));
  if (v_out && !vm) IDTO_CUDA_CHECK(cudaMemcpyAsync(v_out, s->bf.st.v, B * T1 * c.nv * 8, cudaMemcpyDeviceToHost, s->stream));
  if (tau_out && !tm) IDTO_CUDA_CHECK(cudaMemcpyAsync(tau_out, s->bf.st.tau, B * c.T * c.nv * 8, cudaMemcpyDeviceToHost, s->stream));
  if (stats_out && !sm) {
    if (size_t(max_iterations) == s->stats_cap) {  // contiguous: one copy
      IDTO_CUDA_CHECK(cudaMemcpyAsync(stats_out, s->bf.stats, B * size_t(max_iterations) * IDTO_NUM_STATS * 8,
                                      cudaMemcpyDeviceToHost, s->stream));
    } else {
      IDTO_CUDA_CHECK(cudaMemcpy2DAsync(stats_out, size_t(max_iterations) * IDTO_NUM_STATS * 8, s->bf.stats,
                                        s->bf.stats_cap * IDTO_NUM_STATS * 8,
                                        size_t(max_iterations) * IDTO_NUM_STATS * 8, B, cudaMemcpyDeviceToHost,
                                        s->stream));
    }
  }
  if (iters_out && !im) {
    k_pack_iters<<<(c.B + 127) / 128, 128, 0, s->stream>>>(s->bf.ctl, s->iters_dev, c.B);
    IDTO_CUDA_CHECK(cudaMemcpyAsync(iters_out, s->iters_dev, B * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  }
  return IDTO_OK;
}

static int mpc_advance_enqueue(idto_solver_t s, const MpcArgs& m) {
  const double *elapsed = m.elapsed, *q0 = m.q0, *v0 = m.v0, *q_nom_selector = m.selector;
  const SolverConsts& c = s->sc;
  const size_t B = c.B;
  use_main(s);
  double* d_el = s->mpc_in;
  double* d_q0 = d_el + B;
  double* d_v0 = d_q0 + B * c.nq;
  double* d_sel = d_v0 + B * c.nv;
  // pinned inputs (19 KB for 64 quadrupeds) are read by the kernel itself through their mapped aliases: four
  // copy-engine transfers with their fixed costs are not worth it for that
  const double *m_el = mapped_alias(elapsed), *m_q0 = mapped_alias(q0), *m_v0 = mapped_alias(v0),
               *m_sel = mapped_alias(q_nom_selector);
  if (m_el && m_q0 && m_v0 && (m_sel || !q_nom_selector)) {
    d_el = const_cast<double*>(m_el), d_q0 = const_cast<double*>(m_q0), d_v0 = const_cast<double*>(m_v0);
    d_sel = const_cast<double*>(m_sel);
  } else {
    IDTO_CUDA_CHECK(cudaMemcpyAsync(d_el, elapsed, B * 8, cudaMemcpyHostToDevice, s->stream));
    IDTO_CUDA_CHECK(cudaMemcpyAsync(d_q0, q0, B * c.nq * 8, cudaMemcpyHostToDevice, s->stream));
    IDTO_CUDA_CHECK(cudaMemcpyAsync(d_v0, v0, B * c.nv * 8, cudaMemcpyHostToDevice, s->stream));
    if (q_nom_selector)
      IDTO_CUDA_CHECK(cudaMemcpyAsync(d_sel, q_nom_selector, size_t(c.nq) * 8, cudaMemcpyHostToDevice, s->stream));
  }
  if (int rc = launch_mpc_advance(c, s->bf, d_el, d_q0, d_v0, q_nom_selector ? d_sel : nullptr, s->q_init, s->v_init,
                                  s->q_nom, s->stream)) {
    set_last_error("idto_mpc_advance: horizon too long for the spline workspace");
    return rc;
  }
  s->main_dirty = true;
  return IDTO_OK;
}

int idto_mpc_advance(idto_solver_t s, const double* elapsed, const double* q0, const double* v0,
                     const double* q_nom_selector) {
  if (!s || !elapsed || !q0 || !v0) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  MpcArgs m;
  m.elapsed = elapsed, m.q0 = q0, m.v0 = v0, m.selector = q_nom_selector;
  return mpc_advance_enqueue(s, m);
}

int idto_fence(idto_solver_t s) {
  if (!s) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  use_main(s);  // joins the sub-streams into the caller's stream; later sub-stream work forks after it
  return IDTO_OK;
}

int idto_flush_l2(idto_solver_t s, void* scratch, size_t bytes) {
  if (!s || !scratch) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  // One memset of the whole scratch on the caller's stream: use_main() orders it after everything the sub-batch
  // streams hold, and the next sub-stream work forks after it — every step starts with a cold L2 on all streams
  // (per-sub-stream partial memsets, the first version, left up to half of the L2 warm).
  use_main(s);
  IDTO_CUDA_CHECK(cudaMemsetAsync(scratch, 0, bytes, s->stream));
  return IDTO_OK;
}

int idto_synchronize(idto_solver_t s) {
  if (!s) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  use_main(s);
  return check_status(s);
}

int idto_debug_pair_trace(idto_solver_t s, int enable) {
  if (!s) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);
  use_main(s);
  IDTO_CUDA_CHECK(cudaStreamSynchronize(s->stream));
  if (!enable) {  // (the buffers stay allocated until the solver is destroyed)
    s->bf.act_base = s->bf.act_fd = nullptr;
    return IDTO_OK;
  }
  if (!use_chain_kernels(s->model->dm)) {
    set_last_error("the contact-pair trace is implemented by the chain-lane inverse-dynamics kernels only");
    return IDTO_ERR_UNSUPPORTED;
  }
  const size_t np = size_t(std::max(s->model->dm.np, 1)), BT = size_t(s->sc.B) * s->sc.T;
  if (s->mem.get(&s->bf.act_base, BT * np) != cudaSuccess || s->mem.get(&s->bf.act_fd, BT * s->sc.nq * 4 * np) != cudaSuccess) {
    s->bf.act_base = s->bf.act_fd = nullptr;
    set_last_error("cudaMalloc failed for the contact-pair trace");
    return IDTO_ERR_CUDA;
  }
  k_set_ctl<<<(s->sc.B + 127) / 128, 128, 0, s->stream>>>(s->bf.ctl, s->sc.B, 1, 0.0);  // re-evaluate everything
  return IDTO_OK;
}

long idto_launch_count(idto_solver_t s) { return s ? g_launch_counter - s->launches0 : IDTO_ERR_INVALID_ARG; }

int idto_profile_enable(idto_solver_t s, int enable) {
  if (!s) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  use_main(s);
  cudaStreamSynchronize(s->stream);
  for (auto& kv : s->prof)
    for (auto& e : kv.second) cudaEventDestroy(e.a), cudaEventDestroy(e.b);
  s->prof.clear();
  s->profile = enable != 0;
  return IDTO_OK;
}

int idto_profile_read(idto_solver_t s, const char* kernel, double* total_ms, long* launches) {
  if (!s || !kernel) return IDTO_ERR_INVALID_ARG;
  cudaSetDevice(s->device);  // the caller may have changed the current device since creation
  IDTO_CUDA_CHECK(cudaStreamSynchronize(s->stream));
  double tot = 0.0;
  long n = 0;
  auto it = s->prof.find(kernel);
  if (it != s->prof.end())
    for (auto& e : it->second) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) tot += ms, ++n;
    }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = n;
  return IDTO_OK;
}

}  // extern "C"
