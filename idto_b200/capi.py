"""ctypes binding of the C ABI (include/idto_b200.h) — the product path.

Fails loudly when libidto_b200.so is missing or no CUDA device is present: there is no CPU
fallback and nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from .bake import BakedModel, ModelDesc, model_desc
from .types import NUM_STATS, Params, ProblemDefinition, ProblemDesc, SolverParameters

_HERE = os.path.dirname(os.path.abspath(__file__))
# IDTO_B200_LIB selects another build of the same library (e.g. a -DIDTO_KKT_TIMING debug build)
LIB_PATH = os.environ.get("IDTO_B200_LIB") or os.path.join(_HERE, "lib", "libidto_b200.so")
_D = ctypes.POINTER(ctypes.c_double)
_I = ctypes.POINTER(ctypes.c_int)
_LIB = None

ERRORS = {-1: "invalid argument", -2: "unsupported", -3: "CUDA error", -4: "factorisation failed",
          -5: "no CUDA device (no CPU fallback)", -6: "too many simultaneous contacts"}


class IdtoError(RuntimeError):
    pass


def build(extra: str = ""):
    """Compile every CUDA source for sm_100a into idto_b200/lib/libidto_b200.so (nvcc cross-compiles
    without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8", "-s"]
    if extra:
        cmd.append(f"EXTRA={extra}")
    subprocess.run(cmd, check=True)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise IdtoError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(the CUDA extension is mandatory; there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        H = ctypes.c_void_p
        L.idto_last_error.restype = ctypes.c_char_p
        L.idto_set_device.argtypes = [ctypes.c_int]
        L.idto_params_default.argtypes = [ctypes.POINTER(Params)]
        L.idto_model_create.argtypes = [ctypes.POINTER(ModelDesc), ctypes.POINTER(H)]
        L.idto_model_destroy.argtypes = [H]
        L.idto_model_num_unactuated.argtypes = [H]
        L.idto_model_unactuated_dofs.argtypes = [H, _I]
        L.idto_solver_create.argtypes = [H, ctypes.POINTER(ProblemDesc), ctypes.POINTER(Params), ctypes.c_int,
                                         ctypes.POINTER(H)]
        L.idto_solver_destroy.argtypes = [H]
        L.idto_solver_set_stream.argtypes = [H, ctypes.c_void_p]
        L.idto_solver_set_substreams.argtypes = [H, ctypes.c_int]
        L.idto_set_q.argtypes = [H, _D]
        L.idto_reset_initial_conditions.argtypes = [H, _D, _D]
        L.idto_update_nominal_trajectory.argtypes = [H, _D, _D]
        L.idto_invalidate.argtypes = [H]
        L.idto_set_delta.argtypes = [H, _D]
        L.idto_get_delta.argtypes = [H, _D]
        for f in ("idto_eval_trajectory", "idto_eval_derivatives", "idto_eval_assembly", "idto_eval_dogleg",
                  "idto_eval_trust_ratio", "idto_synchronize"):
            getattr(L, f).argtypes = [H]
        L.idto_field_size.argtypes = [H, ctypes.c_char_p]
        L.idto_field_size.restype = ctypes.c_long
        L.idto_get.argtypes = [H, ctypes.c_char_p, _D]
        L.idto_solve.argtypes = [H, ctypes.c_int, _I, _I, _D]
        L.idto_resolve_async.argtypes = [H, ctypes.c_int] + [ctypes.c_void_p] * 8 + [_I, ctypes.c_void_p]
        L.idto_mpc_advance.argtypes = [H, _D, _D, _D, _D]
        L.idto_mpc_resolve_async.argtypes = [H, _D, _D, _D, _D, ctypes.c_int] + [ctypes.c_void_p] * 3 + [_I, ctypes.c_void_p]
        L.idto_fence.argtypes = [H]
        L.idto_flush_l2.argtypes = [H, ctypes.c_void_p, ctypes.c_size_t]
        L.idto_debug_pair_trace.argtypes = [H, ctypes.c_int]
        L.idto_launch_count.argtypes = [H]
        L.idto_launch_count.restype = ctypes.c_long
        L.idto_profile_enable.argtypes = [H, ctypes.c_int]
        L.idto_profile_read.argtypes = [H, ctypes.c_char_p, _D, ctypes.POINTER(ctypes.c_long)]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        msg = lib().idto_last_error().decode()
        raise IdtoError(f"idto_b200: {ERRORS.get(rc, rc)}: {msg}")


def _p(a):
    return a.ctypes.data_as(_D)


def device_count():
    return lib().idto_device_count()


def set_device(device: int):
    """Make CUDA device `device` current for the calling thread (models bind to the current device)."""
    _check(lib().idto_set_device(int(device)))


class Model:
    """Device copy of the baked tables (idto_model_t)."""

    def __init__(self, baked: BakedModel):
        self.baked = baked
        d, self._keep = model_desc(baked)
        self.h = ctypes.c_void_p()
        _check(lib().idto_model_create(ctypes.byref(d), ctypes.byref(self.h)))
        self.nu = lib().idto_model_num_unactuated(self.h)

    def unactuated_dofs(self):
        out = np.zeros(max(self.nu, 1), np.int32)
        _check(lib().idto_model_unactuated_dofs(self.h, out.ctypes.data_as(_I)))
        return out[:self.nu].tolist()

    def __del__(self):
        if getattr(self, "h", None) and self.h.value and _LIB is not None and ctypes is not None:
            _LIB.idto_model_destroy(self.h)
            self.h = ctypes.c_void_p()


class BatchSolver:
    """`batch` independent WarmStarts of one TrajectoryOptimizer, device resident (idto_solver_t)."""

    def __init__(self, model: Model, time_step: float, prob: ProblemDefinition, params: SolverParameters,
                 batch: int = 1):
        self.model, self.prob, self.params, self.B = model, prob, params, int(batch)
        if getattr(params, "unsupported", None):  # options outside the CUDA hot path (yaml_config.SetSolverParameters)
            raise IdtoError("unsupported: " + ", ".join(params.unsupported) + " (no CPU fallback)")
        self.T, self.nq, self.nv = prob.num_steps, model.baked.nq, model.baked.nv
        pd, self._keep = prob.to_c(time_step, self.nq, self.nv)
        pc = params.to_c()
        self.h = ctypes.c_void_p()
        _check(lib().idto_solver_create(model.h, ctypes.byref(pd), ctypes.byref(pc), self.B, ctypes.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None) and self.h.value and _LIB is not None and ctypes is not None:
            _LIB.idto_solver_destroy(self.h)  # (module globals may already be gone at interpreter shutdown)
            self.h = ctypes.c_void_p()

    def _arr(self, x, shape):
        a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        if a.shape == shape[1:]:
            a = np.ascontiguousarray(np.broadcast_to(a, shape))
        if a.shape != shape:
            raise ValueError(f"expected shape {shape} (or {shape[1:]} to broadcast), got {a.shape}")
        return a

    def set_q(self, q):
        a = self._arr(q, (self.B, self.T + 1, self.nq))
        _check(lib().idto_set_q(self.h, _p(a)))
        _check(lib().idto_synchronize(self.h))

    def reset_initial_conditions(self, q0, v0):
        a, b = self._arr(q0, (self.B, self.nq)), self._arr(v0, (self.B, self.nv))
        _check(lib().idto_reset_initial_conditions(self.h, _p(a), _p(b)))
        _check(lib().idto_synchronize(self.h))

    def update_nominal_trajectory(self, qn, vn):
        a = self._arr(qn, (self.B, self.T + 1, self.nq))
        b = self._arr(vn, (self.B, self.T + 1, self.nv))
        _check(lib().idto_update_nominal_trajectory(self.h, _p(a), _p(b)))
        _check(lib().idto_synchronize(self.h))

    def set_substreams(self, n):
        _check(lib().idto_solver_set_substreams(self.h, int(n)))

    def invalidate(self):
        _check(lib().idto_invalidate(self.h))

    def resolve_async(self, max_iterations, q_guess=None, q_init=None, v_init=None, q_nom=None, v_nom=None,
                      q_out=None, v_out=None, tau_out=None, stats_out=None, iters_out=None):
        """End-to-end MPC re-solve: arguments are raw host pointers (ints) of pinned buffers or None."""
        it = None if iters_out is None else ctypes.cast(ctypes.c_void_p(iters_out), _I)
        _check(lib().idto_resolve_async(self.h, int(max_iterations), q_guess, q_init, v_init, q_nom, v_nom, q_out,
                                        v_out, tau_out, it, stats_out))

    def debug_pair_trace(self, on=True):
        """Record which contact pairs every inverse-dynamics evaluation applies a force for; then
        get("pair_active") -> [B, T*npairs], get("pair_active_fd") -> [B, T*nq*4*npairs]."""
        _check(lib().idto_debug_pair_trace(self.h, int(on)))

    def mpc_advance(self, elapsed, q0, v0, q_nom_selector=None):
        """Device-side MPC shell between two re-solves (examples/mpc_controller.cc:43-98): spline-shifted
        guess, shifted nominal trajectory, new initial conditions.  elapsed: scalar or [batch] seconds."""
        def ready(x, shape):  # fast path of the MPC loop: the caller's (pinned) float64 arrays go through untouched
            return isinstance(x, np.ndarray) and x.dtype == np.float64 and x.shape == shape and x.flags.c_contiguous

        el = elapsed if ready(elapsed, (self.B,)) else np.ascontiguousarray(
            np.broadcast_to(np.asarray(elapsed, float), (self.B,)))
        q0 = q0 if ready(q0, (self.B, self.nq)) else self._arr(q0, (self.B, self.nq))
        v0 = v0 if ready(v0, (self.B, self.nv)) else self._arr(v0, (self.B, self.nv))
        sel = None if q_nom_selector is None else np.ascontiguousarray(np.asarray(q_nom_selector, float).reshape(self.nq))
        _check(lib().idto_mpc_advance(self.h, _p(el), _p(q0), _p(v0), None if sel is None else _p(sel)))

    def mpc_resolve_async(self, elapsed, q0, v0, max_iterations, q_nom_selector=None, q_out=None, v_out=None,
                          tau_out=None, stats_out=None, iters_out=None):
        """One MPC re-plan in one call (ModelPredictiveController::UpdateAbstractState, examples/mpc_controller.cc:
        43-85): mpc_advance(elapsed, q0, v0) followed by resolve_async(max_iterations, outputs).  elapsed [batch],
        q0 [batch, nq], v0 [batch, nv]: float64 arrays (pinned ones are read by the device directly); outputs: raw
        host pointers (ints) or None."""
        def ready(x, shape):
            return isinstance(x, np.ndarray) and x.dtype == np.float64 and x.shape == shape and x.flags.c_contiguous

        el = elapsed if ready(elapsed, (self.B,)) else np.ascontiguousarray(
            np.broadcast_to(np.asarray(elapsed, float), (self.B,)))
        q0 = q0 if ready(q0, (self.B, self.nq)) else self._arr(q0, (self.B, self.nq))
        v0 = v0 if ready(v0, (self.B, self.nv)) else self._arr(v0, (self.B, self.nv))
        sel = None if q_nom_selector is None else np.ascontiguousarray(np.asarray(q_nom_selector, float).reshape(self.nq))
        it = None if iters_out is None else ctypes.cast(ctypes.c_void_p(iters_out), _I)
        _check(lib().idto_mpc_resolve_async(self.h, _p(el), _p(q0), _p(v0), None if sel is None else _p(sel),
                                            int(max_iterations), q_out, v_out, tau_out, it, stats_out))
        self._mpc_keep = (el, q0, v0, sel)  # temporaries must outlive the asynchronous call

    def fence(self):
        _check(lib().idto_fence(self.h))

    def flush_l2(self, ptr, nbytes):
        _check(lib().idto_flush_l2(self.h, ptr, nbytes))

    def synchronize(self):
        _check(lib().idto_synchronize(self.h))

    def set_delta(self, d):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(d, float), (self.B,)))
        _check(lib().idto_set_delta(self.h, _p(a)))
        _check(lib().idto_synchronize(self.h))

    def eval(self, stage=4):
        f = ("idto_eval_trajectory", "idto_eval_derivatives", "idto_eval_assembly", "idto_eval_dogleg",
             "idto_eval_trust_ratio")[stage]
        _check(getattr(lib(), f)(self.h))

    def get(self, name):
        n = lib().idto_field_size(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.zeros((self.B, n))
        _check(lib().idto_get(self.h, name.encode(), _p(out)))
        return out

    def solve(self, max_iterations):
        iters = np.zeros(self.B, np.int32)
        reason = np.zeros(self.B, np.int32)
        stats = np.zeros((self.B, max(max_iterations, 1), NUM_STATS))
        _check(lib().idto_solve(self.h, int(max_iterations), iters.ctypes.data_as(_I), reason.ctypes.data_as(_I),
                                _p(stats)))
        return iters, reason, stats

    def solution(self):
        self.eval(0)
        return (self.get("q").reshape(self.B, self.T + 1, self.nq), self.get("v").reshape(self.B, self.T + 1, self.nv),
                self.get("tau").reshape(self.B, self.T, self.nv))

    def launch_count(self):
        return lib().idto_launch_count(self.h)

    def profile_enable(self, on=True):
        _check(lib().idto_profile_enable(self.h, int(on)))

    def profile_read(self, name):
        ms = ctypes.c_double(0)
        n = ctypes.c_long(0)
        _check(lib().idto_profile_read(self.h, name.encode(), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value


class MultiDeviceBatchSolver:
    """`batch` independent WarmStarts spread over the GPUs of this process: one Model + BatchSolver per device,
    one host thread per device for every call (ctypes releases the GIL), contiguous batch slices, no collective
    (SURVEY.md 8e).  The in-process twin of `torchrun bench.py --gpus N`; arrays are indexed by GLOBAL problem."""

    def __init__(self, baked: BakedModel, time_step, prob, params, batch, devices=None):
        from concurrent.futures import ThreadPoolExecutor

        from .sharding import shard_slice
        n = device_count()
        if n == 0:
            raise IdtoError("idto_b200: no CUDA device (no CPU fallback)")
        self.devices = list(range(n)) if devices is None else list(devices)
        self.devices = self.devices[:max(1, min(len(self.devices), int(batch)))]
        self.B, self.T, self.nq, self.nv = int(batch), prob.num_steps, baked.nq, baked.nv
        self.slices = [shard_slice(self.B, r, len(self.devices)) for r in range(len(self.devices))]
        self._pool = ThreadPoolExecutor(len(self.devices))

        def make(r):
            set_device(self.devices[r])
            sl = self.slices[r]
            return BatchSolver(Model(baked), time_step, prob, params, sl.stop - sl.start)
        self.shards = list(self._pool.map(make, range(len(self.devices))))

    def _each(self, fn):
        return list(self._pool.map(lambda r: fn(self.shards[r], self.slices[r]), range(len(self.shards))))

    def set_q(self, q):
        q = np.asarray(q, float)
        self._each(lambda s, sl: s.set_q(q[sl] if q.ndim == 3 else q))

    def reset_initial_conditions(self, q0, v0):
        q0, v0 = np.asarray(q0, float), np.asarray(v0, float)
        self._each(lambda s, sl: s.reset_initial_conditions(q0[sl], v0[sl]))

    def solve(self, max_iterations):
        parts = self._each(lambda s, sl: s.solve(max_iterations))
        return tuple(np.concatenate([p[i] for p in parts]) for i in range(3))

    def solution(self):
        parts = self._each(lambda s, sl: s.solution())
        return tuple(np.concatenate([p[i] for p in parts]) for i in range(3))

    def get(self, name):
        return np.concatenate(self._each(lambda s, sl: s.get(name)))
