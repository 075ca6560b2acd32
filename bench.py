#!/usr/bin/env python
"""bench.py — Gauss-Newton iterations/sec on Mini Cheetah T=40 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            the CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path on the host cores

A "step" is one pass of the hot path over one batch: one trust-region iteration (derivative
pipeline + penta-diagonal solves + dogleg + trust ratio, optimizer/trajectory_optimizer.cc:2495-2625)
of each of the `batch` = 64 independent warm-started MPC problems resident on a GPU, i.e. one
`SolveFromWarmStart(max_iterations=1)` per problem, exactly what the reference's MPC loop issues per
re-plan (`mpc_iters: 1`).  Every step starts from a state whose caches are stale (as after the MPC
guess shift), so no step is a cheap "rejected step" replay.
  value = problems * steps / device time, inputs resident in HBM.
  e2e   = the same through the public C-ABI calls an MPC loop makes (ModelPredictiveController::
          UpdateAbstractState, examples/mpc_controller.cc:43-98) with pinned HOST buffers: per step H2D of
          the measured state (q0, v0, elapsed time) into idto_mpc_resolve_async (= idto_mpc_advance, which shifts
          the previous solution on the device, + the re-solve), and D2H of the solution trajectory and stats.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from idto_b200 import problems  # noqa: E402
from idto_b200.types import GRAD_CENTRAL, GRAD_FORWARD, NUM_STATS  # noqa: E402

BATCH = 64
METRIC = "Gauss-Newton iters/sec, Mini Cheetah T=40"
# name -> (problems.<fn>, BASELINE horizon, description); mini_cheetah is the configuration the metric is quoted on,
# the others are BASELINE.json's other configs (driver-timed lines for them: --workload)
WORKLOADS = {
    "mini_cheetah": ("mini_cheetah", 40, "mini_cheetah trot (floating base + 12 DOF, 4 foot-ground pairs)"),
    "allegro_hand": ("allegro_hand", 60, "allegro_hand in-hand sphere rotation (16 finger joints + free ball, 188 "
                                         "candidate contact pairs)"),
    "hopper": ("hopper", 50, "hopper planar contact (planar base + 2 joints, 2 foot-ground pairs)"),
    "spinner": ("spinner", 40, "spinner (2-link finger + spinner, 1 sphere-sphere pair)"),
}


def workload(method, name="mini_cheetah"):
    fn, T, _ = WORKLOADS[name]
    m, dt, prob, params, guess = getattr(problems, fn)(T=T, gradients_method=method, max_iterations=1)
    return m, dt, prob, params


def b_idp_bytes(nq, nv, T):
    """Algorithmic bytes of one ID-partials evaluation of one problem."""
    return 8 * ((T + 1) * (nq + nv) + 2 * T * nv + (T + 1) * nv * nq + 3 * T * nv * nq)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            probe = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True)
            if probe.returncode != 0 or "not a valid" in (probe.stdout + probe.stderr).lower():
                self.Q = self.Q.replace("clocks_event_reasons", "clocks_throttle_reasons")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_arm(method, steps, warmup, nprob=BATCH, cores=None, name="mini_cheetah"):
    """The reference's CPU path (restated: oracle/idto_oracle.cc — Drake is not installable, so this is
    'kind: port').  Independent optimizers run on independent host threads (the reference's threading
    contract), one WarmStart each, OpenMP off inside (every core already has a problem)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle
    m, dt, prob, params = workload(method, name)
    cores = cores or os.cpu_count() or 1
    nthreads = min(cores, nprob)
    omp = max(1, cores // nprob)  # spare cores go to the reference's own OpenMP loops over t (cc:214, 455)
    q0, v0, qg = problems.perturbed_batch(m, prob, nprob)
    orcs = []
    for b in range(nprob):
        o = oracle.Oracle(m, dt, prob, params, num_threads=omp)
        o.reset_initial_conditions(q0[b], v0[b])
        o.set_q(qg[b])
        orcs.append(o)

    def one(o):
        o.set_q(o.get("q"))  # stale caches, like an MPC re-solve
        return o.solve(1)[0]

    with ThreadPoolExecutor(nthreads) as ex:
        for _ in range(warmup):
            list(ex.map(one, orcs))
        t0 = time.perf_counter()
        iters = 0
        for _ in range(steps):
            iters += sum(ex.map(one, orcs))
        el = time.perf_counter() - t0
    return (iters / el, el / steps * 1e3, nthreads * omp,
            f"{nprob} problems x {steps} iteration(s), {nthreads} host threads x {omp} OpenMP threads each")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="idto_b200", choices=["idto_b200", "reference"])
    ap.add_argument("--method", default="central", choices=["central", "forward"])
    ap.add_argument("--batch", type=int, default=BATCH, help="problems per GPU (weak) or in the whole job (strong)")
    ap.add_argument("--workload", default="mini_cheetah", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch problems per GPU; strong: --batch problems split over the GPUs "
                         "(BASELINE config 4: 'batch=64 MPC solves across 1/2/4/8 GPUs')")
    ap.add_argument("--linear-solver", default="sweep", choices=["sweep", "cyclic_reduction"],
                    help="KKT / Newton-step solver: the two-sided block sweep (default) or block cyclic reduction "
                         "(kernels_cr.cu; the parallel-in-time order BASELINE.json's north star names)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--l2", default="rotate", choices=["rotate", "flush"],
                    help="cold-L2 protocol between timed steps: 'rotate' = inputs larger than L2 (several resident "
                         "batches take turns, one batch per step), 'flush' = a 160 MiB memset before every step")
    args = ap.parse_args()
    method = GRAD_CENTRAL if args.method == "central" else GRAD_FORWARD
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)
    m, dt, prob, params = workload(method, args.workload)
    if args.linear_solver == "cyclic_reduction":
        from idto_b200.types import LINSOLVE_CYCLIC_REDUCTION
        params.linear_solver = LINSOLVE_CYCLIC_REDUCTION
    T = WORKLOADS[args.workload][1]
    metric = METRIC if args.workload == "mini_cheetah" else f"Gauss-Newton iters/sec, {args.workload} T={T}"
    if args.scaling == "strong":
        if args.batch % world:
            raise SystemExit(f"bench.py: --scaling strong needs --batch ({args.batch}) divisible by the GPU count ({world})")
        B = args.batch // world
    else:
        B = args.batch
    config = {"workload": f"{WORKLOADS[args.workload][2]}, T={T}, "
                          + (f"batch={args.batch} independent MPC re-solves per GPU" if args.scaling == "weak" else
                             f"batch={args.batch} independent MPC re-solves in the whole job, {B} per GPU")
                          + f", 1 iteration per step, gradients={args.method}_differences, equality_constraints=on, "
                            "scaling=double_sqrt"
                          + (", linear_solver=cyclic_reduction" if args.linear_solver == "cyclic_reduction" else ""),
              "model": m.name, "nq": m.nq, "nv": m.nv, "T": T, "batch_per_gpu": B, "global_batch": B * world,
              "l2": None}

    if args.impl == "reference":
        if rank != 0:
            return
        # every step is one iteration of each of the --batch problems (the whole workload of one GPU step: it is
        # small enough on the host, ~0.1-0.3 s, that the sample is not cut down); exactly --steps / --warmup are run
        ref_warmup = max(args.warmup, 1)
        val, ms, cores, sample = cpu_arm(method, max(1, args.steps), ref_warmup, nprob=args.batch, name=args.workload)
        line = {"impl": "reference", "metric": metric, "value": val, "unit": "iters/s", "n_gpus": args.gpus,
                "steps": max(1, args.steps), "warmup": ref_warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    from idto_b200 import capi
    if not torch.cuda.is_available() or capi.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device — the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    model = capi.Model(m)
    free0 = torch.cuda.mem_get_info()[0]
    gs = capi.BatchSolver(model, dt, prob, params, B)
    state_bytes = max(free0 - torch.cuda.mem_get_info()[0], 1)  # HBM held by one resident batch
    # batch element b of rank r is problem r*B + b of the global job
    from idto_b200.sharding import shard_slice
    q0, v0, qg = problems.perturbed_batch(m, prob, B * world)
    sl = shard_slice(B * world, rank, world)  # contiguous slice; no data-path collective
    q0, v0, qg = q0[sl], v0[sl], qg[sl]
    gs.reset_initial_conditions(q0, v0)
    gs.set_q(qg)
    # Cold L2 between timed steps.  'rotate': inputs larger than L2 — NSETS resident copies of the batch (same
    # problems, separate HBM) take turns, ONE batch per step, steps serialised on the stream; by the time a copy is
    # used again the others have pushed more than two L2 capacities through the cache.  'flush': one copy and a
    # 160 MiB memset (L2 = 126 MB) in stream order before every step, inside the timed region.
    L2_BYTES = 126e6
    nsets = 1
    if args.l2 == "rotate":
        nsets = max(3, int(np.ceil(2.5 * L2_BYTES / state_bytes)) + 1)
        if nsets > 48:  # tiny problems: too many copies; fall back to the memset
            nsets, args.l2 = 1, "flush"
    sets = [gs]
    for _ in range(nsets - 1):
        g2 = capi.BatchSolver(model, dt, prob, params, B)
        g2.reset_initial_conditions(q0, v0)
        g2.set_q(qg)
        sets.append(g2)
    config["l2"] = (f"inputs larger than L2: {nsets} resident copies of the batch ({state_bytes / 1e6:.0f} MB of solver "
                    f"state each, L2 = 126 MB) take turns, one batch per step, steps serialised"
                    if args.l2 == "rotate" else
                    "flushed before every step: a 160 MiB memset (L2 = 126 MB) on the solver's stream, ordered after "
                    "the previous step and before the next on every internal stream, inside the timed region")
    step_no = [0]

    def next_set():
        g = sets[step_no[0] % nsets]
        step_no[0] += 1
        return g

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ----------------------------------------------------------
    # K re-solves back to back, timed with ONE CUDA-event pair on the solver's stream around the whole
    # region (idto_fence orders the internal sub-batch streams against the events).  L2 hygiene: before
    # every step a 160 MiB scratch buffer is overwritten in stream order INSIDE the timed region (the
    # per-step working set, ~110 MB of bands / KKT sweep / partials, would otherwise fit the 126 MB L2).
    flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")  # L2 is 126 MB

    def step_resident(g=None):
        g = g or next_set()
        if args.l2 == "flush":
            g.flush_l2(flush.data_ptr(), flush.numel())
        g.invalidate()
        g.resolve_async(1)  # no host pointers: everything stays in HBM, nothing synchronises
        g.fence()           # joins the internal streams: the next step (another copy) starts after this one

    # clocks / throttle reasons are sampled every 20 ms from the warm-up to the end of the end-to-end region (the
    # timed regions sit inside that window, back to back, all of it under load)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(warmup, 2 * nsets)):  # (every copy captures its CUDA graph on its second call)
        step_resident()
    barrier()
    l0 = gs.launch_count()  # (the counter is process-wide: launches of every copy)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    launches = gs.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers ------------------------------------
    T1 = T + 1
    hq0, hv0 = torch.from_numpy(q0.copy()).pin_memory(), torch.from_numpy(v0.copy()).pin_memory()
    hel = torch.zeros(B, dtype=torch.float64).pin_memory()  # elapsed time since the stored solution: 0 keeps the
    # workload of the resident arm (guess = previous solution at its knots; q_guess[0] = the measured q0)
    oq = torch.empty((B, T1, m.nq), dtype=torch.float64).pin_memory()
    ov = torch.empty((B, T1, m.nv), dtype=torch.float64).pin_memory()
    ot = torch.empty((B, T, m.nv), dtype=torch.float64).pin_memory()
    ost = torch.empty((B, 1, NUM_STATS), dtype=torch.float64).pin_memory()
    hq0_np, hv0_np, hel_np = hq0.numpy(), hv0.numpy(), hel.numpy()
    h2d = (hq0.numel() + hv0.numel() + hel.numel()) * 8
    d2h = (oq.numel() + ov.numel() + ot.numel() + ost.numel()) * 8

    def step_e2e():
        g = next_set()
        if args.l2 == "flush":
            g.flush_l2(flush.data_ptr(), flush.numel())
        # one MPC re-plan = one call (ModelPredictiveController::UpdateAbstractState, examples/mpc_controller.cc:43-85):
        # measured state in (H2D), guess shifted on the device, one iteration, solution + stats out (D2H)
        g.mpc_resolve_async(hel_np, hq0_np, hv0_np, 1, q_out=oq.data_ptr(), v_out=ov.data_ptr(),
                            tau_out=ot.data_ptr(), stats_out=ost.data_ptr())
        g.synchronize()

    for _ in range(max(warmup, 2 * nsets)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()  # H2D + iteration + D2H + host synchronisation each step, wall clock
    barrier()
    el = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(el.item())
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel (ID partials), CUDA events on the launching stream --------
    gs.profile_enable(True)
    for _ in range(2):  # untimed: first single-stream launches of this copy (side stream, function attributes)
        step_resident(gs)
    gs.profile_enable(True)  # (drops the timers of the two steps above)
    for _ in range(5):
        gs.flush_l2(flush.data_ptr(), flush.numel())  # (stage times: one copy, cold L2 by memset)
        step_resident(gs)
    gs.synchronize()
    stage_ms = {}
    for name in ("id_partials", "trajectory", "assemble", "factor", "lagrange", "dogleg", "trajectory_scratch",
                 "trust_update"):
        tot, n = gs.profile_read(name)
        stage_ms[name] = tot / max(n, 1)
    gs.profile_enable(False)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    algo_bytes = b_idp_bytes(m.nq, m.nv, T) * B
    achieved = algo_bytes / (stage_ms["id_partials"] * 1e-3) / 1e9
    traffic, fp64 = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            prof = json.load(f)
        key = f"id_partials_{args.method}" if args.workload == "mini_cheetah" and B == BATCH else "none"
        traffic = prof.get(key)
        flop = prof.get(f"{key}_fp64_flop")
        if flop:  # what actually binds this stage: fp64 issue, not HBM
            fp64 = {"flop_per_launch": flop, "achieved_tflops": flop / (stage_ms["id_partials"] * 1e-3) / 1e12,
                    "peak_tflops": 148 * 64 * 2 * 1.965e9 / 1e12, "source": "ncu instruction counts, profiles/ncu_traffic.json"}
    except OSError:
        pass
    # The stage that dominates the step is the KKT sweep (lagrange): its roof is neither HBM nor flops but the serial
    # pivot chain; reported against the fp64 peak with the flops the reference's block recurrence needs
    # (penta_diagonal_solver.h:124-248 on blocks of kb = nq + nu: per block row one LU (2/3 kb^3), two kb^3
    # triangular solve pairs for Y and Z and three kb x kb products for K, G: ~ (2/3 + 2*2 + 3*2) kb^3).
    kb = m.nq + (len(m.unactuated_dofs) if params.equality_constraints else 0)
    kkt_flop = B * (T + 1) * (2.0 / 3 + 4 + 6) * kb ** 3
    fp64_peak = 148 * 64 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    kkt_ms = stage_ms["lagrange"]
    roofline_kkt = {"kernel": "lagrange stage: KKT sweep (k_kkt_*) + k_gm_matvec", "bound": "fp64 (latency of the serial "
                    "pivot chain in practice)", "achieved": kkt_flop / (kkt_ms * 1e-3) / 1e12 if kkt_ms > 0 else None,
                    "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": kkt_flop / (kkt_ms * 1e-3) / 1e12 / fp64_peak if kkt_ms > 0 else None,
                    "algorithmic_flop_per_launch": kkt_flop, "avg_launch_ms": kkt_ms, "block_size": kb,
                    "peak_source": "148 SMs x 64 fp64 FMA/clk x 2 x sm_max_mhz"}
    roofline = {"kernel": "id_partials stage: k_partials_path + k_partials_chain", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": algo_bytes,
                "avg_launch_ms": stage_ms["id_partials"],
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "note": "arithmetic intensity ~50 fp64 flop/B: this kernel is fp64-issue/latency bound, not HBM "
                        "bound (SURVEY.md §8d); the HBM fraction is reported as BASELINE.json defines it",
                "stage_ms": stage_ms, "fp64": fp64}

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "iters/s", "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches), "roofline": roofline, "roofline_kkt": roofline_kkt}
        if world == 1 and not args.no_cpu_baseline:
            val, cms, cores, sample = cpu_arm(method, 2, 1, nprob=B, name=args.workload)
            line["cpu_baseline"] = {"value": val, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample,
                                    "ms_per_step": cms}
        out_line = json.dumps(line)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        # printed last (after the process group is gone) and pushed through the file descriptor right away
        sys.stdout.write(out_line + "\n")
        sys.stdout.flush()
        try:
            os.fsync(sys.stdout.fileno())
        except OSError:
            pass


if __name__ == "__main__":
    main()
