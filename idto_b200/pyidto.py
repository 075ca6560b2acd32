"""`pyidto`-shaped Python API over the CUDA path.

Mirrors the names and signatures bound in the reference's python_bindings/*.cc:
  TrajectoryOptimizer(diagram, plant, problem, params): time_step, num_steps, Solve(q_guess, solution,
  stats) -> None, SolveFromWarmStart(warm_start, solution, stats) -> None, CreateWarmStart(q_guess),
  ResetInitialConditions, UpdateNominalTrajectory, params, prob (trajectory_optimizer_py.cc:34-59);
  WarmStart.{set_q, get_q, Delta, dq, dqH} (:60-67); ProblemDefinition; SolverParameters;
  TrajectoryOptimizerSolution; TrajectoryOptimizerStats; FindIdtoResource (find_resource.cc:11-14).
`diagram` / `plant` are Drake objects in the reference; here `plant` is a `BakedPlant` (baked tables +
time step) and `diagram` is ignored (kept for signature compatibility).
"""
from __future__ import annotations

import copy
import os
import time

import numpy as np

from . import capi
from .bake import BakedModel, load_model
from .types import (ProblemDefinition, SolverParameters, TrajectoryOptimizerSolution,  # noqa: F401
                    TrajectoryOptimizerStats)


class BakedPlant:
    """Stand-in for `MultibodyPlant` after `Finalize()`: baked tables + discrete time step."""

    def __init__(self, model: BakedModel | str, time_step: float):
        self.model = load_model(model) if isinstance(model, str) else model
        self._time_step = float(time_step)
        self._device = None

    def time_step(self):
        return self._time_step

    def num_positions(self):
        return self.model.nq

    def num_velocities(self):
        return self.model.nv

    def device_model(self):
        if self._device is None:
            self._device = capi.Model(self.model)
        return self._device


def FindIdtoResource(path: str) -> str:
    """utils/find_resource.cc:11-14: paths must start with 'idto/'."""
    if not path.startswith("idto/"):
        raise RuntimeError(f"FindIdtoResource: '{path}' does not start with 'idto/'")
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), path[len("idto/"):])


class WarmStart:
    """optimizer/warm_start.h:23-76 (no Python constructor in the reference either: use CreateWarmStart)."""

    def __init__(self, solver: capi.BatchSolver, q_guess, version):
        self._s, self._version = solver, version
        self.set_q(q_guess)

    def set_q(self, q_guess):
        self._s.set_q(np.asarray(q_guess, float)[None])

    def get_q(self):
        return list(self._s.get("q")[0].reshape(self._s.T + 1, self._s.nq))

    @property
    def Delta(self):
        return float(self._s.get("delta")[0, 0])

    @property
    def dq(self):
        return self._s.get("dq")[0]

    @property
    def dqH(self):
        return self._s.get("dqH")[0]


class TrajectoryOptimizer:
    def __init__(self, diagram, plant: BakedPlant, problem: ProblemDefinition, params: SolverParameters = None):
        self._plant = plant
        self._prob = copy.deepcopy(problem)   # copied by value, like the reference (cc:47)
        self._params = copy.deepcopy(params) if params is not None else SolverParameters()
        T = self._prob.num_steps
        if len(self._prob.q_nom) != T + 1 or len(self._prob.v_nom) != T + 1:  # cc:75-76
            raise ValueError("q_nom and v_nom must have num_steps + 1 entries")
        self._version = 0
        self._model = plant.device_model()

    def time_step(self):
        return self._plant.time_step()

    def num_steps(self):
        return self._prob.num_steps

    def params(self):
        return copy.deepcopy(self._params)

    def prob(self):
        return copy.deepcopy(self._prob)

    def unactuated_dofs(self):
        return self._model.unactuated_dofs()

    def num_equality_constraints(self):
        return len(self.unactuated_dofs()) * self.num_steps()

    def CreateWarmStart(self, q_guess) -> WarmStart:
        q_guess = np.asarray(q_guess, float)
        if q_guess.shape != (self.num_steps() + 1, self._plant.num_positions()):  # cc:1356-1357
            raise ValueError("q_guess must be (num_steps + 1, nq)")
        s = capi.BatchSolver(self._model, self.time_step(), self._prob, self._params, 1)
        return WarmStart(s, q_guess, self._version)

    def ResetInitialConditions(self, q_init, v_init):
        q_init, v_init = np.asarray(q_init, float), np.asarray(v_init, float)
        if q_init.shape != (self._plant.num_positions(),) or v_init.shape != (self._plant.num_velocities(),):
            raise ValueError("wrong initial condition sizes")
        self._prob.q_init, self._prob.v_init = q_init.copy(), v_init.copy()
        self._version += 1

    def UpdateNominalTrajectory(self, q_nom, v_nom):
        if len(q_nom) != self.num_steps() + 1 or np.asarray(q_nom[0]).size != self._plant.num_positions():
            raise ValueError("wrong nominal trajectory sizes")
        self._prob.q_nom = [np.array(x, float) for x in q_nom]
        self._prob.v_nom = [np.array(x, float) for x in v_nom]
        self._version += 1

    def _sync_problem(self, ws: WarmStart):
        if ws._version != self._version:
            ws._s.reset_initial_conditions(self._prob.q_init[None], self._prob.v_init[None])
            ws._s.update_nominal_trajectory(np.asarray(self._prob.q_nom)[None], np.asarray(self._prob.v_nom)[None])
            ws._version = self._version

    def SolveFromWarmStart(self, warm_start: WarmStart, solution: TrajectoryOptimizerSolution,
                           stats: TrajectoryOptimizerStats):
        """cc:2449-2651.  Stats are appended (not required empty)."""
        self._sync_problem(warm_start)
        t0 = time.perf_counter()
        iters, reason, st = warm_start._s.solve(self._params.max_iterations)
        el = time.perf_counter() - t0
        n = int(iters[0])
        for k in range(n):
            stats.push_row(el / max(n, 1), st[0, k])
        stats.solve_time = el
        q, v, tau = warm_start._s.solution()
        solution.q, solution.v, solution.tau = list(q[0]), list(v[0]), list(tau[0])

    def Solve(self, q_guess, solution: TrajectoryOptimizerSolution, stats: TrajectoryOptimizerStats):
        """cc:2213-2234: the guess must start at q_init and stats must be empty (abort in the reference)."""
        q_guess = np.asarray(q_guess, float)
        if not np.array_equal(q_guess[0], self._prob.q_init):
            raise RuntimeError("Solve: q_guess[0] must equal q_init")
        if not stats.is_empty():
            raise RuntimeError("Solve: stats must be empty")
        self.SolveFromWarmStart(self.CreateWarmStart(q_guess), solution, stats)
