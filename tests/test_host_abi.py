"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/idto_b200.h declares, the value types mirror the reference's defaults
(python_bindings/test/solver_parameters_test.py, problem_definition_test.py), the bake step restates
the model conventions the oracle is pinned on, and the product fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from idto_b200 import problems
from idto_b200.bake import BakedModel, load_model, make_from_one_vector, rpy_to_R
from idto_b200.types import (GRAD_FORWARD, SCALING_DOUBLE_SQRT, ProblemDefinition, SolverParameters,
                             TrajectoryOptimizerSolution, TrajectoryOptimizerStats)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from idto_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "idto_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(idto_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 25
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_params_default_matches_reference():
    """optimizer/solver_parameters.h:64-167 defaults through the C ABI and the Python mirror."""
    from idto_b200 import capi
    from idto_b200.types import Params
    p = Params()
    capi.lib().idto_params_default(ctypes.byref(p))
    assert (p.max_iterations, p.gradients_method, p.normalize_quaternions) == (100, GRAD_FORWARD, 0)
    assert (p.contact_stiffness, p.dissipation_velocity, p.stiction_velocity) == (100.0, 0.1, 0.05)
    assert (p.friction_coefficient, p.smoothing_factor) == (0.5, 0.1)
    assert (p.scaling, p.scaling_method, p.equality_constraints) == (1, SCALING_DOUBLE_SQRT, 1)
    assert (p.Delta0, p.Delta_max, p.check_convergence) == (1e-1, 1e5, 0)
    s = SolverParameters()
    c = s.to_c()
    for f, _ in Params._fields_:
        assert getattr(c, f) == getattr(p, f), f
    assert s.num_threads == 1 and s.verbose is True


def test_no_gpu_fails_loudly():
    """There is no CPU fallback: without a device, model creation must raise, not degrade."""
    from idto_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.IdtoError, match="no CUDA device"):
        capi.Model(load_model("pendulum"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "idto_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                bad = re.findall(r"^\s*(?:from\s+oracle|import\s+oracle|#include\s+.*oracle).*$|liboracle|oracle_[a-z_]+\(",
                                 src, flags=re.M)
                assert not bad, (os.path.join(dirpath, f), bad)


def test_stats_and_solution_types():
    """trajectory_optimizer_solution.h:58-185: 13 series, push_data order, is_empty, CSV header."""
    st = TrajectoryOptimizerStats()
    assert st.is_empty()
    st.push_row(0.01, [1.0, 0.1, 2.0, 0.3, 0.4, 0.9, 5.0, -0.1, 0.2, 1.5])
    assert not st.is_empty()
    assert st.iteration_costs == [1.0] and st.trust_region_radii == [0.1] and st.q_norms == [2.0]
    assert st.dq_norms == [0.3] and st.dqH_norms == [0.4] and st.trust_ratios == [0.9]
    assert st.gradient_norms == [5.0] and st.dL_dqs == [-0.1] and st.h_norms == [0.2] and st.merits == [1.5]
    assert st.linesearch_iterations == [0] and np.isnan(st.linesearch_alphas[0])
    sol = TrajectoryOptimizerSolution()
    assert sol.q == [] and sol.v == [] and sol.tau == []


def test_problem_definition_size_checks():
    """cc:74-82: wrong q_nom / v_nom sizes are rejected (ValueError instead of DRAKE_DEMAND abort)."""
    m, dt, prob, params, guess = problems.spinner()
    prob.to_c(dt, m.nq, m.nv)
    bad = ProblemDefinition(**{**prob.__dict__, "q_nom": prob.q_nom[:-1]})
    with pytest.raises(ValueError):
        bad.to_c(dt, m.nq, m.nv)


# ----------------------------------------------------------------------------- bake step
def test_bake_conventions():
    ch = load_model("mini_cheetah")
    assert (ch.nbodies, ch.nq, ch.nv) == (13, 19, 18)
    assert ch.body_names[0] == "body" and ch.joint_type[0] == 3 and ch.quat_q_starts == [0]
    # depth-first dof order: base, then each leg abduct-thigh-shank (mini_cheetah.yaml:8-13)
    assert ch.body_names[1:4] == ["abduct_fl", "thigh_fl", "shank_fl"]
    assert ch.parent.tolist() == [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11]
    assert ch.unactuated_dofs == [0, 1, 2, 3, 4, 5]
    # welded feet merged into the shanks; ground box anchored to the world and registered first
    assert ch.geom_body.tolist() == [-1, 3, 6, 9, 12] and ch.geom_type.tolist() == [1, 0, 0, 0, 0]
    assert list(zip(ch.pair_geomA.tolist(), ch.pair_geomB.tolist())) == [(0, 1), (0, 2), (0, 3), (0, 4)]
    assert np.isclose(ch.mass.sum(), 3.3 + 4 * (0.54 + 0.634 + 0.064))
    sp = load_model("spinner")
    assert sp.unactuated_dofs == [2] and sp.npairs == 1 and np.allclose(sp.damping, 0.1)
    hp = load_model("hopper")
    assert hp.unactuated_dofs == [0, 1, 2] and hp.npairs == 2  # tt:1598; sphere(A)-box(B) x2
    assert hp.geom_type.tolist() == [0, 0, 1] and hp.geom_body.tolist() == [2, 2, -1]
    ac = load_model("acrobot")
    assert ac.unactuated_dofs == [0]  # shoulder: no transmission / effort limit 0


def test_planar_joint_frame_matches_hopper_example():
    """examples/hopper/hopper.yaml:7-8: q = [height, horizontal, theta, knee, ankle]."""
    hp = load_model("hopper")
    R = hp.X_PF[0][:9].reshape(3, 3)
    assert np.allclose(R[:, 0], [0, 0, 1])   # planar x = world z (height)
    assert np.allclose(R[:, 1], [-1, 0, 0])  # planar y = -world x
    assert np.allclose(R[:, 2], [0, -1, 0])  # normal = URDF axis
    assert np.allclose(make_from_one_vector([0, 0, 1], 2), np.eye(3))
    assert np.allclose(rpy_to_R(0, -np.pi / 2, 0) @ [1, 0, 0], [0, 0, 1])


def test_baked_model_json_roundtrip(tmp_path):
    m = load_model("hopper")
    p = tmp_path / "m.json"
    m.save(p)
    m2 = BakedModel.load(p)
    for k in BakedModel._INT + BakedModel._DBL:
        assert np.array_equal(np.asarray(getattr(m, k)), np.asarray(getattr(m2, k))), k


def test_top_level_pyidto_module_reexports_the_binding():
    """python_bindings/pyidto.cc:12-23: the examples do `from pyidto import TrajectoryOptimizer, ...`."""
    import pyidto
    from idto_b200 import pyidto as impl
    for name in ("TrajectoryOptimizer", "ProblemDefinition", "SolverParameters", "TrajectoryOptimizerSolution",
                 "TrajectoryOptimizerStats", "FindIdtoResource"):
        assert getattr(pyidto, name) is getattr(impl, name)
    p = pyidto.SolverParameters()
    assert p.max_iterations == 100 and p.Delta0 == 1e-1  # python_bindings/test/solver_parameters_test.py


def test_python_constants_match_the_header_enums():
    """idto_b200/types.py restates the enums of include/idto_b200.h: every IDTO_<GROUP>_<NAME> = value there must
    equal <GROUP>_<NAME> here (gradient methods, scaling methods, linear solvers, status codes)."""
    import re
    from idto_b200 import types as T
    hdr = open(os.path.join(ROOT, "include", "idto_b200.h")).read()
    found = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"IDTO_([A-Z0-9_]+)\s*=\s*(-?\d+)", hdr))
    checked = 0
    for name, val in found.items():
        for prefix in ("GRAD_", "SCALING_", "LINSOLVE_"):
            if name.startswith(prefix):
                assert getattr(T, name) == val, name
                checked += 1
    assert checked >= 11 and found["LINSOLVE_CYCLIC_REDUCTION"] == 3
