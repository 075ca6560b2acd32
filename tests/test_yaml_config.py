"""YAML front-end (examples/yaml_config.h, examples/example_base.cc:377-543): the reference's own example
configs must parse into the same ProblemDefinition / SolverParameters as the restated problems."""
import os

import numpy as np
import pytest

from idto_b200 import problems, yaml_config
from idto_b200.types import (GRAD_CENTRAL, GRAD_FORWARD, SCALING_DOUBLE_SQRT, ProblemDefinition)

REF = "/root/reference/examples"

HOPPER_LIKE = """
q_init : [ 0.61, 0.0, 0.3,-0.5, 0.2]
v_init : [ 0.0, 0.0, 0.0, 0.0, 0.0]
q_nom_start : [ 0.61, 0.0, 0.3,-0.5, 0.2]
q_nom_end :   [ 0.61,-0.5, 0.3,-0.5, 0.2]
q_nom_relative_to_q_init : [false, true, false, false, false]
q_guess :   [ 0.61,-0.5, 0.3,-0.5, 0.2]
Qq : [1.0, 1.0, 1.0, 1.0, 1.0]
Qv : [0.1, 0.1, 0.1, 0.1, 0.1]
R : [1e2, 1e2, 1e2, 0.1, 0.1]
Qfq : [10, 10, 10, 10, 10]
Qfv : [1.0, 1.0, 1.0, 1.0, 1.0]
time_step : 0.05
num_steps : 10
max_iters : 7
method : "trust_region"
Delta0 : 1e-3
tolerances:
  rel_cost_reduction: 1e-6
gradients_method: "central_differences"
contact_stiffness : 800
smoothing_factor : 0.01
friction_coefficient : 1.0
"""


def test_inline_yaml_defaults_and_mapping():
    o = yaml_config.load_yaml_string(HOPPER_LIKE)
    assert o.num_steps == 10 and o.max_iters == 7 and o.scaling_method == "double_sqrt" and o.linesearch == "armijo"
    assert o.stiction_velocity == 0.05 and o.dissipation_velocity == 0.1  # struct defaults (yaml_config.h:140-143)
    m = problems.load_model("hopper")
    prob = yaml_config.SetProblemDefinition(o, m)
    assert isinstance(prob, ProblemDefinition) and len(prob.q_nom) == 11 and len(prob.v_nom) == 11
    # q_nom relative to q_init for dof 1 only (example_base.cc:394-407); v_nom by differences (nq == nv)
    assert np.allclose(prob.q_nom[-1], [0.61, -0.5 + 0.0, 0.3, -0.5, 0.2])
    assert np.allclose(prob.v_nom[1], (prob.q_nom[1] - prob.q_nom[0]) / 0.05) and np.allclose(prob.v_nom[0], 0)
    assert np.allclose(np.diag(prob.R), [1e2, 1e2, 1e2, 0.1, 0.1]) and prob.Qq.shape == (5, 5)
    p = yaml_config.SetSolverParameters(o)
    assert p.gradients_method == GRAD_CENTRAL and p.max_iterations == 7 and p.Delta0 == 1e-3
    assert p.scaling and p.scaling_method == SCALING_DOUBLE_SQRT and p.equality_constraints
    # tolerances are copied, the check itself stays off (example_base.cc:427-543 never sets check_convergence)
    assert not p.check_convergence and p.convergence_tolerances.rel_cost_reduction == 1e-6
    assert p.q_nom_relative_to_q_init.tolist() == [False, True, False, False, False] and p.unsupported == []
    g = yaml_config.MakeInitialGuess(o)
    assert len(g) == 11 and np.allclose(g[0], o.q_init) and np.allclose(g[-1], o.q_guess)


@pytest.mark.parametrize("key,val,msg", [("gradients_method", "bogus", "Unknown gradient method 'bogus'"),
                                         ("method", "bogus", "Unknown solver method 'bogus'"),
                                         ("linesearch", "bogus", "Unknown linesearch method 'bogus'"),
                                         ("linear_solver", "bogus", "Unknown linear solver 'bogus'"),
                                         ("scaling_method", "bogus", "Unknown scaling method 'bogus'")])
def test_unknown_option_strings_raise_like_the_reference(key, val, msg):
    o = yaml_config.load_yaml_string(HOPPER_LIKE)
    setattr(o, key, val)
    with pytest.raises(RuntimeError, match=msg):
        yaml_config.SetSolverParameters(o)


def test_unknown_key_is_an_error_and_off_path_options_are_recorded():
    with pytest.raises(RuntimeError, match="unknown key"):
        yaml_config.load_yaml_string("not_a_field: 1")
    o = yaml_config.load_yaml_string(HOPPER_LIKE)
    o.method, o.linear_solver = "linesearch", "dense_ldlt"
    p = yaml_config.SetSolverParameters(o)
    assert p.unsupported == ["method: linesearch"] and p.linear_solver == 2  # dense_ldlt is provided (cross-check solver)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not mounted on this box")
@pytest.mark.parametrize("name,yaml_rel,T", [("acrobot", "acrobot/acrobot.yaml", None),
                                             ("hopper", "hopper/hopper.yaml", None),
                                             ("allegro_hand", "allegro_hand/allegro_hand.yaml", None),
                                             ("allegro_hand_upside_down",
                                              "allegro_hand/allegro_hand_upside_down.yaml", None)])
def test_reference_example_yaml_equals_restated_problem(name, yaml_rel, T):
    o = yaml_config.load_yaml(os.path.join(REF, yaml_rel))
    m, dt, prob_r, params_r, guess_r = getattr(problems, name)(T=o.num_steps)
    prob = yaml_config.SetProblemDefinition(o, m)
    params = yaml_config.SetSolverParameters(o)
    assert dt == o.time_step and prob.num_steps == prob_r.num_steps
    for f in ("q_init", "v_init", "Qq", "Qv", "Qf_q", "Qf_v", "R"):
        assert np.allclose(getattr(prob, f), getattr(prob_r, f)), f
    assert np.allclose(np.array(prob.q_nom), np.array(prob_r.q_nom))
    assert np.allclose(np.array(prob.v_nom), np.array(prob_r.v_nom))
    # (problems.spinner restates python_bindings/test/trajectory_optimizer_test.py, whose nominal trajectory
    # differs from examples/spinner/spinner.yaml, so it is not compared here)
    contact = ("contact_stiffness", "dissipation_velocity", "stiction_velocity", "friction_coefficient",
               "smoothing_factor")  # acrobot.yaml leaves them at the yaml struct's defaults; it has no contact
    for f in (() if name == "acrobot" else contact) + ("scaling", "equality_constraints", "Delta0"):
        assert getattr(params, f) == getattr(params_r, f), f
    assert params.gradients_method == GRAD_FORWARD


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not mounted on this box")
def test_every_reference_example_yaml_parses():
    n = 0
    for root, _, files in os.walk(REF):
        for f in files:
            if f.endswith(".yaml"):
                o = yaml_config.load_yaml(os.path.join(root, f))
                yaml_config.SetSolverParameters(o)
                assert o.num_steps > 0 and o.q_init.size > 0, f
                n += 1
    assert n >= 5


@pytest.mark.gpu
def test_yaml_driven_solve_runs_on_the_gpu():
    from idto_b200 import capi
    o = yaml_config.load_yaml_string(HOPPER_LIKE)
    m = problems.load_model("hopper")
    prob, params = yaml_config.SetProblemDefinition(o, m), yaml_config.SetSolverParameters(o)
    gs = capi.BatchSolver(capi.Model(m), o.time_step, prob, params, 2)
    gs.set_q(np.array(yaml_config.MakeInitialGuess(o)))
    it, reason, stats = gs.solve(o.max_iters)
    assert it[0] >= 1 and np.all(np.isfinite(stats[0, :it[0], 0])) and stats[0, it[0] - 1, 0] < stats[0, 0, 0]
    o.method = "linesearch"
    with pytest.raises(capi.IdtoError, match="unsupported"):
        capi.BatchSolver(capi.Model(m), o.time_step, prob, yaml_config.SetSolverParameters(o), 1)
