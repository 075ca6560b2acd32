"""`import pyidto` — the module name the reference's examples import (python_bindings/pyidto.cc:12-23,
python_examples/*.py), served by the CUDA path: a re-export of idto_b200.pyidto.  pydrake is optional here (the
reference imports pydrake.multibody.plant at load, pyidto.cc:15): `plant` arguments are BakedPlant objects (baked
tables + time step; idto_b200.bake reads URDF / SDF, include/idto_b200_drake.hpp bakes from a real Drake plant)."""
from idto_b200.pyidto import (BakedPlant, FindIdtoResource, ProblemDefinition, SolverParameters,  # noqa: F401
                              TrajectoryOptimizer, TrajectoryOptimizerSolution, TrajectoryOptimizerStats, WarmStart)

__all__ = ["BakedPlant", "FindIdtoResource", "ProblemDefinition", "SolverParameters", "TrajectoryOptimizer",
           "TrajectoryOptimizerSolution", "TrajectoryOptimizerStats", "WarmStart"]
