"""odefilters_b200 -- host-side mirror of ProbNumDiffEq.jl's interface for the ODE-filter hot path.

The reference is Julia; Julia is not available in the build image, so this Python layer stands
where the Julia wrapper (``julia/ProbNumDiffEqB200.jl``, shipped as source) would stand: it
marshals problems to structure-of-arrays, calls the C-ABI library ``libpnde.so`` once per ensemble
and rebuilds the reference's solution fields (``src/solution.jl:8-24``).  All numerics run in the
CUDA kernels behind the C ABI; there is no CPU path here.

Names follow the reference: ``EK0``/``EK1`` (src/algorithms.jl:23-51), ``solve``, ``ODEProblem``,
``EnsembleProblem``, ``SRMatrix`` (src/squarerootmatrix.jl), ``Gaussian``.
"""
from .api import (  # noqa: F401
    CustomVectorField,
    EK0,
    EK1,
    EnsembleB200,
    EnsembleProblem,
    EnsembleSolution,
    FilterSolver,
    Gaussian,
    IEKS,
    ODEProblem,
    ProbODESolution,
    pinned_empty,
    SRMatrix,
    shard_range,
    shard_strided,
    solve,
    solve_ieks,
    mean,
    std,
    var,
)
from . import _lib  # noqa: F401

__all__ = ["CustomVectorField", "EK0", "EK1", "EnsembleB200", "EnsembleProblem", "EnsembleSolution", "FilterSolver", "Gaussian",
           "IEKS", "ODEProblem", "ProbODESolution", "SRMatrix", "pinned_empty", "shard_range", "shard_strided", "solve", "solve_ieks", "mean", "std", "var"]
