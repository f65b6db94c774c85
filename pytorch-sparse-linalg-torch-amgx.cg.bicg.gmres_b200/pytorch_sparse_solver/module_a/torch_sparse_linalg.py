"""Import-path compatibility: the reference keeps its solvers in module_a/torch_sparse_linalg.py.
Everything lives in .krylov here; this module only re-exports the public names."""
from .krylov import (  # noqa: F401
    cg, bicgstab, gmres,
    cg_differentiable, bicgstab_differentiable, gmres_differentiable,
    LinearSolveFunction,
)
