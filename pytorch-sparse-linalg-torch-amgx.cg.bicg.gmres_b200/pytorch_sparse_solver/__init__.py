"""pytorch_sparse_solver — drop-in for the reference package of the same name, with Module A's Krylov
inner loop (CG / BiCGStab / GMRES on CSR) running on a hand-written sm_100a CUDA library.
Same public names as the reference __init__.py:46-113."""

__version__ = '1.0.0'

from .solver import (  # noqa: F401
    SparseSolver, SolverResult, SolverMethod, SolverBackend,
    solve, cg, bicgstab, gmres, amg, direct_solve,
)
from .utils.availability import (  # noqa: F401
    check_module_a_available, check_module_b_available, check_module_c_available,
    get_available_backends, print_availability_report,
)
from .utils.matrix_utils import (  # noqa: F401
    dense_to_sparse_csr, sparse_coo_to_csr, ensure_sparse_format,
    create_tridiagonal_sparse_coo, create_poisson_2d_sparse_coo,
    compute_residual, compute_relative_residual,
)

__all__ = [
    '__version__',
    'SparseSolver', 'SolverResult', 'SolverMethod', 'SolverBackend',
    'solve', 'cg', 'bicgstab', 'gmres', 'amg', 'direct_solve',
    'check_module_a_available', 'check_module_b_available', 'check_module_c_available',
    'get_available_backends', 'print_availability_report',
    'dense_to_sparse_csr', 'sparse_coo_to_csr', 'ensure_sparse_format',
    'create_tridiagonal_sparse_coo', 'create_poisson_2d_sparse_coo',
    'compute_residual', 'compute_relative_residual',
]
