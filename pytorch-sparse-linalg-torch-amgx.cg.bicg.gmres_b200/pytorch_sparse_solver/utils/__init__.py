from .availability import (  # noqa: F401
    check_module_a_available, check_module_b_available, check_module_c_available,
    get_available_backends, get_available_backend_list, print_availability_report,
)
from .matrix_utils import (  # noqa: F401
    dense_to_sparse_csr, sparse_coo_to_csr, ensure_sparse_format, get_csr_components,
    create_sparse_csr_from_components, create_tridiagonal_sparse_coo, create_poisson_2d_sparse_coo,
    compute_residual, compute_relative_residual,
)
