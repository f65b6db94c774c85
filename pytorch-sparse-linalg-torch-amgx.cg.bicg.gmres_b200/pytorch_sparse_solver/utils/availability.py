"""Backend availability probes (reference utils/availability.py:13-178).  This build ships Module A only:
module_b (AMGX) and module_c (cuDSS) are outside the hot-path scope and always report unavailable."""
from functools import lru_cache
from typing import Dict, List


@lru_cache(maxsize=1)
def check_module_a_available() -> bool:
    """Module A here IS the CUDA library: available only when libbk_krylov.so is built and a CUDA device exists
    (there is no CPU fallback to select)."""
    try:
        import torch
        from .. import module_a  # noqa: F401
        from .. import _native
        return _native.library_path().exists() and torch.cuda.is_available()
    except Exception:
        return False


def check_module_b_available() -> bool:
    return False


def check_module_c_available() -> bool:
    return False


def get_available_backends() -> Dict[str, bool]:
    return {'module_a': check_module_a_available(), 'module_b': False, 'module_c': False}


def get_available_backend_list() -> List[str]:
    return [k for k, v in get_available_backends().items() if v]


def print_availability_report() -> None:
    for name, ok in get_available_backends().items():
        print(f"  {name}: {'available' if ok else 'not available in this build'}")
