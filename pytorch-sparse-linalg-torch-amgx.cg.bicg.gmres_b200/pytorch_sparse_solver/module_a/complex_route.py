"""Complex systems (reference torch_sparse_linalg.py:100-127 `_vdot_real_part`, :1220 conjugate transpose)."""


def complex_solve(name, A, b, x0=None, **kw):
    raise NotImplementedError("complex systems: route under construction")
