"""pytest configuration: import paths, the `gpu` marker, golden-fixture helpers."""
import json
import sys
import warnings
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
PKG_DIR = ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"
for p in (str(PKG_DIR), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

warnings.filterwarnings("ignore", message=".*Sparse CSR tensor support is in beta.*")
warnings.filterwarnings("ignore", message=".*Sparse invariant checks.*")

GOLD = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def manifest():
    with open(GOLD / "manifest.json") as f:
        return json.load(f)


def load_case(name):
    with np.load(GOLD / f"{name}.npz") as z:
        return {k: torch.from_numpy(z[k].copy()) for k in z.files}


def build_matrix(gen, device="cpu", index_dtype=torch.int64):
    from pytorch_sparse_solver import problems
    kind = gen["matrix"]
    if kind == "poisson2d":
        return problems.poisson2d_csr(gen["nx"], gen["ny"], device=device, index_dtype=index_dtype)
    if kind == "poisson3d":
        return problems.poisson3d_csr(gen["n"], device=device, index_dtype=index_dtype)
    if kind == "convdiff3d":
        return problems.convdiff3d_csr(gen["n"], device=device, index_dtype=index_dtype)
    if kind == "scaled_convdiff3d":
        return problems.scaled_convdiff3d_csr(gen["n"], seed=gen.get("seed", 7), device=device, index_dtype=index_dtype)
    if kind == "scaled_poisson3d":
        return problems.scaled_poisson3d_csr(gen["n"], seed=gen.get("seed", 7), device=device, index_dtype=index_dtype)
    if kind == "ldc":
        return problems.ldc_pressure_csr(gen["nx"], device=device, index_dtype=index_dtype)
    if kind == "tridiag":
        n = gen["n"]
        A = (2.0 * torch.eye(n, dtype=torch.float64) - torch.diag(torch.ones(n - 1, dtype=torch.float64), 1)
             - torch.diag(torch.ones(n - 1, dtype=torch.float64), -1))
        return A.to_sparse_csr().to(device)
    raise KeyError(kind)


def rel_diff(x, ref):
    x = x.detach().cpu().double()
    ref = ref.detach().cpu().double()
    d = torch.linalg.norm(x - ref)
    n = torch.linalg.norm(ref)
    return float(d / n) if float(n) > 0 else float(d)
