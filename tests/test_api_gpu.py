"""Parity of the CUDA path (through the C ABI and the custom ops) with the reference's golden
outputs and with the oracle.  Needs a GPU: `pytest -m gpu`."""

import pytest
import torch

from tests import api_checks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize(
    "name,orders,aligned",
    [
        ("structured4x4", (1, 2, 3, 4), True),
        ("structured6x5_jitter", (3,), True),
        ("delaunay60", (2, 3, 4), True),
        ("structured3x3_neighbors", (2,), False),
    ],
)
def test_single_mesh(golden, name, orders, aligned):
    api_checks.check_single_mesh(golden(name), orders, "cuda", aligned)


def test_patches(golden):
    api_checks.check_patches(golden("patches_l2"), "cuda")


@pytest.mark.parametrize("name", ["fractures2_4x2", "fractures2_8x4"])
def test_fractures(golden, name):
    api_checks.check_fractures(golden(name), "cuda")


def test_native_library_is_the_one_running():
    """The ops must have gone through libtfem_b200.so (no silent fallback)."""
    from pytorch_fem_solver_b200 import _lib

    assert _lib._lib is not None
    assert sum(_lib.LAUNCHES.values()) > 0


def test_cpu_tensors_are_refused():
    from pytorch_fem_solver_b200 import _lib, ops

    coords = torch.zeros((3, 2), dtype=torch.float64)
    conn = torch.zeros((1, 3), dtype=torch.int32)
    with pytest.raises(_lib.TfemError):
        ops.tri_geometry(coords, conn, 1, 3, 2)


@pytest.mark.parametrize("nx,ny", [(8, 4), (64, 32)])
def test_seven_fracture_network(nx, ny):
    """BASELINE config 5, scaled down (CSR output above 8192 DOFs): kernels vs the oracle + properties."""
    assert api_checks.check_seven_fractures("cuda", nx=nx, ny=ny) == 7 * (nx + 1) * (ny + 1) - 6 * (ny + 1)
