"""Host-side glue of the API classes vs. the reference's golden outputs, on CPU.

The CUDA ops are swapped for oracle-backed stand-ins (tests/cpu_shim.py) so that shapes, index
maps, batch flattening and result layouts are verified without a GPU; `test_api_gpu.py` runs the
same checks against the real kernels."""

import pytest

from tests import api_checks, cpu_shim


@pytest.mark.parametrize(
    "name,orders,aligned",
    [
        ("structured4x4", (1, 2, 3, 4), True),
        ("structured6x5_jitter", (3,), True),
        ("delaunay60", (2, 3, 4), True),
        ("structured3x3_neighbors", (2,), False),
    ],
)
def test_single_mesh(golden, monkeypatch, name, orders, aligned):
    cpu_shim.install(monkeypatch)
    api_checks.check_single_mesh(golden(name), orders, "cpu", aligned)


def test_patches(golden, monkeypatch):
    cpu_shim.install(monkeypatch)
    api_checks.check_patches(golden("patches_l2"), "cpu")


@pytest.mark.parametrize("name", ["fractures2_4x2", "fractures2_8x4"])
def test_fractures(golden, monkeypatch, name):
    cpu_shim.install(monkeypatch)
    api_checks.check_fractures(golden(name), "cpu")


def test_seven_fracture_network(monkeypatch):
    """BASELINE config 5, scaled down: 7 planes, 6 trace lines (the reference cannot build this one)."""
    cpu_shim.install(monkeypatch)
    assert api_checks.check_seven_fractures("cpu", nx=8, ny=4) == 7 * 45 - 6 * 5
