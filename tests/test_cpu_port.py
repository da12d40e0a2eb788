"""The torch CPU port used as bench baseline reproduces the oracle (and hence the reference)."""

import numpy as np
import torch

from oracle import fem_oracle as fo
from oracle.torch_cpu_port import reference_assembly_cpu
from pytorch_fem_solver_b200 import meshgen


def test_port_matches_oracle():
    mesh = meshgen.structured_rectangle(13, 9, jitter=0.25, seed=8, topology=False)
    coords, conn = mesh["vertices"], mesh["triangles"]
    n = coords.shape[0]
    timings = {}
    matrix, load = reference_assembly_cpu(torch.from_numpy(coords), torch.from_numpy(conn), 3, timings)
    geo = fo.tri_geometry(coords, conn, 3)
    crow, col, vals = fo.scatter_bilinear_csr(fo.quad_reduce(fo.form_stiffness_mass(geo), geo["dx"]), conn, n)
    assert np.array_equal(matrix.crow_indices().numpy(), crow) and np.array_equal(matrix.col_indices().numpy(), col)
    np.testing.assert_allclose(matrix.values().numpy(), vals, rtol=1e-13, atol=1e-16)
    f_q = fo.source_sinsin(geo["integration_points"])
    ref_load = fo.scatter_linear(fo.quad_reduce(fo.form_load(geo, f_q), geo["dx"]), conn, n)
    np.testing.assert_allclose(load.numpy(), ref_load, rtol=1e-13, atol=1e-16)
    assert set(timings) == {"geometry", "local_matrix", "local_load", "coo_to_csr", "total"}
