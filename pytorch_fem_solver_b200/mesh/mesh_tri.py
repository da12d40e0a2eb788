"""Single triangular mesh (reference torch_fem/mesh/mesh_tri.py)."""

import torch

from .abstract_mesh import AbstractMesh


class MeshTri(AbstractMesh):
    """Triangular mesh built from a `triangle`-style dictionary."""

    @property
    def _edges_permutations(self):
        return torch.tensor([[0, 1], [1, 2], [0, 2]])
