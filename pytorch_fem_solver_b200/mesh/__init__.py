"""Mesh containers with the reference's key layout."""

from .abstract_mesh import AbstractMesh
from .fractures_tri import FracturesTri
from .mesh_tri import MeshTri
from .meshes_tri import MeshesTri
from .patches import Patches

__all__ = ["AbstractMesh", "FracturesTri", "MeshTri", "MeshesTri", "Patches"]
