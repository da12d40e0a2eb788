"""B200-native element-assembly hot path of torch_fem behind the torch_fem Python API.

Same public names as the reference package (`/root/reference/torch_fem/__init__.py:3-28`);
geometry, quadrature, weak residuals, jump terms and the local-to-global scatter run as
hand-written sm_100a CUDA kernels behind the C ABI of `include/tfem_b200.h`.
"""

from . import csr, forms, meshgen, ops, sparse
from .basis import Basis, FractureBasis, InteriorEdgesBasis, InteriorEdgesFractureBasis, PatchesBasis
from .element import ElementLine, ElementTri
from .mesh import FracturesTri, MeshesTri, MeshTri, Patches
from .model import FeedForwardNeuralNetwork, Model

__all__ = [
    "Basis",
    "FractureBasis",
    "InteriorEdgesBasis",
    "InteriorEdgesFractureBasis",
    "PatchesBasis",
    "ElementLine",
    "ElementTri",
    "FracturesTri",
    "MeshTri",
    "MeshesTri",
    "Patches",
    "Model",
    "FeedForwardNeuralNetwork",
    "forms",
    "csr",
    "meshgen",
    "ops",
    "sparse",
]
