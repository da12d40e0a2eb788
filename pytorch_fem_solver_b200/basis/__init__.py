"""Finite-element bases: geometry, index maps, integrate_* and interpolate on the CUDA path."""

from .abstract_basis import AbstractBasis
from .basis import Basis
from .fracture_basis import FractureBasis
from .interior_edges_basis import InteriorEdgesBasis
from .interior_edges_fracture_basis import InteriorEdgesFractureBasis
from .patches_basis import PatchesBasis

__all__ = [
    "AbstractBasis",
    "Basis",
    "FractureBasis",
    "InteriorEdgesBasis",
    "InteriorEdgesFractureBasis",
    "PatchesBasis",
]
