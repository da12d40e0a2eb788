"""Module kept for import parity with the reference
(torch_fem/basis/interior_edges_fracture_basis.py); the class lives beside its planar sibling."""

from .interior_edges_basis import InteriorEdgesFractureBasis

__all__ = ["InteriorEdgesFractureBasis"]
