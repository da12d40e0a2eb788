"""Element interface mirrored from the reference (torch_fem/element/abstract_element.py:8-62).

In this package the element object is a *descriptor*: the CUDA kernels carry the same
quadrature tables and closed-form maps (csrc/common.cuh), selected by
`polynomial_order` / `integration_order`.  The tensor-valued helpers below exist so scripts
written against the reference API (`element.gaussian_nodes`, `compute_inverse_map`, ...) keep
working; the assembly path never calls them.
"""

from __future__ import annotations

import abc
from typing import Tuple

import torch


class AbstractElement(abc.ABC):
    """Reference element: polynomial order + quadrature rule."""

    def __init__(self, polynomial_order: int, integration_order: int):
        self.polynomial_order = polynomial_order
        self.integration_order = integration_order
        self.gaussian_nodes, self.gaussian_weights = self._compute_gauss_values()

    def compute_inverse_map(self, first_node, integration_points, inv_map_jacobian):
        """Physical -> reference coordinates, `(x - p0) J^-T` (reference :18-26)."""
        return (integration_points - first_node) @ inv_map_jacobian.mT

    @property
    def n_quadrature_points(self) -> int:
        return self.gaussian_nodes.shape[0]

    @abc.abstractmethod
    def compute_shape_functions(self, bar_coords, inv_map_jacobian) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    @abc.abstractmethod
    def _compute_gauss_values(self) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    @abc.abstractmethod
    def compute_barycentric_coordinates(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    @abc.abstractmethod
    def compute_det_and_inv_map(self, map_jacobian: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    @property
    @abc.abstractmethod
    def reference_element_area(self) -> float:
        raise NotImplementedError

    @property
    @abc.abstractmethod
    def barycentric_grad(self) -> torch.Tensor:
        raise NotImplementedError
