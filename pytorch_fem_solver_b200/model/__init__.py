"""Training loop and MLP (API kept from torch_fem/model; not part of the assembly hot path)."""

from .model import Model
from .neural_network import FeedForwardNeuralNetwork, IdentityBC

__all__ = ["Model", "FeedForwardNeuralNetwork", "IdentityBC"]
