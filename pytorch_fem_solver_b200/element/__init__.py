"""Reference elements (quadrature tables and P1 shape functions)."""

from .abstract_element import AbstractElement
from .element_line import ElementLine
from .element_tri import ElementTri

__all__ = ["AbstractElement", "ElementLine", "ElementTri"]
