"""Minimal nested tensor container with the surface the torch_fem mesh API uses.

The reference keeps every mesh in a `tensordict.TensorDict`
(reference: torch_fem/mesh/abstract_mesh.py:31-74, meshes_tri.py:17-31,
basis/fracture_basis.py:115-127).  `tensordict` is not part of this image, so the
package carries the small subset it needs: nested `td[("cells", "vertices")]`
get/set, `in`, `.items()`, `.batch_size`, `.auto_batch_size_()`, iteration over
the leading batch axis, `.to(device)` and a module-level `stack`.
"""

from __future__ import annotations

from typing import Any, Iterable

import numpy as np
import torch


def _as_value(value: Any) -> Any:
    if isinstance(value, TensorDict):
        return value
    if isinstance(value, dict):
        return TensorDict(value)
    if isinstance(value, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(value))
    if isinstance(value, (list, tuple)) and value and not isinstance(value[0], torch.Tensor):
        try:
            return torch.as_tensor(np.asarray(value))
        except (TypeError, ValueError):
            return value
    return value


class TensorDict:
    """Nested dict of tensors sharing leading batch dimensions."""

    def __init__(self, source: Any = None, batch_size: Iterable[int] | None = None, **kwargs: Any):
        self._data: dict[str, Any] = {}
        items = {}
        if source is not None:
            items.update(source.items() if hasattr(source, "items") else dict(source))
        items.update(kwargs)
        for key, value in items.items():
            self._data[key] = _as_value(value)
        self._batch_size = torch.Size(batch_size) if batch_size is not None else torch.Size([])

    # ---- mapping surface -------------------------------------------------
    def __getitem__(self, key):
        if isinstance(key, tuple) and all(isinstance(k, str) for k in key):
            node = self
            for k in key:
                node = node._data[k] if isinstance(node, TensorDict) else node[k]
            return node
        if isinstance(key, str):
            return self._data[key]
        # integer / slice indexing along the batch axis
        return TensorDict(
            {k: v[key] for k, v in self._data.items()},
            batch_size=self._batch_size[1:] if isinstance(key, int) else None,
        )

    def __setitem__(self, key, value):
        value = _as_value(value)
        if isinstance(key, tuple):
            node = self
            for k in key[:-1]:
                if k not in node._data:
                    node._data[k] = TensorDict()
                node = node._data[k]
            node._data[key[-1]] = value
        else:
            self._data[key] = value

    def __contains__(self, key) -> bool:
        if isinstance(key, tuple):
            node = self
            for k in key:
                if not isinstance(node, TensorDict) or k not in node._data:
                    return False
                node = node._data[k]
            return True
        return key in self._data

    def __iter__(self):
        if len(self._batch_size) == 0:
            raise TypeError("iteration over a 0-d TensorDict")
        for i in range(self._batch_size[0]):
            yield self[i]

    def __len__(self) -> int:
        return self._batch_size[0] if len(self._batch_size) else 0

    def keys(self):
        return self._data.keys()

    def values(self):
        return self._data.values()

    def items(self):
        return self._data.items()

    def get(self, key, default=None):
        return self[key] if key in self else default

    # ---- batch bookkeeping -----------------------------------------------
    @property
    def batch_size(self) -> torch.Size:
        return self._batch_size

    def _leaf_shapes(self):
        for value in self._data.values():
            if isinstance(value, TensorDict):
                value.auto_batch_size_()
                if value._data:
                    yield tuple(value._batch_size)
            elif isinstance(value, torch.Tensor):
                yield tuple(value.shape)

    def auto_batch_size_(self) -> "TensorDict":
        shapes = list(self._leaf_shapes())
        if not shapes:
            self._batch_size = torch.Size([])
            return self
        common = []
        for dims in zip(*shapes):
            if all(d == dims[0] for d in dims):
                common.append(dims[0])
            else:
                break
        self._batch_size = torch.Size(common)
        return self

    def to(self, *args, **kwargs) -> "TensorDict":
        out = TensorDict(batch_size=self._batch_size)
        for key, value in self._data.items():
            if isinstance(value, (torch.Tensor, TensorDict)):
                out._data[key] = value.to(*args, **kwargs)
            else:
                out._data[key] = value
        return out

    def __repr__(self) -> str:
        def describe(v):
            if isinstance(v, torch.Tensor):
                return f"Tensor{tuple(v.shape)}:{str(v.dtype).replace('torch.', '')}"
            return repr(v)

        body = ", ".join(f"{k}={describe(v)}" for k, v in self._data.items())
        return f"TensorDict({body}, batch_size={list(self._batch_size)})"


def stack(items: list, dim: int = 0) -> TensorDict:
    """Stack TensorDicts with identical keys along a new leading axis."""
    if dim != 0:
        raise NotImplementedError("only dim=0 is supported")
    first = items[0]
    out = TensorDict()
    for key, value in first.items():
        column = [item[key] for item in items]
        if isinstance(value, TensorDict):
            out._data[key] = stack(column, dim=0)
        else:
            out._data[key] = torch.stack([torch.as_tensor(c) for c in column], dim=0)
    out._batch_size = torch.Size([len(items)]) + first.batch_size
    return out
