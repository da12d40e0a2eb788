"""Golden vectors of the reference's P2 shape functions (`ElementTri(2, k).compute_shape_functions`,
/root/reference/torch_fem/element/element_tri.py:43-70) on a few seeded Jacobians.  Run in the build container:
    python tests/golden/make_golden_p2.py
Uses the stand-ins of make_golden.py to import the UNMODIFIED reference package."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "p2_shape.npz")


def main():
    make_golden.import_reference()
    from torch_fem.element.element_tri import ElementTri

    torch.set_default_dtype(torch.float64)
    out = {}
    for order in (2, 3, 4):
        element = ElementTri(2, order)
        generator = torch.Generator().manual_seed(order)
        jac = torch.eye(2) + 0.3 * torch.randn(5, 1, 2, 2, generator=generator)  # (N,1,2,2): one inverse map per element
        inv = torch.linalg.inv(jac)
        bar = element.compute_barycentric_coordinates(element.gaussian_nodes)
        v, v_grad = element.compute_shape_functions(bar, inv)
        out[f"o{order}_inv"] = inv.numpy()
        out[f"o{order}_bar"] = bar.numpy()
        out[f"o{order}_v"] = v.numpy()
        out[f"o{order}_v_grad"] = v_grad.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
