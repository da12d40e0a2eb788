"""Deterministic synthetic mesh generators emitting `triangle`-style dictionaries.

The reference builds every mesh with the `triangle` package (e.g.
examples/example_weak.py:46-49, tests/test_assembly.py:23-25) and hands the
resulting dict to `MeshTri` (torch_fem/mesh/abstract_mesh.py:31-41).  `triangle`
is not available offline, so workloads are generated here in the same format:

    vertices        (N_v, 2) float64
    vertex_markers  (N_v, 1) int32   1 on the boundary
    triangles       (N_c, 3) int32   counter-clockwise
    edges           (E, 2)   int32
    edge_markers    (E, 1)   int32   1 on the boundary
    neighbors       (N_c, 3) int32   cell opposite local vertex k, -1 = none

All generators are pure numpy and seeded, so a mesh is reproducible from its
arguments alone (BASELINE.json configs 1, 2, 4, 5).
"""

from __future__ import annotations

import numpy as np


def build_topology(vertices: np.ndarray, triangles: np.ndarray) -> dict:
    """Edges, boundary markers and `triangle -n` style neighbours of a mesh."""
    n_cells = triangles.shape[0]
    n_vertices = vertices.shape[0]
    tri = triangles.astype(np.int64)
    # local edge opposite local vertex k joins vertices (k+1, k+2)
    a = tri[:, [1, 2, 0]].reshape(-1)
    b = tri[:, [2, 0, 1]].reshape(-1)
    lo = np.minimum(a, b)
    hi = np.maximum(a, b)
    key = lo * n_vertices + hi
    order = np.argsort(key, kind="stable")
    sorted_key = key[order]
    first = np.ones(sorted_key.shape[0], dtype=bool)
    first[1:] = sorted_key[1:] != sorted_key[:-1]
    edge_id_sorted = np.cumsum(first) - 1
    n_edges = int(edge_id_sorted[-1]) + 1
    edge_of_half = np.empty_like(edge_id_sorted)
    edge_of_half[order] = edge_id_sorted
    counts = np.bincount(edge_id_sorted, minlength=n_edges)

    # order unique edges by first appearance in the cell list
    first_half = np.full(n_edges, key.shape[0], dtype=np.int64)
    np.minimum.at(first_half, edge_of_half, np.arange(key.shape[0]))
    appearance = np.argsort(first_half, kind="stable")
    rank = np.empty(n_edges, dtype=np.int64)
    rank[appearance] = np.arange(n_edges)
    edge_of_half = rank[edge_of_half]

    edges = np.empty((n_edges, 2), dtype=np.int32)
    edges[edge_of_half, 0] = lo
    edges[edge_of_half, 1] = hi
    edge_counts = np.empty(n_edges, dtype=np.int64)
    edge_counts[rank] = counts
    edge_markers = (edge_counts == 1).astype(np.int32).reshape(-1, 1)

    # neighbours: for each half-edge, the other cell sharing the edge
    cell_of_half = np.repeat(np.arange(n_cells), 3)
    sum_cells = np.zeros(n_edges, dtype=np.int64)
    np.add.at(sum_cells, edge_of_half, cell_of_half)
    other = sum_cells[edge_of_half] - cell_of_half
    neighbors = np.where(edge_counts[edge_of_half] == 2, other, -1).astype(np.int32).reshape(n_cells, 3)

    vertex_markers = np.zeros((n_vertices, 1), dtype=np.int32)
    boundary_edges = edges[edge_markers[:, 0] == 1]
    vertex_markers[boundary_edges.reshape(-1), 0] = 1
    return {
        "edges": edges,
        "edge_markers": edge_markers,
        "neighbors": neighbors,
        "vertex_markers": vertex_markers,
    }


def structured_rectangle(
    nx: int,
    ny: int,
    x0: float = 0.0,
    x1: float = 1.0,
    y0: float = 0.0,
    y1: float = 1.0,
    jitter: float = 0.0,
    seed: int = 1234,
    topology: bool = True,
    corners_first: bool = False,
) -> dict:
    """`nx` x `ny` squares, each split along the a-d diagonal into [a,b,d],[a,d,c].

    Vertex id is `j*(nx+1)+i`; interior vertices are displaced by
    `U(-jitter*h, jitter*h)` from `default_rng(seed)`.
    With `corners_first` the four corner vertices are renumbered 0..3 so that the
    first three are not collinear, as FracturesTri needs
    (torch_fem/mesh/fractures_tri.py:37-45).
    """
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    gx, gy = np.meshgrid(xs, ys, indexing="xy")
    vertices = np.stack([gx.reshape(-1), gy.reshape(-1)], axis=1).astype(np.float64)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        hx = (x1 - x0) / nx
        hy = (y1 - y0) / ny
        disp = rng.uniform(-jitter, jitter, size=vertices.shape) * np.array([hx, hy])
        interior = np.zeros((ny + 1, nx + 1), dtype=bool)
        interior[1:-1, 1:-1] = True
        vertices += disp * interior.reshape(-1, 1)

    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    a = (j * (nx + 1) + i).reshape(-1)
    b = a + 1
    c = a + (nx + 1)
    d = c + 1
    triangles = np.empty((2 * nx * ny, 3), dtype=np.int32)
    triangles[0::2] = np.stack([a, b, d], axis=1)
    triangles[1::2] = np.stack([a, d, c], axis=1)

    boundary = np.zeros((ny + 1, nx + 1), dtype=np.int32)
    boundary[0, :] = boundary[-1, :] = 1
    boundary[:, 0] = boundary[:, -1] = 1
    vertex_markers = boundary.reshape(-1, 1)

    if corners_first:
        n_v = vertices.shape[0]
        corners = np.array([0, nx, ny * (nx + 1), ny * (nx + 1) + nx])
        rest = np.setdiff1d(np.arange(n_v), corners, assume_unique=True)
        new_to_old = np.concatenate([corners, rest])
        old_to_new = np.empty(n_v, dtype=np.int64)
        old_to_new[new_to_old] = np.arange(n_v)
        vertices = vertices[new_to_old]
        vertex_markers = vertex_markers[new_to_old]
        triangles = old_to_new[triangles].astype(np.int32)

    mesh = {"vertices": vertices, "triangles": triangles, "vertex_markers": vertex_markers}
    if topology:
        topo = build_topology(vertices, triangles)
        mesh["edges"] = topo["edges"]
        mesh["edge_markers"] = topo["edge_markers"]
        mesh["neighbors"] = topo["neighbors"]
    return mesh


def delaunay_unit_square(n_interior: int = 60, seed: int = 0, boundary_per_side: int = 0) -> dict:
    """Delaunay triangulation of the unit-square corners plus random interior points."""
    from scipy.spatial import Delaunay  # host-side setup only

    rng = np.random.default_rng(seed)
    points = [np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])]
    if boundary_per_side:
        t = np.linspace(0.0, 1.0, boundary_per_side + 2)[1:-1]
        zeros, ones = np.zeros_like(t), np.ones_like(t)
        points += [np.stack([t, zeros], 1), np.stack([t, ones], 1), np.stack([zeros, t], 1), np.stack([ones, t], 1)]
    points.append(rng.random((n_interior, 2)))
    vertices = np.concatenate(points, axis=0)
    tri = Delaunay(vertices).simplices.astype(np.int64)
    p = vertices[tri]
    area2 = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 2, 0] - p[:, 0, 0]) * (
        p[:, 1, 1] - p[:, 0, 1]
    )
    keep = np.abs(area2) > 1e-14
    tri, area2 = tri[keep], area2[keep]
    flip = area2 < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    triangles = tri.astype(np.int32)
    mesh = {"vertices": vertices, "triangles": triangles}
    mesh.update(build_topology(vertices, triangles))
    return mesh


def permute_mesh(mesh: dict, seed: int = 7) -> dict:
    """Random renumbering of vertices and cells (locality stress)."""
    rng = np.random.default_rng(seed)
    n_v = mesh["vertices"].shape[0]
    n_c = mesh["triangles"].shape[0]
    new_to_old_v = rng.permutation(n_v)
    old_to_new_v = np.empty(n_v, dtype=np.int64)
    old_to_new_v[new_to_old_v] = np.arange(n_v)
    cell_perm = rng.permutation(n_c)
    out = {
        "vertices": mesh["vertices"][new_to_old_v],
        "vertex_markers": mesh["vertex_markers"][new_to_old_v],
        "triangles": old_to_new_v[mesh["triangles"][cell_perm]].astype(np.int32),
    }
    if "edges" in mesh:
        out.update(build_topology(out["vertices"], out["triangles"]))
    return out


def two_fracture_network(nx: int = 4, ny: int = 2, jitter: float = 0.0, seed: int = 3):
    """The two-plane DFN of examples/example_fractures_fem.py:31-64.

    Both fractures share one 2-D mesh of `[-1,1]x[0,1]` (an even `nx` puts a
    vertex column on the trace `x=0`); fracture 0 lies in `z=0`, fracture 1 in
    `x=0`.  Returns `(list_of_mesh_dicts, fractures_3d_data (2,4,3))`.
    """
    if nx % 2:
        raise ValueError("nx must be even so the trace x=0 is a mesh line")
    mesh = structured_rectangle(nx, ny, -1.0, 1.0, 0.0, 1.0, jitter=0.0, corners_first=True)
    if jitter > 0.0:
        # displace only vertices off the trace and off the boundary, along x,
        # by dyadic amounts so mapped 3-D trace vertices stay bit-identical
        rng = np.random.default_rng(seed)
        v = mesh["vertices"]
        free = (mesh["vertex_markers"][:, 0] == 0) & (v[:, 0] != 0.0)
        h = 2.0 / nx
        steps = rng.integers(-8, 9, size=v.shape[0]) / 64.0
        v[:, 0] += free * steps * h * jitter * 4
    data = np.array(
        [
            [[-1.0, 0.0, 0.0], [1.0, 0.0, 0.0], [-1.0, 1.0, 0.0], [1.0, 1.0, 0.0]],
            [[0.0, 0.0, -1.0], [0.0, 0.0, 1.0], [0.0, 1.0, -1.0], [0.0, 1.0, 1.0]],
        ]
    )
    return [mesh, {k: v.copy() for k, v in mesh.items()}], data


def seven_fracture_network(nx: int = 8, ny: int = 4):
    """Backbone `z=0` plus six planes `x = +-0.25, +-0.5, +-0.75`.

    Every fracture carries the same `nx` x `ny` mesh of `[-1,1]x[0,1]`; `nx` must
    be a multiple of 8 so each crossing plane meets the backbone on a mesh line
    and its own mesh has a vertex column at local `x=0`.  Coordinates are dyadic,
    so trace vertices coincide exactly after the affine map.
    """
    if nx % 8:
        raise ValueError("nx must be a multiple of 8")
    mesh = structured_rectangle(nx, ny, -1.0, 1.0, 0.0, 1.0, corners_first=True)
    data = [[[-1.0, 0.0, 0.0], [1.0, 0.0, 0.0], [-1.0, 1.0, 0.0], [1.0, 1.0, 0.0]]]
    for x in (-0.75, -0.5, -0.25, 0.25, 0.5, 0.75):
        data.append([[x, 0.0, -1.0], [x, 0.0, 1.0], [x, 1.0, -1.0], [x, 1.0, 1.0]])
    meshes = [{k: v.copy() for k, v in mesh.items()} for _ in range(7)]
    return meshes, np.array(data)


def generate_patches_info(levels: int):
    """Centres and radii of the `4**levels` uniform patches of the unit square.

    Same enumeration as examples/example_patches.py:49-70.
    """
    centers = [(0.5, 0.5)]
    radius = [0.5]
    for _ in range(levels):
        new_centers, new_radius = [], []
        for (cx, cy), r in zip(centers, radius):
            half = r / 2
            new_centers += [(cx - half, cy - half), (cx - half, cy + half), (cx + half, cy - half), (cx + half, cy + half)]
            new_radius += [half] * 4
        centers, radius = new_centers, new_radius
    return np.asarray(centers, dtype=np.float64), np.asarray(radius, dtype=np.float64).reshape(-1, 1)
