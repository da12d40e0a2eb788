// Edge topology on the device: the sort-based edge <-> cell adjacency that replaces
// AbstractMesh._compute_interior_and_boundary_edges / _compute_cells_4_edges (mesh/abstract_mesh.py:104-255; batched:
// mesh/meshes_tri.py:54-123), whose O(E C) broadcast match and per-mesh Python loops do not scale.
//   tfem_half_edges           every cell edge as a key  mesh * V^2 + min(v) * V + max(v), stably sorted (CUB radix sort over
//                             the bits a key can have) with its cell as payload; unique keys and their incidence counts
//   tfem_edge_cells           the cells adjacent to each edge of an edge list, by binary search in the sorted half edges;
//                             the sides come in increasing cell id ("aligned": the same order for every consumer)
//   tfem_interior_edge_geometry   end points, length and unit normal of each interior edge, the normal oriented from the
//                             first listed cell's centroid towards the second's (abstract_mesh.py:143-162)
#include <cub/cub.cuh>

#include "common.cuh"

namespace tfem {

__global__ void half_edge_keys_kernel(int64_t n_cells_total, int64_t n_cells, int64_t n_vert, const int32_t* __restrict__ conn, int p00, int p01,
                                      int p10, int p11, int p20, int p21, int64_t* __restrict__ keys, int32_t* __restrict__ cells) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells_total) return;
  const int64_t mesh = c / n_cells;
  const int32_t v[3] = {__ldg(conn + 3 * c), __ldg(conn + 3 * c + 1), __ldg(conn + 3 * c + 2)};
  const int pa[3] = {p00, p10, p20}, pb[3] = {p01, p11, p21};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int64_t a = v[pa[k]], b = v[pb[k]];
    const int64_t lo = a < b ? a : b, hi = a < b ? b : a;
    keys[3 * c + k] = mesh * n_vert * n_vert + lo * n_vert + hi;
    cells[3 * c + k] = (int32_t)(c - mesh * n_cells);
  }
}

// status bits: 1 = an edge of the list belongs to no cell, 2 = an edge listed with two sides has a single adjacent cell
__global__ void edge_cells_kernel(int64_t n_edges_total, int64_t n_edges, int64_t n_vert, const int32_t* __restrict__ edge_vertices, int n_sides,
                                  const int64_t* __restrict__ he_sorted, const int32_t* __restrict__ cell_sorted, int64_t n_half,
                                  int32_t* __restrict__ cells_out, int32_t* __restrict__ status) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges_total) return;
  const int64_t mesh = e / n_edges;
  const int64_t a = __ldg(edge_vertices + 2 * e), b = __ldg(edge_vertices + 2 * e + 1);
  const int64_t key = mesh * n_vert * n_vert + (a < b ? a : b) * n_vert + (a < b ? b : a);
  int64_t lo = 0, hi = n_half;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (he_sorted[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  int flags = 0;
  if (lo >= n_half || he_sorted[lo] != key) flags |= 1;
  for (int s = 0; s < n_sides; ++s) {
    const int64_t pos = lo + s < n_half ? lo + s : n_half - 1;
    if (s > 0 && he_sorted[pos] != key) flags |= 2;
    cells_out[e * n_sides + s] = cell_sorted[pos];
  }
  if (flags) atomicOr(status, flags);
}

template <typename T>
__global__ void interior_edge_geometry_kernel(int64_t n_edges_total, int64_t n_edges, int64_t n_vert, int64_t n_cells, const T* __restrict__ coords,
                                              const int32_t* __restrict__ conn, const int32_t* __restrict__ edge_vertices,
                                              const int32_t* __restrict__ edge_cells, T* __restrict__ x_out, T* __restrict__ length_out,
                                              T* __restrict__ normal_out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges_total) return;
  const int64_t mesh = e / n_edges;
  const T* xy = coords + 2 * mesh * n_vert;
  const int64_t a = __ldg(edge_vertices + 2 * e), b = __ldg(edge_vertices + 2 * e + 1);
  T ax, ay, bx, by;
  load_xy(xy, a, ax, ay);
  load_xy(xy, b, bx, by);
  x_out[4 * e] = ax; x_out[4 * e + 1] = ay; x_out[4 * e + 2] = bx; x_out[4 * e + 3] = by;
  const T vx = bx - ax, vy = by - ay;
  const T len = sqrt(vx * vx + vy * vy);
  T nx = -vy / len, ny = vx / len;
  // centroid of each adjacent cell: mean of its three vertices, summed in vertex order like torch's mean(dim=-2)
  T cx[2], cy[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int32_t* tri = conn + 3 * (mesh * n_cells + __ldg(edge_cells + 2 * e + s));
    T x0, y0, x1, y1, x2, y2;
    load_xy(xy, (int64_t)__ldg(tri), x0, y0);
    load_xy(xy, (int64_t)__ldg(tri + 1), x1, y1);
    load_xy(xy, (int64_t)__ldg(tri + 2), x2, y2);
    cx[s] = ((x0 + x1) + x2) / T(3);
    cy[s] = ((y0 + y1) + y2) / T(3);
  }
  const T towards = nx * (cx[1] - cx[0]) + ny * (cy[1] - cy[0]);
  if (towards < T(0)) {
    nx = -nx;
    ny = -ny;
  }
  length_out[e] = len;
  normal_out[2 * e] = nx;
  normal_out[2 * e + 1] = ny;
}

inline int topo_bits_for(int64_t v) {
  int b = 1;
  while (b < 63 && (int64_t(1) << b) < v) ++b;
  return b;
}
inline size_t topo_align(size_t v) { return (v + 255) & ~size_t(255); }

struct HalfEdgeLayout {
  size_t keys_in, cells_in, cub_temp, cub_bytes, total;
};

inline int half_edge_layout(int64_t n_half, HalfEdgeLayout* out) {
  size_t sort_bytes = 0, rle_bytes = 0;
  if (cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const int64_t*)nullptr, (int64_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr,
                                      (int)n_half) != cudaSuccess)
    return TFEM_ERR_LAUNCH;
  if (cub::DeviceRunLengthEncode::Encode(nullptr, rle_bytes, (const int64_t*)nullptr, (int64_t*)nullptr, (int32_t*)nullptr, (int64_t*)nullptr,
                                         (int)n_half) != cudaSuccess)
    return TFEM_ERR_LAUNCH;
  size_t at = 0;
  out->keys_in = at;  at += topo_align(8 * (size_t)n_half);
  out->cells_in = at; at += topo_align(4 * (size_t)n_half);
  out->cub_bytes = sort_bytes > rle_bytes ? sort_bytes : rle_bytes;
  out->cub_temp = at; at += topo_align(out->cub_bytes);
  out->total = at;
  return TFEM_OK;
}

}  // namespace tfem

extern "C" int tfem_half_edges_workspace(int64_t n_mesh, int64_t n_cells, int64_t* bytes) {
  if (n_mesh < 0 || n_cells < 0 || !bytes) return TFEM_ERR_BAD_ARG;
  if (3 * n_mesh * n_cells > tfem::kMaxIndex) return TFEM_ERR_TOO_LARGE;
  tfem::HalfEdgeLayout lay{};
  const int64_t n_half = 3 * n_mesh * n_cells;
  const int status = tfem::half_edge_layout(n_half > 0 ? n_half : 1, &lay);
  if (status != TFEM_OK) return status;
  *bytes = (int64_t)lay.total;
  return TFEM_OK;
}

extern "C" int tfem_half_edges(int64_t n_mesh, int64_t n_cells, int64_t n_vert, const int32_t* conn, const int32_t* local_pairs, void* workspace,
                               int64_t workspace_bytes, int64_t* he_sorted, int32_t* cell_sorted, int64_t* unique_keys, int32_t* counts,
                               int64_t* n_unique, void* stream) {
  using namespace tfem;
  if (n_mesh <= 0 || n_cells <= 0 || n_vert <= 0) return TFEM_ERR_BAD_ARG;
  if (!conn || !local_pairs || !workspace || !he_sorted || !cell_sorted || !unique_keys || !counts || !n_unique) return TFEM_ERR_BAD_ARG;
  const int64_t n_total = n_mesh * n_cells, n_half = 3 * n_total;
  if (n_half > kMaxIndex) return TFEM_ERR_TOO_LARGE;
  // keys must fit 63 bits: mesh * V^2 + lo * V + hi < n_mesh * V^2
  if ((double)n_mesh * (double)n_vert * (double)n_vert >= 9.0e18) return TFEM_ERR_TOO_LARGE;
  for (int k = 0; k < 6; ++k)
    if (local_pairs[k] < 0 || local_pairs[k] > 2) return TFEM_ERR_BAD_ARG;
  HalfEdgeLayout lay{};
  int status = half_edge_layout(n_half, &lay);
  if (status != TFEM_OK) return status;
  if ((size_t)workspace_bytes < lay.total) return TFEM_ERR_BAD_ARG;
  auto s = static_cast<cudaStream_t>(stream);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  int64_t* keys_in = reinterpret_cast<int64_t*>(ws + lay.keys_in);
  int32_t* cells_in = reinterpret_cast<int32_t*>(ws + lay.cells_in);
  half_edge_keys_kernel<<<blocks_for(n_total, 256), 256, 0, s>>>(n_total, n_cells, n_vert, conn, local_pairs[0], local_pairs[1], local_pairs[2],
                                                                 local_pairs[3], local_pairs[4], local_pairs[5], keys_in, cells_in);
  size_t cub_bytes = lay.cub_bytes;
  const int key_bits = topo_bits_for(n_mesh * n_vert * n_vert);
  if (cub::DeviceRadixSort::SortPairs(ws + lay.cub_temp, cub_bytes, keys_in, he_sorted, cells_in, cell_sorted, (int)n_half, 0, key_bits, s) !=
      cudaSuccess)
    return TFEM_ERR_LAUNCH;
  cub_bytes = lay.cub_bytes;
  if (cub::DeviceRunLengthEncode::Encode(ws + lay.cub_temp, cub_bytes, he_sorted, unique_keys, counts, n_unique, (int)n_half, s) != cudaSuccess)
    return TFEM_ERR_LAUNCH;
  return check_launch();
}

extern "C" int tfem_edge_cells(int64_t n_mesh, int64_t n_edges, int64_t n_vert, const int32_t* edge_vertices, int n_sides, const int64_t* he_sorted,
                               const int32_t* cell_sorted, int64_t n_half, int32_t* cells, int32_t* status, void* stream) {
  using namespace tfem;
  if (n_mesh < 0 || n_edges < 0 || n_sides < 1 || n_sides > 2 || n_half <= 0) return TFEM_ERR_BAD_ARG;
  if (n_mesh * n_edges == 0) return TFEM_OK;
  if (!edge_vertices || !he_sorted || !cell_sorted || !cells || !status) return TFEM_ERR_BAD_ARG;
  edge_cells_kernel<<<blocks_for(n_mesh * n_edges, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n_mesh * n_edges, n_edges, n_vert, edge_vertices, n_sides, he_sorted, cell_sorted, n_half, cells, status);
  return check_launch();
}

#define TFEM_TOPOLOGY_API(T, SUF)                                                                                                      \
  extern "C" int tfem_interior_edge_geometry_##SUF(int64_t n_mesh, int64_t n_edges, int64_t n_vert, int64_t n_cells, const T* coords,  \
                                                   const int32_t* conn, const int32_t* edge_vertices, const int32_t* edge_cells,       \
                                                   T* x, T* length, T* normal, void* stream) {                                         \
    if (n_mesh < 0 || n_edges < 0) return TFEM_ERR_BAD_ARG;                                                                            \
    if (n_mesh * n_edges == 0) return TFEM_OK;                                                                                         \
    if (!coords || !conn || !edge_vertices || !edge_cells || !x || !length || !normal) return TFEM_ERR_BAD_ARG;                        \
    tfem::interior_edge_geometry_kernel<T><<<tfem::blocks_for(n_mesh * n_edges, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(    \
        n_mesh * n_edges, n_edges, n_vert, n_cells, coords, conn, edge_vertices, edge_cells, x, length, normal);                       \
    return tfem::check_launch();                                                                                                       \
  }

TFEM_TOPOLOGY_API(double, f64)
TFEM_TOPOLOGY_API(float, f32)
