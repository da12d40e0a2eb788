// FE interpolation to cell / interior-edge quadrature points and the jump estimator.
// Reference: basis/basis.py:98-159, basis/fracture_basis.py:212-257,
// element/abstract_element.py:18-26, examples/example_jump.py:75-87.
#include "common.cuh"

namespace tfem {

template <typename T>
__global__ void __launch_bounds__(256) interp_cells_kernel(int n_el, const int32_t* __restrict__ dof_conn,
                                                           const T* __restrict__ v_grad, int d,
                                                           const QuadT<T> quad, const T* __restrict__ u,
                                                           T* __restrict__ val, T* __restrict__ grad) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
  const T u0 = __ldg(u + __ldg(dof_conn + 3 * (int64_t)e + 0));
  const T u1 = __ldg(u + __ldg(dof_conn + 3 * (int64_t)e + 1));
  const T u2 = __ldg(u + __ldg(dof_conn + 3 * (int64_t)e + 2));
  if (val) {
    for (int q = 0; q < quad.n_q; ++q)
      val[(int64_t)e * quad.n_q + q] = u0 * quad.l0[q] + u1 * quad.l1[q] + u2 * quad.l2[q];
  }
  if (grad) {
    const T* g = v_grad + (int64_t)e * 3 * d;
    for (int c = 0; c < d; ++c)
      grad[(int64_t)e * d + c] = u0 * __ldg(g + c) + u1 * __ldg(g + d + c) + u2 * __ldg(g + 2 * d + c);
  }
}

// One thread per (edge, side).
template <typename T, int D>
__global__ void __launch_bounds__(256) interp_edges_kernel(
    int n_edge, int n_edge_per_mesh, int n_el_per_mesh, const int32_t* __restrict__ edge_cells,
    const int32_t* __restrict__ conn, const T* __restrict__ first_vertex, const T* __restrict__ inv_jac,
    const T* __restrict__ x_q, int n_q, const T* __restrict__ u, T* __restrict__ val, T* __restrict__ grad) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * n_edge) return;
  const int edge = t >> 1;
  const int mesh = edge / n_edge_per_mesh;
  const int64_t cell = (int64_t)mesh * n_el_per_mesh + __ldg(edge_cells + t);
  const T u0 = __ldg(u + __ldg(conn + 3 * cell + 0));
  const T u1 = __ldg(u + __ldg(conn + 3 * cell + 1));
  const T u2 = __ldg(u + __ldg(conn + 3 * cell + 2));
  T inv[2][D], p0[D];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    inv[0][c] = __ldg(inv_jac + cell * 2 * D + c);
    inv[1][c] = __ldg(inv_jac + cell * 2 * D + D + c);
    p0[c] = __ldg(first_vertex + cell * D + c);
  }
  if (val) {
    for (int q = 0; q < n_q; ++q) {
      T r0 = T(0), r1 = T(0);  // (x_q - p0) @ J^-T  (abstract_element.py:18-26)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const T dxc = __ldg(x_q + ((int64_t)edge * n_q + q) * D + c) - p0[c];
        r0 += dxc * inv[0][c];
        r1 += dxc * inv[1][c];
      }
      const T l0 = T(1) - r0 - r1;  // element_tri.py:23-26
      val[(int64_t)t * n_q + q] = u0 * l0 + u1 * r0 + u2 * r1;
    }
  }
  if (grad) {
#pragma unroll
    for (int c = 0; c < D; ++c)
      grad[(int64_t)t * D + c] = u0 * (-inv[0][c] - inv[1][c]) + u1 * inv[0][c] + u2 * inv[1][c];
  }
}

// local[e,i] = sum_q val_bar[e,q] l_i(q) + sum_c grad_bar[e,c] v_grad[e,i,c]
template <typename T>
__global__ void __launch_bounds__(256) interp_cells_bwd_kernel(int n_el, const T* __restrict__ v_grad, int d,
                                                               const QuadT<T> quad, const T* __restrict__ val_bar,
                                                               const T* __restrict__ grad_bar, T* __restrict__ local) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
  T a0 = T(0), a1 = T(0), a2 = T(0);
  if (val_bar) {
    for (int q = 0; q < quad.n_q; ++q) {
      const T vb = __ldg(val_bar + (int64_t)e * quad.n_q + q);
      a0 += vb * quad.l0[q];
      a1 += vb * quad.l1[q];
      a2 += vb * quad.l2[q];
    }
  }
  if (grad_bar) {
    const T* g = v_grad + (int64_t)e * 3 * d;
    for (int c = 0; c < d; ++c) {
      const T gb = __ldg(grad_bar + (int64_t)e * d + c);
      a0 += gb * __ldg(g + c);
      a1 += gb * __ldg(g + d + c);
      a2 += gb * __ldg(g + 2 * d + c);
    }
  }
  local[3 * (int64_t)e + 0] = a0;
  local[3 * (int64_t)e + 1] = a1;
  local[3 * (int64_t)e + 2] = a2;
}

// One thread per (edge, side): same geometry as interp_edges_kernel, transposed.
template <typename T, int D>
__global__ void __launch_bounds__(256) interp_edges_bwd_kernel(
    int n_edge, int n_edge_per_mesh, int n_el_per_mesh, const int32_t* __restrict__ edge_cells,
    const T* __restrict__ first_vertex, const T* __restrict__ inv_jac, const T* __restrict__ x_q, int n_q,
    const T* __restrict__ val_bar, const T* __restrict__ grad_bar, T* __restrict__ local) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * n_edge) return;
  const int edge = t >> 1;
  const int mesh = edge / n_edge_per_mesh;
  const int64_t cell = (int64_t)mesh * n_el_per_mesh + __ldg(edge_cells + t);
  T inv[2][D], p0[D];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    inv[0][c] = __ldg(inv_jac + cell * 2 * D + c);
    inv[1][c] = __ldg(inv_jac + cell * 2 * D + D + c);
    p0[c] = __ldg(first_vertex + cell * D + c);
  }
  T a0 = T(0), a1 = T(0), a2 = T(0);
  if (val_bar) {
    for (int q = 0; q < n_q; ++q) {
      T r0 = T(0), r1 = T(0);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const T dxc = __ldg(x_q + ((int64_t)edge * n_q + q) * D + c) - p0[c];
        r0 += dxc * inv[0][c];
        r1 += dxc * inv[1][c];
      }
      const T vb = __ldg(val_bar + (int64_t)t * n_q + q);
      a0 += vb * (T(1) - r0 - r1);
      a1 += vb * r0;
      a2 += vb * r1;
    }
  }
  if (grad_bar) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const T gb = __ldg(grad_bar + (int64_t)t * D + c);
      a0 += gb * (-inv[0][c] - inv[1][c]);
      a1 += gb * inv[0][c];
      a2 += gb * inv[1][c];
    }
  }
  local[3 * (int64_t)t + 0] = a0;
  local[3 * (int64_t)t + 1] = a1;
  local[3 * (int64_t)t + 2] = a2;
}

template <typename T>
__global__ void __launch_bounds__(256) edge_jump_kernel(int n_edge, int d, int n_q,
                                                        const T* __restrict__ grad_edges,
                                                        const T* __restrict__ normals,
                                                        const T* __restrict__ h_e, const T* __restrict__ dx,
                                                        T* __restrict__ eta) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edge) return;
  T plus = T(0), minus = T(0);
  for (int c = 0; c < d; ++c) {
    const T n = __ldg(normals + (int64_t)e * d + c);
    plus += __ldg(grad_edges + (int64_t)e * 2 * d + c) * n;
    minus += __ldg(grad_edges + (int64_t)e * 2 * d + d + c) * (-n);
  }
  const T jump = plus + minus;
  const T integrand = __ldg(h_e + e) * (jump * jump);
  T acc = T(0);
  for (int q = 0; q < n_q; ++q) acc += integrand * __ldg(dx + (int64_t)e * n_q + q);
  eta[e] = acc;
}

template <typename T>
int interp_cells(int64_t n_el, const int32_t* dof_conn, const T* v_grad, int d, int quad_order, const T* u,
                 T* val, T* grad, void* stream) {
  if (n_el < 0 || (d != 2 && d != 3)) return TFEM_ERR_BAD_ARG;
  if (n_el == 0) return TFEM_OK;
  if (!dof_conn || !u || (grad && !v_grad)) return TFEM_ERR_BAD_ARG;
  if (n_el > kMaxIndex / 9) return TFEM_ERR_TOO_LARGE;
  if (tri_n_q(quad_order) == 0) return TFEM_ERR_UNSUPPORTED;
  interp_cells_kernel<T><<<blocks_for(n_el, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (int)n_el, dof_conn, v_grad, d, make_quad<T>(quad_order), u, val, grad);
  return check_launch();
}

template <typename T>
int interp_edges(int64_t n_edge, int64_t n_edge_per_mesh, int64_t n_el_per_mesh, const int32_t* edge_cells,
                 const int32_t* conn, const T* first_vertex, const T* inv_jac, int d, const T* x_q, int n_q,
                 const T* u, T* val, T* grad, void* stream) {
  if (n_edge < 0 || n_edge_per_mesh <= 0 || n_el_per_mesh <= 0 || (d != 2 && d != 3) || n_q <= 0)
    return TFEM_ERR_BAD_ARG;
  if (n_edge == 0) return TFEM_OK;
  if (!edge_cells || !conn || !first_vertex || !inv_jac || !u || (val && !x_q)) return TFEM_ERR_BAD_ARG;
  if (n_edge > kMaxIndex / 16) return TFEM_ERR_TOO_LARGE;
  auto s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = blocks_for(2 * n_edge, 256);
  if (d == 2)
    interp_edges_kernel<T, 2><<<blocks, 256, 0, s>>>((int)n_edge, (int)n_edge_per_mesh, (int)n_el_per_mesh,
                                                     edge_cells, conn, first_vertex, inv_jac, x_q, n_q, u, val, grad);
  else
    interp_edges_kernel<T, 3><<<blocks, 256, 0, s>>>((int)n_edge, (int)n_edge_per_mesh, (int)n_el_per_mesh,
                                                     edge_cells, conn, first_vertex, inv_jac, x_q, n_q, u, val, grad);
  return check_launch();
}

template <typename T>
int interp_cells_bwd(int64_t n_el, const T* v_grad, int d, int quad_order, const T* val_bar, const T* grad_bar,
                     T* local, void* stream) {
  if (n_el < 0 || (d != 2 && d != 3)) return TFEM_ERR_BAD_ARG;
  if (n_el == 0) return TFEM_OK;
  if (!local || (grad_bar && !v_grad)) return TFEM_ERR_BAD_ARG;
  if (n_el > kMaxIndex / 9) return TFEM_ERR_TOO_LARGE;
  if (tri_n_q(quad_order) == 0) return TFEM_ERR_UNSUPPORTED;
  interp_cells_bwd_kernel<T><<<blocks_for(n_el, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (int)n_el, v_grad, d, make_quad<T>(quad_order), val_bar, grad_bar, local);
  return check_launch();
}

template <typename T>
int interp_edges_bwd(int64_t n_edge, int64_t n_edge_per_mesh, int64_t n_el_per_mesh, const int32_t* edge_cells,
                     const T* first_vertex, const T* inv_jac, int d, const T* x_q, int n_q, const T* val_bar,
                     const T* grad_bar, T* local, void* stream) {
  if (n_edge < 0 || n_edge_per_mesh <= 0 || n_el_per_mesh <= 0 || (d != 2 && d != 3) || n_q <= 0)
    return TFEM_ERR_BAD_ARG;
  if (n_edge == 0) return TFEM_OK;
  if (!edge_cells || !first_vertex || !inv_jac || !local || (val_bar && !x_q)) return TFEM_ERR_BAD_ARG;
  if (n_edge > kMaxIndex / 16) return TFEM_ERR_TOO_LARGE;
  auto s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = blocks_for(2 * n_edge, 256);
  if (d == 2)
    interp_edges_bwd_kernel<T, 2><<<blocks, 256, 0, s>>>((int)n_edge, (int)n_edge_per_mesh, (int)n_el_per_mesh, edge_cells,
                                                         first_vertex, inv_jac, x_q, n_q, val_bar, grad_bar, local);
  else
    interp_edges_bwd_kernel<T, 3><<<blocks, 256, 0, s>>>((int)n_edge, (int)n_edge_per_mesh, (int)n_el_per_mesh, edge_cells,
                                                         first_vertex, inv_jac, x_q, n_q, val_bar, grad_bar, local);
  return check_launch();
}

template <typename T>
int edge_jump(int64_t n_edge, int d, int n_q, const T* grad_edges, const T* normals, const T* h_e, const T* dx,
              T* eta, void* stream) {
  if (n_edge < 0 || (d != 2 && d != 3) || n_q <= 0) return TFEM_ERR_BAD_ARG;
  if (n_edge == 0) return TFEM_OK;
  if (!grad_edges || !normals || !h_e || !dx || !eta) return TFEM_ERR_BAD_ARG;
  if (n_edge > kMaxIndex / 16) return TFEM_ERR_TOO_LARGE;
  edge_jump_kernel<T><<<blocks_for(n_edge, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (int)n_edge, d, n_q, grad_edges, normals, h_e, dx, eta);
  return check_launch();
}

}  // namespace tfem

#define TFEM_INTERP_API(T, SUF)                                                                     \
  extern "C" int tfem_interp_cells_##SUF(int64_t n_el, const int32_t* dof_conn, const T* v_grad,    \
                                         int d, int quad_order, const T* u, T* val, T* grad,        \
                                         void* stream) {                                            \
    return tfem::interp_cells<T>(n_el, dof_conn, v_grad, d, quad_order, u, val, grad, stream);      \
  }                                                                                                 \
  extern "C" int tfem_interp_edges_##SUF(                                                           \
      int64_t n_edge, int64_t n_edge_per_mesh, int64_t n_el_per_mesh, const int32_t* edge_cells,    \
      const int32_t* conn, const T* first_vertex, const T* inv_jac, int d, const T* x_q, int n_q,   \
      const T* u, T* val, T* grad, void* stream) {                                                  \
    return tfem::interp_edges<T>(n_edge, n_edge_per_mesh, n_el_per_mesh, edge_cells, conn,          \
                                 first_vertex, inv_jac, d, x_q, n_q, u, val, grad, stream);         \
  }                                                                                                 \
  extern "C" int tfem_interp_cells_bwd_##SUF(int64_t n_el, const T* v_grad, int d, int quad_order,   \
                                             const T* val_bar, const T* grad_bar, T* local,         \
                                             void* stream) {                                        \
    return tfem::interp_cells_bwd<T>(n_el, v_grad, d, quad_order, val_bar, grad_bar, local, stream);\
  }                                                                                                 \
  extern "C" int tfem_interp_edges_bwd_##SUF(                                                       \
      int64_t n_edge, int64_t n_edge_per_mesh, int64_t n_el_per_mesh, const int32_t* edge_cells,    \
      const T* first_vertex, const T* inv_jac, int d, const T* x_q, int n_q, const T* val_bar,      \
      const T* grad_bar, T* local, void* stream) {                                                  \
    return tfem::interp_edges_bwd<T>(n_edge, n_edge_per_mesh, n_el_per_mesh, edge_cells,            \
                                     first_vertex, inv_jac, d, x_q, n_q, val_bar, grad_bar, local,  \
                                     stream);                                                       \
  }                                                                                                 \
  extern "C" int tfem_edge_jump_##SUF(int64_t n_edge, int d, int n_q, const T* grad_edges,          \
                                      const T* normals, const T* h_e, const T* dx, T* eta,          \
                                      void* stream) {                                               \
    return tfem::edge_jump<T>(n_edge, d, n_q, grad_edges, normals, h_e, dx, eta, stream);           \
  }

TFEM_INTERP_API(double, f64)
TFEM_INTERP_API(float, f32)
