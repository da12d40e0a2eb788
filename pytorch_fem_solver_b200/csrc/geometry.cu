// Per-element / per-edge geometry: replaces the tensor program of
// AbstractBasis._compute_integral_values (reference basis/abstract_basis.py:42-63).
#include "common.cuh"

namespace tfem {

template <typename T>
struct FracPtrs {
  const T* jac;  // [n_mesh,3,2]
  const T* inv;  // [n_mesh,2,3]
  const T* det;  // [n_mesh]
  const T* t;    // [n_mesh,3]
};

// One thread per element.  conn is read as three coalesced int32 streams, coordinates as one
// 16 B (f64) / 8 B (f32) request per vertex served by L1/L2 (neighbouring elements share vertices).
template <typename T, bool FRAC>
__global__ void __launch_bounds__(256) tri_geometry_kernel(
    int n_el, int n_el_per_mesh, int n_vert_per_mesh, const T* __restrict__ coords,
    const int32_t* __restrict__ conn, const QuadT<T> quad, const FracPtrs<T> frac,
    T* __restrict__ inv_jac, T* __restrict__ v_grad, T* __restrict__ x_q, T* __restrict__ dx) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
  const int mesh = e / n_el_per_mesh;
  const int64_t voff = (int64_t)mesh * n_vert_per_mesh;
  const int v0 = __ldg(conn + 3 * (int64_t)e + 0);
  const int v1 = __ldg(conn + 3 * (int64_t)e + 1);
  const int v2 = __ldg(conn + 3 * (int64_t)e + 2);
  T x0, y0, x1, y1, x2, y2;
  load_xy(coords, voff + v0, x0, y0);
  load_xy(coords, voff + v1, x1, y1);
  load_xy(coords, voff + v2, x2, y2);
  const TriGeom<T> g = tri_geom(x0, y0, x1, y1, x2, y2);

  // physical gradients, row i = grad(phi_i) (element_tri.py:37-41: barycentric_grad @ J^-1)
  T gr[3][2];
  gr[0][0] = -g.i00 - g.i10;
  gr[0][1] = -g.i01 - g.i11;
  gr[1][0] = g.i00;
  gr[1][1] = g.i01;
  gr[2][0] = g.i10;
  gr[2][1] = g.i11;

  T detf = T(1);
  if constexpr (FRAC) {
    T jf[3][2], jinv[2][3], tf[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      jf[r][0] = __ldg(frac.jac + 6 * mesh + 2 * r);
      jf[r][1] = __ldg(frac.jac + 6 * mesh + 2 * r + 1);
      tf[r] = __ldg(frac.t + 3 * mesh + r);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) jinv[a][c] = __ldg(frac.inv + 6 * mesh + 3 * a + c);
    detf = __ldg(frac.det + mesh);
    if (inv_jac) {  // fracture_basis.py:24-26: J^-1 @ J_f^+  -> (2,3)
      const T i2[2][2] = {{g.i00, g.i01}, {g.i10, g.i11}};
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) inv_jac[6 * (int64_t)e + 3 * r + c] = i2[r][0] * jinv[0][c] + i2[r][1] * jinv[1][c];
    }
    if (v_grad) {  // fracture_basis.py:20-22
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) v_grad[9 * (int64_t)e + 3 * i + c] = gr[i][0] * jinv[0][c] + gr[i][1] * jinv[1][c];
    }
    if (x_q) {  // fracture_basis.py:199-207
      for (int q = 0; q < quad.n_q; ++q) {
        const T px = quad.l0[q] * x0 + quad.l1[q] * x1 + quad.l2[q] * x2;
        const T py = quad.l0[q] * y0 + quad.l1[q] * y1 + quad.l2[q] * y2;
#pragma unroll
        for (int r = 0; r < 3; ++r) x_q[((int64_t)e * quad.n_q + q) * 3 + r] = jf[r][0] * px + jf[r][1] * py + tf[r];
      }
    }
  } else {
    if (inv_jac) {
      using V = typename Vec2<T>::type;
      V* o = reinterpret_cast<V*>(inv_jac) + 2 * (int64_t)e;
      o[0] = V{g.i00, g.i01};
      o[1] = V{g.i10, g.i11};
    }
    if (v_grad) {
      using V = typename Vec2<T>::type;
      V* o = reinterpret_cast<V*>(v_grad) + 3 * (int64_t)e;
      o[0] = V{gr[0][0], gr[0][1]};
      o[1] = V{gr[1][0], gr[1][1]};
      o[2] = V{gr[2][0], gr[2][1]};
    }
    if (x_q) {  // basis.py:90-91
      using V = typename Vec2<T>::type;
      V* o = reinterpret_cast<V*>(x_q) + (int64_t)e * quad.n_q;
      for (int q = 0; q < quad.n_q; ++q) {
        const T px = quad.l0[q] * x0 + quad.l1[q] * x1 + quad.l2[q] * x2;
        const T py = quad.l0[q] * y0 + quad.l1[q] * y1 + quad.l2[q] * y2;
        o[q] = V{px, py};
      }
    }
  }
  if (dx) {  // basis.py:93-96 (and fracture_basis.py:189-197)
    for (int q = 0; q < quad.n_q; ++q) dx[(int64_t)e * quad.n_q + q] = quad.w[q] * g.det * detf;
  }
}

// One thread per interior edge (interior_edges_basis.py:63-72, element_line.py:45-73).
template <typename T, bool FRAC>
__global__ void __launch_bounds__(256) edge_geometry_kernel(
    int n_edge, int n_edge_per_mesh, const T* __restrict__ edge_coords, int n_q, T node,
    T w0, T w1, T w2, const T* __restrict__ frac_jac, const T* __restrict__ frac_det,
    const T* __restrict__ frac_t, T* __restrict__ inv_jac, T* __restrict__ v_grad,
    T* __restrict__ x_q, T* __restrict__ dx) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edge) return;
  const int mesh = e / n_edge_per_mesh;
  T ax, ay, bx, by;
  load_xy(edge_coords, 2 * (int64_t)e, ax, ay);
  load_xy(edge_coords, 2 * (int64_t)e + 1, bx, by);
  // J = X^T @ [[-1/2],[1/2]]  (interior_edges_basis.py:63-64)
  const T jx = T(-0.5) * ax + T(0.5) * bx;
  const T jy = T(-0.5) * ay + T(0.5) * by;
  const T det = sqrt(jx * jx + jy * jy);  // element_line.py:63-69
  const T inv = T(1) / det;
  if (inv_jac) inv_jac[e] = inv;
  if (v_grad) {
    v_grad[2 * (int64_t)e] = T(-0.5) * inv;
    v_grad[2 * (int64_t)e + 1] = T(0.5) * inv;
  }
  T detf = T(1);
  if constexpr (FRAC) detf = __ldg(frac_det + mesh);
  // reference nodes (element_line.py:21-43): order 2 -> (-n, n); order 3 -> (0, -n, n)
  const T nodes[3] = {n_q == 2 ? -node : T(0), n_q == 2 ? node : -node, node};
  const T w[3] = {w0, w1, w2};
  for (int q = 0; q < n_q; ++q) {
    const T l0 = T(0.5) * (T(1) - nodes[q]);
    const T l1 = T(0.5) * (T(1) + nodes[q]);
    const T px = l0 * ax + l1 * bx;
    const T py = l0 * ay + l1 * by;
    if (x_q) {
      if constexpr (FRAC) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
          x_q[((int64_t)e * n_q + q) * 3 + r] = __ldg(frac_jac + 6 * mesh + 2 * r) * px +
                                                  __ldg(frac_jac + 6 * mesh + 2 * r + 1) * py +
                                                  __ldg(frac_t + 3 * mesh + r);
      } else {
        x_q[((int64_t)e * n_q + q) * 2] = px;
        x_q[((int64_t)e * n_q + q) * 2 + 1] = py;
      }
    }
    if (dx) dx[(int64_t)e * n_q + q] = T(2) * w[q] * det * detf;
  }
}

template <typename T>
int tri_geometry(int64_t n_el, int64_t n_el_per_mesh, int64_t n_vert_per_mesh, const T* coords,
                 const int32_t* conn, int quad_order, const T* frac_jac, const T* frac_inv,
                 const T* frac_det, const T* frac_t, T* inv_jac, T* v_grad, T* x_q, T* dx,
                 void* stream) {
  if (n_el < 0 || n_el_per_mesh <= 0 || n_vert_per_mesh <= 0) return TFEM_ERR_BAD_ARG;
  if (n_el == 0) return TFEM_OK;
  if (!coords || !conn) return TFEM_ERR_BAD_ARG;
  if (n_el > kMaxIndex / 9) return TFEM_ERR_TOO_LARGE;
  if (tri_n_q(quad_order) == 0) return TFEM_ERR_UNSUPPORTED;
  const bool frac = frac_jac != nullptr;
  if (frac && (!frac_inv || !frac_det || !frac_t)) return TFEM_ERR_BAD_ARG;
  const QuadT<T> quad = make_quad<T>(quad_order);
  const FracPtrs<T> fp{frac_jac, frac_inv, frac_det, frac_t};
  const int threads = 256;
  auto s = static_cast<cudaStream_t>(stream);
  if (frac)
    tri_geometry_kernel<T, true><<<blocks_for(n_el, threads), threads, 0, s>>>(
        (int)n_el, (int)n_el_per_mesh, (int)n_vert_per_mesh, coords, conn, quad, fp, inv_jac, v_grad, x_q, dx);
  else
    tri_geometry_kernel<T, false><<<blocks_for(n_el, threads), threads, 0, s>>>(
        (int)n_el, (int)n_el_per_mesh, (int)n_vert_per_mesh, coords, conn, quad, fp, inv_jac, v_grad, x_q, dx);
  return check_launch();
}

template <typename T>
int edge_geometry(int64_t n_edge, int64_t n_edge_per_mesh, const T* edge_coords, int quad_order,
                  const T* frac_jac, const T* frac_det, const T* frac_t, T* inv_jac, T* v_grad,
                  T* x_q, T* dx, void* stream) {
  if (n_edge < 0 || n_edge_per_mesh <= 0) return TFEM_ERR_BAD_ARG;
  if (n_edge == 0) return TFEM_OK;
  if (!edge_coords) return TFEM_ERR_BAD_ARG;
  if (n_edge > kMaxIndex / 9) return TFEM_ERR_TOO_LARGE;
  const int n_q = line_n_q(quad_order);
  if (n_q == 0) return TFEM_ERR_UNSUPPORTED;
  const bool frac = frac_jac != nullptr;
  if (frac && (!frac_det || !frac_t)) return TFEM_ERR_BAD_ARG;
  // element_line.py:21-43, evaluated in the working dtype like torch.sqrt(torch.tensor(.))
  T node, w0, w1, w2;
  if (n_q == 2) {
    node = T(1) / std::sqrt(T(3));
    w0 = w1 = T(0.5);
    w2 = T(0);
  } else {
    node = std::sqrt(T(3.0 / 5.0));
    w0 = T(8.0 / 18.0);
    w1 = w2 = T(5.0 / 18.0);
  }
  const int threads = 256;
  auto s = static_cast<cudaStream_t>(stream);
  if (frac)
    edge_geometry_kernel<T, true><<<blocks_for(n_edge, threads), threads, 0, s>>>(
        (int)n_edge, (int)n_edge_per_mesh, edge_coords, n_q, node, w0, w1, w2, frac_jac, frac_det, frac_t,
        inv_jac, v_grad, x_q, dx);
  else
    edge_geometry_kernel<T, false><<<blocks_for(n_edge, threads), threads, 0, s>>>(
        (int)n_edge, (int)n_edge_per_mesh, edge_coords, n_q, node, w0, w1, w2, frac_jac, frac_det, frac_t,
        inv_jac, v_grad, x_q, dx);
  return check_launch();
}

}  // namespace tfem

#define TFEM_GEOMETRY_API(T, SUF)                                                                   \
  extern "C" int tfem_tri_p1_geometry_##SUF(                                                        \
      int64_t n_el, int64_t n_el_per_mesh, int64_t n_vert_per_mesh, const T* coords,                \
      const int32_t* conn, int quad_order, const T* frac_jac, const T* frac_inv, const T* frac_det, \
      const T* frac_t, T* inv_jac, T* v_grad, T* x_q, T* dx, void* stream) {                        \
    return tfem::tri_geometry<T>(n_el, n_el_per_mesh, n_vert_per_mesh, coords, conn, quad_order,    \
                                 frac_jac, frac_inv, frac_det, frac_t, inv_jac, v_grad, x_q, dx,    \
                                 stream);                                                           \
  }                                                                                                 \
  extern "C" int tfem_edge_p1_geometry_##SUF(                                                       \
      int64_t n_edge, int64_t n_edge_per_mesh, const T* edge_coords, int quad_order,                \
      const T* frac_jac, const T* frac_det, const T* frac_t, T* inv_jac, T* v_grad, T* x_q, T* dx,  \
      void* stream) {                                                                               \
    return tfem::edge_geometry<T>(n_edge, n_edge_per_mesh, edge_coords, quad_order, frac_jac,       \
                                  frac_det, frac_t, inv_jac, v_grad, x_q, dx, stream);              \
  }

TFEM_GEOMETRY_API(double, f64)
TFEM_GEOMETRY_API(float, f32)
