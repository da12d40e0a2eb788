// Shared device helpers for the tfem_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "tfem_b200.h"

namespace tfem {

constexpr int kMaxQ = 6;  // largest triangle table of the reference (element_tri.py:99-126)

// ---------------------------------------------------------------------------------------------
// Quadrature tables (reference: element/element_tri.py:77-130, element/element_line.py:21-43).
// Kept as double literals; the f32 kernels narrow them exactly as torch.tensor() does.
// ---------------------------------------------------------------------------------------------
struct TriTable {
  int n_q;
  double xi[kMaxQ];
  double eta[kMaxQ];
  double w[kMaxQ];
};

__host__ __device__ inline TriTable tri_table(int order) {
  TriTable t{};
  switch (order) {
    case 1:
      t.n_q = 1;
      t.xi[0] = 1.0 / 3.0; t.eta[0] = 1.0 / 3.0; t.w[0] = 1.0;
      break;
    case 2:
      t.n_q = 3;
      t.xi[0] = 1.0 / 6.0; t.eta[0] = 1.0 / 6.0;
      t.xi[1] = 2.0 / 3.0; t.eta[1] = 1.0 / 6.0;
      t.xi[2] = 1.0 / 6.0; t.eta[2] = 2.0 / 3.0;
      t.w[0] = t.w[1] = t.w[2] = 1.0 / 3.0;
      break;
    case 3:
      t.n_q = 4;
      t.xi[0] = 1.0 / 3.0; t.eta[0] = 1.0 / 3.0;
      t.xi[1] = 0.6; t.eta[1] = 0.2;
      t.xi[2] = 0.2; t.eta[2] = 0.6;
      t.xi[3] = 0.2; t.eta[3] = 0.2;
      t.w[0] = -9.0 / 16.0;
      t.w[1] = t.w[2] = t.w[3] = 25.0 / 48.0;
      break;
    case 4:
      t.n_q = 6;
      t.xi[0] = 0.816847572980459; t.eta[0] = 0.091576213509771;
      t.xi[1] = 0.091576213509771; t.eta[1] = 0.816847572980459;
      t.xi[2] = 0.091576213509771; t.eta[2] = 0.091576213509771;
      t.xi[3] = 0.108103018168070; t.eta[3] = 0.445948490915965;
      t.xi[4] = 0.445948490915965; t.eta[4] = 0.108103018168070;
      t.xi[5] = 0.445948490915965; t.eta[5] = 0.445948490915965;
      t.w[0] = t.w[1] = t.w[2] = 0.109951743655322;
      t.w[3] = t.w[4] = t.w[5] = 0.223381589678011;
      break;
    default:
      t.n_q = 0;
  }
  return t;
}

inline int tri_n_q(int order) { return tri_table(order).n_q; }
inline int line_n_q(int order) { return order == 2 ? 2 : (order == 3 ? 3 : 0); }

// Per-kernel quadrature data in the working precision, passed by value as a kernel parameter
// (lands in the constant bank; no cudaMemcpyToSymbol, so calls stay re-entrant).
template <typename T>
struct QuadT {
  int n_q;
  T l0[kMaxQ], l1[kMaxQ], l2[kMaxQ];  // barycentric coordinates of each point
  T w[kMaxQ];                          // reference_element_area * weight
  T mref[9];                           // sum_q w[q] * l_i(q) * l_j(q)   (reference mass matrix / det)
  T wsum;                              // sum_q w[q]
};

template <typename T>
inline QuadT<T> make_quad(int order) {
  TriTable t = tri_table(order);
  QuadT<T> q{};
  q.n_q = t.n_q;
  T wsum = T(0);
  for (int k = 0; k < t.n_q; ++k) {
    T xi = T(t.xi[k]), eta = T(t.eta[k]);
    q.l0[k] = T(1.0) - xi - eta;  // element_tri.py:23-26, evaluated in the working dtype
    q.l1[k] = xi;
    q.l2[k] = eta;
    q.w[k] = T(0.5) * T(t.w[k]);  // basis.py:93-96
    wsum += q.w[k];
  }
  q.wsum = wsum;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      T acc = T(0);
      for (int k = 0; k < t.n_q; ++k) {
        const T li = i == 0 ? q.l0[k] : (i == 1 ? q.l1[k] : q.l2[k]);
        const T lj = j == 0 ? q.l0[k] : (j == 1 ? q.l1[k] : q.l2[k]);
        acc += q.w[k] * (li * lj);
      }
      q.mref[3 * i + j] = acc;
    }
  return q;
}

template <typename T>
struct SourceT {
  int kind;
  T p0, p1, p2, p3;
};

template <typename T>
inline SourceT<T> make_source(const tfem_source* s) {
  SourceT<T> out{};
  if (s == nullptr) {
    out.kind = TFEM_SRC_NONE;
    return out;
  }
  out.kind = s->kind;
  out.p0 = T(s->p[0]);
  out.p1 = T(s->p[1]);
  out.p2 = T(s->p[2]);
  out.p3 = T(s->p[3]);
  return out;
}

template <typename T>
__device__ __forceinline__ T source_eval(const SourceT<T>& s, T x, T y) {
  if (s.kind == TFEM_SRC_SINSIN) return s.p0 * sin(s.p1 * x) * sin(s.p2 * y);
  if (s.kind == TFEM_SRC_CONST) return s.p0;
  return T(0);
}

// 2-wide vector types for coordinate loads (one 16 B / 8 B request per vertex)
template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

template <typename T>
__device__ __forceinline__ void load_xy(const T* __restrict__ coords, int64_t v, T& x, T& y) {
  using V = typename Vec2<T>::type;
  const V p = __ldg(reinterpret_cast<const V*>(coords) + v);
  x = p.x;
  y = p.y;
}

// Planar P1 triangle map (basis.py:87-88, element_tri.py:132-145): signed det, no abs().
template <typename T>
struct TriGeom {
  T x0, y0, x1, y1, x2, y2;
  T det;
  T i00, i01, i10, i11;  // J^-1
};

template <typename T>
__device__ __forceinline__ TriGeom<T> tri_geom(T x0, T y0, T x1, T y1, T x2, T y2) {
  TriGeom<T> g;
  g.x0 = x0; g.y0 = y0; g.x1 = x1; g.y1 = y1; g.x2 = x2; g.y2 = y2;
  const T a = x1 - x0, b = x2 - x0;  // J = [[a, b], [c, d]]
  const T c = y1 - y0, d = y2 - y0;
  g.det = a * d - b * c;
  const T r = T(1) / g.det;
  g.i00 = r * d;
  g.i01 = r * (-b);
  g.i10 = r * (-c);
  g.i11 = r * a;
  return g;
}

inline int check_launch() { return cudaGetLastError() == cudaSuccess ? TFEM_OK : TFEM_ERR_LAUNCH; }

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

constexpr int64_t kMaxIndex = 2147483647LL;

}  // namespace tfem
