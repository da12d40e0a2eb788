// Fused P1 assembly to CSR + load vector, one CTA per row tile (BASELINE.json config 2).
//
// Replaces, in one pass and without materialising any per-element tensor in HBM, the reference
// pipeline  Basis.__init__ geometry (basis/abstract_basis.py:42-63)  ->  user form evaluation
// (tests/test_assembly.py:68-84)  ->  (integrand*dx).sum(-3) (abstract_basis.py:83,104)  ->
// index_put_(accumulate=True) (abstract_basis.py:87-91,106-110).
//
// A tile owns a set of CSR rows; all of its index data sits in ONE contiguous, 16 B aligned blob
// that a single elected thread pulls into shared memory with a TMA bulk copy
// (cp.async.bulk ... mbarrier::complete_tx) while the other threads wait on the mbarrier.  Then
//   A. every tile vertex is gathered once: coordinates (one 16 B load) and, for the sin*sin
//      source, sin/cos of (w x, w y) -> shared;
//   B. every tile element is integrated ONCE: 6 unique matrix entries (the form is symmetric)
//      and 3 load entries -> shared.  f at the quadrature points comes from the vertex-0
//      sin/cos and a short Taylor rotation by the (tiny) in-element phase, so no fp64 sin() runs
//      per quadrature point.  Elements on a tile border are recomputed by the neighbouring tile
//      (halo ~14-20%);
//   C. one thread per owned row sums its incident-element contributions in increasing element
//      order into a shared image of the row's CSR entries (no atomics: bitwise reproducible);
//   D. the image is streamed to csr_val in runs of consecutive rows (coalesced 8 B stores).
// HBM traffic is therefore coords + index blob + outputs, each touched once.
#include "common.cuh"

namespace tfem {

// ---- mbarrier / TMA bulk-copy wrappers (PTX ISA 8.x, sm_90+) --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }
constexpr int kBlobHeader = 8;

template <int ORDER> struct NQ;
template <> struct NQ<1> { static constexpr int value = 1; };
template <> struct NQ<2> { static constexpr int value = 3; };
template <> struct NQ<3> { static constexpr int value = 4; };
template <> struct NQ<4> { static constexpr int value = 6; };

__device__ __forceinline__ void sincos_full(double x, double& s, double& c) { sincos(x, &s, &c); }
__device__ __forceinline__ void sincos_full(float x, float& s, float& c) { sincosf(x, &s, &c); }

// |phase| below which the 7th/6th-order Taylor rotation is exact to < 1e-18 relative
template <typename T> __device__ __forceinline__ T small_phase_limit() { return T(0.02); }

// sin(t), cos(t) for |t| <= 0.02
template <typename T>
__device__ __forceinline__ void sincos_small(T t, T& s, T& c) {
  const T z = t * t;
  T ps = fma(z, T(-1.0 / 5040.0), T(1.0 / 120.0));
  ps = fma(z, ps, T(-1.0 / 6.0));
  s = fma(t * z, ps, t);
  T pc = fma(z, T(-1.0 / 720.0), T(1.0 / 24.0));
  pc = fma(z, pc, T(-0.5));
  c = fma(z, pc, T(1));
}

template <typename T, int THREADS, int ORDER, int SRC, bool HAS_MAT>
__global__ void __launch_bounds__(THREADS) assemble_tiled_kernel(
    const int32_t* __restrict__ tile_off, const int32_t* __restrict__ blob, const int max_vert,
    const int elem_stride, const int max_blob_words, const T* __restrict__ coords, const QuadT<T> quad,
    const T alpha, const T beta, const SourceT<T> src, T* __restrict__ csr_val, T* __restrict__ load) {
  constexpr int NQV = NQ<ORDER>::value;
  constexpr bool HAS_LOAD = SRC != TFEM_SRC_NONE;
  constexpr bool SINSIN = SRC == TFEM_SRC_SINSIN;
  constexpr int VFIELDS = SINSIN ? 6 : 2;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [ mbarrier (16 B) | blob (16 B aligned, whole 16 B units) | vertex fields | sloc[9][elem_stride] | sout ]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  int32_t* sblob = reinterpret_cast<int32_t*>(smem_raw + 16);
  T* vtx = reinterpret_cast<T*>(sblob + max_blob_words);
  T* sloc = vtx + VFIELDS * max_vert;
  T* sout = sloc + 9 * elem_stride;

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int off0 = __ldg(tile_off + tile);
  const int words = __ldg(tile_off + tile + 1) - off0;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    mbar_expect_tx(bar, (uint32_t)words * 4u);
    bulk_g2s(sblob, blob + off0, (uint32_t)words * 4u, bar);
  }
  __syncthreads();  // barrier object initialised and visible to every waiter
  mbar_wait(bar, 0);

  const int n_vert = sblob[0], n_elem = sblob[1], n_rows = sblob[2], n_runs = sblob[3], n_corner = sblob[4],
            n_out = sblob[5];
  const int32_t* s_vert = sblob + kBlobHeader;
  const uint32_t* s_elem = reinterpret_cast<const uint32_t*>(s_vert + pad4(n_vert));
  const int32_t* s_row_id = reinterpret_cast<const int32_t*>(s_elem + pad4(n_elem));
  const int32_t* s_row_meta = s_row_id + pad4(n_rows);
  const int32_t* s_row_cptr = s_row_meta + pad4(n_rows);
  const uint32_t* s_corner = reinterpret_cast<const uint32_t*>(s_row_cptr + pad4(n_rows + 1));
  const int32_t* s_run_start = reinterpret_cast<const int32_t*>(s_corner + pad4(n_corner));
  const int32_t* s_run_meta = s_run_start + pad4(n_runs);

  // ---- A: tile vertices -> shared -------------------------------------------------------------
  T* vx = vtx;
  T* vy = vtx + max_vert;
  for (int i = tid; i < n_vert; i += THREADS) {
    T x, y;
    load_xy(coords, s_vert[i], x, y);
    vx[i] = x;
    vy[i] = y;
    if constexpr (SINSIN) {
      T s, c;
      sincos_full(src.p1 * x, s, c);
      vtx[2 * max_vert + i] = s;
      vtx[3 * max_vert + i] = c;
      sincos_full(src.p2 * y, s, c);
      vtx[4 * max_vert + i] = s;
      vtx[5 * max_vert + i] = c;
    }
  }
  if constexpr (HAS_MAT) {
    for (int i = tid; i < n_out; i += THREADS) sout[i] = T(0);
  }
  __syncthreads();

  // ---- B: local matrices and loads, each tile element once ------------------------------------
  for (int el = tid; el < n_elem; el += THREADS) {
    const uint32_t packed = s_elem[el];
    const int a = packed & 1023u, b = (packed >> 10) & 1023u, c = (packed >> 20) & 1023u;
    const T x0 = vx[a], y0 = vy[a];
    const T ax = vx[b] - x0, ay = vy[b] - y0;  // J = [[ax, bx], [ay, by]] (basis.py:87-88)
    const T bx = vx[c] - x0, by = vy[c] - y0;
    const T det = ax * by - bx * ay;  // signed (element_tri.py:139)
    if constexpr (HAS_MAT) {
      const T r = T(1) / det;
      const T g1x = r * by, g1y = -(r * bx);  // rows of J^-1 = grad(phi_1), grad(phi_2)
      const T g2x = -(r * ay), g2y = r * ax;
      const T g0x = -g1x - g2x, g0y = -g1y - g2y;
      const T ka = alpha * (quad.wsum * det);
      const T mb = beta * det;
      sloc[0 * elem_stride + el] = fma(ka, g0x * g0x + g0y * g0y, mb * quad.mref[0]);
      sloc[1 * elem_stride + el] = fma(ka, g1x * g1x + g1y * g1y, mb * quad.mref[4]);
      sloc[2 * elem_stride + el] = fma(ka, g2x * g2x + g2y * g2y, mb * quad.mref[8]);
      sloc[3 * elem_stride + el] = fma(ka, g0x * g1x + g0y * g1y, mb * quad.mref[1]);
      sloc[4 * elem_stride + el] = fma(ka, g1x * g2x + g1y * g2y, mb * quad.mref[5]);
      sloc[5 * elem_stride + el] = fma(ka, g2x * g0x + g2y * g0y, mb * quad.mref[6]);
    }
    if constexpr (HAS_LOAD) {
      T b0 = T(0), b1 = T(0), b2 = T(0);
      if constexpr (SINSIN) {
        // phase of the source relative to vertex 0: w*(x_q - x0) = xi*(w ax) + eta*(w bx)
        const T uax = src.p1 * ax, ubx = src.p1 * bx, uay = src.p2 * ay, uby = src.p2 * by;
        const T lim = small_phase_limit<T>();
        const bool small = fabs(uax) < lim && fabs(ubx) < lim && fabs(uay) < lim && fabs(uby) < lim;
        const T sx0 = vtx[2 * max_vert + a], cx0 = vtx[3 * max_vert + a];
        const T sy0 = vtx[4 * max_vert + a], cy0 = vtx[5 * max_vert + a];
        const T amp = src.p0 * det;
#pragma unroll
        for (int q = 0; q < NQV; ++q) {
          T sx, sy;
          if (small) {
            T s, cth;
            sincos_small(fma(quad.l1[q], uax, quad.l2[q] * ubx), s, cth);
            sx = fma(sx0, cth, cx0 * s);  // sin(w x0 + theta)
            sincos_small(fma(quad.l1[q], uay, quad.l2[q] * uby), s, cth);
            sy = fma(sy0, cth, cy0 * s);
          } else {  // coarse element: evaluate the source directly
            sx = sin(src.p1 * fma(quad.l1[q], ax, fma(quad.l2[q], bx, x0)));
            sy = sin(src.p2 * fma(quad.l1[q], ay, fma(quad.l2[q], by, y0)));
          }
          const T wf = (quad.w[q] * amp) * (sx * sy);
          b0 = fma(wf, quad.l0[q], b0);
          b1 = fma(wf, quad.l1[q], b1);
          b2 = fma(wf, quad.l2[q], b2);
        }
      } else {  // constant source: sum_q w_q l_k(q) is a per-order constant
        const T wf = src.p0 * det;
#pragma unroll
        for (int q = 0; q < NQV; ++q) {
          b0 = fma(wf * quad.w[q], quad.l0[q], b0);
          b1 = fma(wf * quad.w[q], quad.l1[q], b1);
          b2 = fma(wf * quad.w[q], quad.l2[q], b2);
        }
      }
      sloc[6 * elem_stride + el] = b0;
      sloc[7 * elem_stride + el] = b1;
      sloc[8 * elem_stride + el] = b2;
    }
  }
  __syncthreads();

  // ---- C: one thread per owned row gathers its corners in increasing element order ----------
  for (int j = tid; j < n_rows; j += THREADS) {
    const int meta = s_row_meta[j];
    const int base = meta & 0xffff, pos_diag = (meta >> 16) & 0xff;
    const int c0 = s_row_cptr[j], c1 = s_row_cptr[j + 1];
    T diag = T(0), rhs = T(0);
    for (int cidx = c0; cidx < c1; ++cidx) {
      const uint32_t cw = s_corner[cidx];
      const int el = cw & 0xfffu, k = (cw >> 12) & 3u;
      if constexpr (HAS_MAT) {
        const int pa = (cw >> 16) & 0xffu, pb = cw >> 24;
        const int kb = k == 0 ? 2 : k - 1;  // (k+2) % 3
        diag += sloc[k * elem_stride + el];
        sout[base + pa] += sloc[(3 + k) * elem_stride + el];   // entry (k+1, k)
        sout[base + pb] += sloc[(3 + kb) * elem_stride + el];  // entry (k+2, k)
      }
      if constexpr (HAS_LOAD) rhs += sloc[(6 + k) * elem_stride + el];
    }
    if constexpr (HAS_MAT) {
      if (c1 > c0) sout[base + pos_diag] += diag;
    }
    if constexpr (HAS_LOAD) load[s_row_id[j]] = rhs;
  }

  // ---- D: stream the row images out, one warp per run of consecutive rows --------------------
  if constexpr (HAS_MAT) {
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < n_runs; r += THREADS / 32) {
      const int gstart = s_run_start[r];
      const int meta = s_run_meta[r];
      const int base = meta & 0xffff, len = (meta >> 16) & 0xffff;
      for (int i = lane; i < len; i += 32) csr_val[(int64_t)gstart + i] = sout[base + i];
    }
  }
}

template <typename T>
size_t tiled_smem_bytes(const tfem_tile_plan* hp, int src_kind, int* elem_stride) {
  const int vfields = src_kind == TFEM_SRC_SINSIN ? 6 : 2;
  *elem_stride = (hp->max_elem + 31) & ~31;
  return 16 + 4 * (size_t)((hp->max_blob_words + 3) & ~3) +
         sizeof(T) * ((size_t)vfields * hp->max_vert + (size_t)9 * *elem_stride + (size_t)hp->max_out);
}

template <typename T, int THREADS, int ORDER, int SRC, bool HAS_MAT>
int launch_tiled(const tfem_tile_plan* hp, const T* coords, const QuadT<T>& quad, T alpha, T beta,
                 const SourceT<T>& src, T* csr_val, T* load, cudaStream_t s) {
  int elem_stride = 0;
  const size_t smem = tiled_smem_bytes<T>(hp, SRC, &elem_stride);
  if (smem > 227 * 1024) return TFEM_ERR_TOO_LARGE;
  auto kern = assemble_tiled_kernel<T, THREADS, ORDER, SRC, HAS_MAT>;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return TFEM_ERR_LAUNCH;
  kern<<<(unsigned)hp->n_tiles, THREADS, smem, s>>>(hp->tile_off, hp->blob, hp->max_vert, elem_stride, (hp->max_blob_words + 3) & ~3,
                                                    coords, quad, alpha, beta, src, csr_val, load);
  return check_launch();
}

template <typename T, int THREADS, int ORDER>
int dispatch_tiled(const tfem_tile_plan* hp, const T* coords, const QuadT<T>& quad, T alpha, T beta,
                   const SourceT<T>& src, T* csr_val, T* load, cudaStream_t s) {
  const int kind = load ? src.kind : TFEM_SRC_NONE;
  if (csr_val) {
    if (kind == TFEM_SRC_SINSIN) return launch_tiled<T, THREADS, ORDER, TFEM_SRC_SINSIN, true>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
    if (kind == TFEM_SRC_CONST) return launch_tiled<T, THREADS, ORDER, TFEM_SRC_CONST, true>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
    return launch_tiled<T, THREADS, ORDER, TFEM_SRC_NONE, true>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
  }
  if (kind == TFEM_SRC_SINSIN) return launch_tiled<T, THREADS, ORDER, TFEM_SRC_SINSIN, false>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
  if (kind == TFEM_SRC_CONST) return launch_tiled<T, THREADS, ORDER, TFEM_SRC_CONST, false>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
  return TFEM_ERR_BAD_ARG;
}

template <typename T>
int assemble_tiled_impl_zero(const tfem_tile_plan* hp, const T* coords, int quad_order, const tfem_bilinear* form,
                             const SourceT<T>& src, T* csr_val, T* load, void* stream) {
  const QuadT<T> quad = make_quad<T>(quad_order);
  const T alpha = form ? T(form->alpha) : T(0), beta = form ? T(form->beta) : T(0);
  auto s = static_cast<cudaStream_t>(stream);
  constexpr int THREADS = 256;
  switch (quad_order) {
    case 1: return dispatch_tiled<T, THREADS, 1>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
    case 2: return dispatch_tiled<T, THREADS, 2>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
    case 3: return dispatch_tiled<T, THREADS, 3>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
    default: return dispatch_tiled<T, THREADS, 4>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
  }
}

template <typename T>
int assemble_tiled(const tfem_tile_plan* hp, const T* coords, int quad_order, const tfem_bilinear* form,
                   const tfem_source* source, T* csr_val, T* load, void* stream) {
  if (!hp || hp->n_tiles < 0) return TFEM_ERR_BAD_ARG;
  if (hp->n_tiles == 0) return TFEM_OK;
  if (!coords || (!csr_val && !load)) return TFEM_ERR_BAD_ARG;
  if (csr_val && !form) return TFEM_ERR_BAD_ARG;
  if (!hp->tile_off || !hp->blob) return TFEM_ERR_BAD_ARG;
  if (hp->max_vert > 1024 || hp->max_elem > 4096 || hp->max_out > 65535) return TFEM_ERR_TOO_LARGE;
  if (hp->n_tiles > kMaxIndex) return TFEM_ERR_TOO_LARGE;
  if (tri_n_q(quad_order) == 0) return TFEM_ERR_UNSUPPORTED;
  const SourceT<T> src = make_source<T>(source);
  if (load && (src.kind == TFEM_SRC_SAMPLED || src.kind < TFEM_SRC_NONE || src.kind > TFEM_SRC_SINSIN))
    return TFEM_ERR_BAD_ARG;  // sampled sources go through tfem_tri_p1_local_forms
  if (load && src.kind == TFEM_SRC_NONE) {
    // f == 0: the load vector is zero; still produced by the kernel so every row is written
    SourceT<T> zero = src;
    zero.kind = TFEM_SRC_CONST;
    zero.p0 = T(0);
    return assemble_tiled_impl_zero(hp, coords, quad_order, form, zero, csr_val, load, stream);
  }
  return assemble_tiled_impl_zero(hp, coords, quad_order, form, src, csr_val, load, stream);
}

}  // namespace tfem

extern "C" int tfem_tri_p1_assemble_csr_f64(const tfem_tile_plan* host_plan, const double* coords,
                                            int quad_order, const tfem_bilinear* host_form,
                                            const tfem_source* host_source, double* csr_val, double* load,
                                            void* stream) {
  return tfem::assemble_tiled<double>(host_plan, coords, quad_order, host_form, host_source, csr_val, load, stream);
}

extern "C" int tfem_tri_p1_assemble_csr_f32(const tfem_tile_plan* host_plan, const float* coords,
                                            int quad_order, const tfem_bilinear* host_form,
                                            const tfem_source* host_source, float* csr_val, float* load,
                                            void* stream) {
  return tfem::assemble_tiled<float>(host_plan, coords, quad_order, host_form, host_source, csr_val, load, stream);
}
