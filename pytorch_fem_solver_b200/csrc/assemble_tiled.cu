// Fused P1 assembly to CSR + load vector, one CTA per row tile (BASELINE.json config 2).
//
// Replaces, in one pass and without materialising any per-element tensor in HBM, the reference
// pipeline  Basis.__init__ geometry (basis/abstract_basis.py:42-63)  ->  user form evaluation
// (tests/test_assembly.py:68-84)  ->  (integrand*dx).sum(-3) (abstract_basis.py:83,104)  ->
// index_put_(accumulate=True) (abstract_basis.py:87-91,106-110).
//
// A tile owns a set of CSR rows.  Its CTA
//   A. stages the coordinates of every vertex the tile touches in shared memory,
//   B. computes each tile element's local matrix (6 unique entries, the form is symmetric) and
//      load (3 entries) ONCE, into shared memory (elements on a tile border are recomputed by the
//      neighbouring tile: a halo of ~13% for 16x16 vertex blocks),
//   C. lets one thread per owned row sum its incident-element contributions in increasing
//      element order into a shared image of the row's CSR entries,
//   D. streams the image to csr_val in runs of consecutive rows (coalesced 8 B stores).
// HBM traffic per element is therefore ~coords + index plan + outputs; nothing is re-read.
#include "common.cuh"

namespace tfem {

template <typename T>
struct PlanDev {
  const int32_t* tile_ptr;
  const int32_t* tile_vert;
  const uint32_t* tile_elem;
  const int32_t* row_id;
  const int32_t* row_meta;
  const int32_t* row_corner_ptr;
  const uint32_t* corner;
  const int32_t* run_start;
  const int32_t* run_meta;
  int max_vert, max_elem, max_out;
};

template <typename T, int THREADS, bool HAS_MAT, bool HAS_LOAD>
__global__ void __launch_bounds__(THREADS) assemble_tiled_kernel(const PlanDev<T> plan,
                                                                 const T* __restrict__ coords,
                                                                 const QuadT<T> quad, const T alpha,
                                                                 const T beta, const SourceT<T> src,
                                                                 T* __restrict__ csr_val,
                                                                 T* __restrict__ load) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sx = reinterpret_cast<T*>(smem_raw);
  T* sy = sx + plan.max_vert;
  T* sloc = sy + plan.max_vert;           // [9][max_elem]: K00 K11 K22 K01 K12 K20 b0 b1 b2
  T* sout = sloc + 9 * plan.max_elem;     // [max_out] image of the tile's CSR entries
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int4 p0 = __ldg(reinterpret_cast<const int4*>(plan.tile_ptr) + tile);
  const int4 p1 = __ldg(reinterpret_cast<const int4*>(plan.tile_ptr) + tile + 1);
  const int n_vert = p1.x - p0.x, n_elem = p1.y - p0.y, n_rows = p1.z - p0.z, n_runs = p1.w - p0.w;
  const int max_elem = plan.max_elem;

  // ---- A: coordinates of the tile's vertices -> shared --------------------------------------
  for (int i = tid; i < n_vert; i += THREADS) {
    const int v = __ldg(plan.tile_vert + p0.x + i);
    T x, y;
    load_xy(coords, v, x, y);
    sx[i] = x;
    sy[i] = y;
  }
  if (HAS_MAT) {
    // runs tile the image in order, so the last run ends it
    const int last = n_runs > 0 ? __ldg(plan.run_meta + p1.w - 1) : 0;
    const int n_out = (last & 0xffff) + ((last >> 16) & 0xffff);
    for (int i = tid; i < n_out; i += THREADS) sout[i] = T(0);
  }
  __syncthreads();

  // ---- B: local matrices and loads, one element per thread ----------------------------------
  for (int el = tid; el < n_elem; el += THREADS) {
    const uint32_t packed = __ldg(plan.tile_elem + p0.y + el);
    const int a = packed & 1023u, b = (packed >> 10) & 1023u, c = (packed >> 20) & 1023u;
    const T x0 = sx[a], y0 = sy[a], x1 = sx[b], y1 = sy[b], x2 = sx[c], y2 = sy[c];
    const TriGeom<T> g = tri_geom(x0, y0, x1, y1, x2, y2);
    if (HAS_MAT) {
      const T g0x = -g.i00 - g.i10, g0y = -g.i01 - g.i11;
      const T g1x = g.i00, g1y = g.i01, g2x = g.i10, g2y = g.i11;
      const T ka = alpha * (quad.wsum * g.det);
      const T mb = beta * g.det;
      sloc[0 * max_elem + el] = ka * (g0x * g0x + g0y * g0y) + mb * quad.mref[0];
      sloc[1 * max_elem + el] = ka * (g1x * g1x + g1y * g1y) + mb * quad.mref[4];
      sloc[2 * max_elem + el] = ka * (g2x * g2x + g2y * g2y) + mb * quad.mref[8];
      sloc[3 * max_elem + el] = ka * (g0x * g1x + g0y * g1y) + mb * quad.mref[1];
      sloc[4 * max_elem + el] = ka * (g1x * g2x + g1y * g2y) + mb * quad.mref[5];
      sloc[5 * max_elem + el] = ka * (g2x * g0x + g2y * g0y) + mb * quad.mref[6];
    }
    if (HAS_LOAD) {
      T b0 = T(0), b1 = T(0), b2 = T(0);
      for (int q = 0; q < quad.n_q; ++q) {
        const T px = quad.l0[q] * x0 + quad.l1[q] * x1 + quad.l2[q] * x2;
        const T py = quad.l0[q] * y0 + quad.l1[q] * y1 + quad.l2[q] * y2;
        const T wf = (quad.w[q] * g.det) * source_eval(src, px, py);
        b0 += wf * quad.l0[q];
        b1 += wf * quad.l1[q];
        b2 += wf * quad.l2[q];
      }
      sloc[6 * max_elem + el] = b0;
      sloc[7 * max_elem + el] = b1;
      sloc[8 * max_elem + el] = b2;
    }
  }
  __syncthreads();

  // ---- C: one thread per owned row gathers its corners in increasing element order ----------
  for (int j = tid; j < n_rows; j += THREADS) {
    const int meta = __ldg(plan.row_meta + p0.z + j);
    const int base = meta & 0xffff, pos_diag = (meta >> 16) & 0xff;
    const int c0 = __ldg(plan.row_corner_ptr + p0.z + j);
    const int c1 = __ldg(plan.row_corner_ptr + p0.z + j + 1);
    T diag = T(0), rhs = T(0);
    for (int c = c0; c < c1; ++c) {
      const uint32_t cw = __ldg(plan.corner + c);
      const int el = cw & 0xfffu, k = (cw >> 12) & 3u;
      if (HAS_MAT) {
        const int pa = (cw >> 16) & 0xffu, pb = cw >> 24;
        const int kb = k == 0 ? 2 : k - 1;  // (k+2) % 3
        diag += sloc[k * max_elem + el];
        sout[base + pa] += sloc[(3 + k) * max_elem + el];   // entry (k+1, k)
        sout[base + pb] += sloc[(3 + kb) * max_elem + el];  // entry (k+2, k)
      }
      if (HAS_LOAD) rhs += sloc[(6 + k) * max_elem + el];
    }
    if (HAS_MAT) sout[base + pos_diag] += diag;
    if (HAS_LOAD) load[__ldg(plan.row_id + p0.z + j)] = rhs;
  }

  // ---- D: stream the row images out, one warp per run of consecutive rows --------------------
  if (HAS_MAT) {
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < n_runs; r += THREADS / 32) {
      const int gstart = __ldg(plan.run_start + p0.w + r);
      const int meta = __ldg(plan.run_meta + p0.w + r);
      const int base = meta & 0xffff, len = (meta >> 16) & 0xffff;
      for (int i = lane; i < len; i += 32) csr_val[(int64_t)gstart + i] = sout[base + i];
    }
  }
}

template <typename T>
int assemble_tiled(const tfem_tile_plan* hp, const T* coords, int quad_order, const tfem_bilinear* form,
                   const tfem_source* source, T* csr_val, T* load, void* stream) {
  if (!hp || hp->n_tiles < 0) return TFEM_ERR_BAD_ARG;
  if (hp->n_tiles == 0) return TFEM_OK;
  if (!coords || (!csr_val && !load)) return TFEM_ERR_BAD_ARG;
  if (csr_val && !form) return TFEM_ERR_BAD_ARG;
  if (!hp->tile_ptr || !hp->tile_vert || !hp->tile_elem || !hp->row_id || !hp->row_meta ||
      !hp->row_corner_ptr || !hp->corner || !hp->run_start || !hp->run_meta)
    return TFEM_ERR_BAD_ARG;
  if (hp->max_vert > 1024 || hp->max_elem > 4096 || hp->max_out > 65535) return TFEM_ERR_TOO_LARGE;
  if (hp->n_tiles > kMaxIndex) return TFEM_ERR_TOO_LARGE;
  if (tri_n_q(quad_order) == 0) return TFEM_ERR_UNSUPPORTED;
  const SourceT<T> src = make_source<T>(source);
  if (load && (src.kind == TFEM_SRC_SAMPLED || src.kind < TFEM_SRC_NONE || src.kind > TFEM_SRC_SINSIN))
    return TFEM_ERR_BAD_ARG;  // sampled sources go through tfem_tri_p1_local_forms
  const QuadT<T> quad = make_quad<T>(quad_order);
  const PlanDev<T> plan{hp->tile_ptr, hp->tile_vert, hp->tile_elem, hp->row_id, hp->row_meta,
                        hp->row_corner_ptr, hp->corner, hp->run_start, hp->run_meta,
                        hp->max_vert, hp->max_elem, hp->max_out};
  const size_t smem = sizeof(T) * ((size_t)2 * hp->max_vert + (size_t)9 * hp->max_elem + (size_t)hp->max_out);
  if (smem > 227 * 1024) return TFEM_ERR_TOO_LARGE;
  const T alpha = form ? T(form->alpha) : T(0), beta = form ? T(form->beta) : T(0);
  constexpr int THREADS = 256;
  auto s = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)hp->n_tiles;
#define TFEM_LAUNCH_TILED(MAT, LOAD)                                                              \
  do {                                                                                            \
    auto kern = assemble_tiled_kernel<T, THREADS, MAT, LOAD>;                                     \
    if (smem > 48 * 1024 &&                                                                       \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=     \
            cudaSuccess)                                                                          \
      return TFEM_ERR_LAUNCH;                                                                     \
    kern<<<grid, THREADS, smem, s>>>(plan, coords, quad, alpha, beta, src, csr_val, load);        \
  } while (0)
  if (csr_val && load) TFEM_LAUNCH_TILED(true, true);
  else if (csr_val) TFEM_LAUNCH_TILED(true, false);
  else TFEM_LAUNCH_TILED(false, true);
#undef TFEM_LAUNCH_TILED
  return check_launch();
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
