// Fused P1 assembly to CSR + load vector: persistent, warp-specialised CTAs over row tiles
// (BASELINE.json config 2).
//
// Replaces, in one pass and without materialising any per-element tensor in HBM, the reference
// pipeline  Basis.__init__ geometry (basis/abstract_basis.py:42-63)  ->  user form evaluation
// (tests/test_assembly.py:68-84)  ->  (integrand*dx).sum(-3) (abstract_basis.py:83,104)  ->
// index_put_(accumulate=True) (abstract_basis.py:87-91,106-110).
//
// A tile owns a set of CSR rows; all of its index data is ONE contiguous, 16 B aligned blob.
// Each CTA is resident for the whole launch and walks tiles  blockIdx.x, +gridDim.x, ...
//
//   producer warp (1 warp)                       consumer warps (8 warps)
//   ------------------------------------------   ------------------------------------------------
//   TMA bulk copies (cp.async.bulk + mbarrier     wait full[t]
//   complete_tx) of the E blob of tile t+2 and   A  rotate the tile's base sin/cos to every vertex
//   the L blob of tile t+1                       B  integrate every tile element ONCE: 6 matrix
//   wait E blob t+1                                 entries (symmetric form) + 3 load entries
//   cp.async gather of the vertex coordinates       -> shared
//   of tile t+1 (16 B per vertex) -> shared      C  one thread per CSR ENTRY of the tile sums the
//   full-range sin/cos at ONE base vertex           entry's contributions in increasing element
//   arrive full[t+1]                                 order (no atomics, no read-modify-write) and
//                                                   stores it: consecutive threads write
//                                                   consecutive csr_val slots (coalesced);
//                                                   one thread per owned row does the load entry
//                                                arrive done[t]
//
// so global-memory latency (blob, coordinate gather) and the only library sin/cos of a tile are
// off the consumers' critical path.  f at the quadrature points is obtained by rotating the
// vertex-0 sin/cos by the in-element phase with a short Taylor series: no fp64 sin() per point.
// Elements on a tile border are recomputed by the neighbouring tile (halo ~15%).
// HBM traffic is coords + index blob + outputs, each touched once.
#include "common.cuh"

// Profiling aid (never set in the shipped build): -DTFEM_DEBUG_SKIP=<mask> removes phases so their
// share of the kernel time can be measured; 1 = A (vertex sin/cos), 2 = B (integration), 4 = C,
// 8 = keep C's arithmetic but suppress its global stores.
#ifndef TFEM_DEBUG_SKIP
#define TFEM_DEBUG_SKIP 0
#endif
// tuning knobs of the shipped build (overridable for experiments)
#ifndef TFEM_L_STAGES
#define TFEM_L_STAGES 1  // one L buffer: refilled while the next tile integrates; leaves room for 3 CTAs/SM
#endif
#ifndef TFEM_MIN_CTAS
#define TFEM_MIN_CTAS 3  // resident CTAs per SM of the 256-consumer build (caps registers at 72)
#endif
#ifndef TFEM_MIN_CTAS_128
#define TFEM_MIN_CTAS_128 5
#endif

// -DTFEM_DEBUG_TIMING: consumer thread 0 / producer lane 0 of a few CTAs print the cycles they spent
// per phase (summed over the CTA's tiles).  Profiling aid only.
#ifdef TFEM_DEBUG_TIMING
#include <cstdio>
#define TFEM_T_DECL long long t_mark = clock64(), t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TFEM_T(i)                      \
  do {                                 \
    const long long t_now = clock64(); \
    t_acc[i] += t_now - t_mark;        \
    t_mark = t_now;                    \
  } while (0)
#else
#define TFEM_T_DECL
#define TFEM_T(i)
#endif

namespace tfem {

// ---- mbarrier / TMA bulk-copy / cp.async wrappers (PTX ISA 8.x, sm_90+) ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Bounded spin: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
template <int BYTES>
__device__ __forceinline__ void cp_async(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int THREADS>
__device__ __forceinline__ void consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}

__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }
constexpr int kBlobHeader = 12;
constexpr int kEStages = 3;       // E blobs in flight
constexpr int kLStages = TFEM_L_STAGES;  // L blobs in flight
constexpr int kSmemHeader = 256;  // mbarriers (9 x 8 B) + two base-point records (2 x 6 values)

template <int ORDER> struct NQ;
template <> struct NQ<1> { static constexpr int value = 1; };
template <> struct NQ<2> { static constexpr int value = 3; };
template <> struct NQ<3> { static constexpr int value = 4; };
template <> struct NQ<4> { static constexpr int value = 6; };

__device__ __forceinline__ void sincos_full(double x, double& s, double& c) { sincos(x, &s, &c); }
__device__ __forceinline__ void sincos_full(float x, float& s, float& c) { sincosf(x, &s, &c); }

// sin(t), cos(t) for |t| <= 0.02 (truncation < 1e-18 relative)
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

// sin(t), cos(t) for |t| <= 0.1 (truncation < 3e-18 relative)
template <typename T>
__device__ __forceinline__ void sincos_medium(T t, T& s, T& c) {
  const T z = t * t;
  T ps = fma(z, T(1.0 / 362880.0), T(-1.0 / 5040.0));
  ps = fma(z, ps, T(1.0 / 120.0));
  ps = fma(z, ps, T(-1.0 / 6.0));
  s = fma(t * z, ps, t);
  T pc = fma(z, T(-1.0 / 3628800.0), T(1.0 / 40320.0));
  pc = fma(z, pc, T(-1.0 / 720.0));
  pc = fma(z, pc, T(1.0 / 24.0));
  pc = fma(z, pc, T(-0.5));
  c = fma(z, pc, T(1));
}

// sin/cos(w*x) from sin/cos(w*xb): rotate by the phase difference when it is small
template <typename T>
__device__ __forceinline__ void sincos_about(T w, T x, T xb, T sb, T cb, T& s, T& c) {
  const T th = w * (x - xb);
  if (fabs(th) <= T(0.1)) {
    T st, ct;
    sincos_medium(th, st, ct);
    s = fma(sb, ct, cb * st);
    c = fma(cb, ct, -(sb * st));
  } else {
    sincos_full(w * x, s, c);
  }
}

struct EView {  // "early" blob: header + vertices + connectivity
  int n_vert, n_elem, n_rows, n_runs, n_out, n_contrib, base_vertex, n_lcontrib, n_heavy;
  const int32_t* vert;
  const uint32_t* elem;
};

__device__ __forceinline__ EView view_e(const int32_t* b) {
  EView v;
  v.n_vert = b[0]; v.n_elem = b[1]; v.n_rows = b[2]; v.n_runs = b[3];
  v.n_out = b[4]; v.n_contrib = b[5]; v.base_vertex = b[6]; v.n_lcontrib = b[7]; v.n_heavy = b[8];
  v.vert = b + kBlobHeader;
  v.elem = reinterpret_cast<const uint32_t*>(v.vert + pad4(v.n_vert));
  return v;
}

struct LView {  // "late" blob: rows, runs and per-entry contribution lists
  const int32_t* row_id;
  const int32_t* run_start;
  const int32_t* run_meta;
  const uint16_t* ent_seg;
  const uint16_t* contrib;
  const uint16_t* lrow_seg;
  const uint16_t* lcontrib;
  const uint32_t* row_diag;
  const uint16_t* heavy;
  const uint32_t* heavy_pos;
};

__device__ __forceinline__ LView view_l(const int32_t* b, const EView& e) {
  LView v;
  v.row_id = b;
  v.run_start = v.row_id + pad4(e.n_rows);
  v.run_meta = v.run_start + pad4(e.n_runs);
  const int32_t* p = v.run_meta + pad4(e.n_runs);
  v.ent_seg = reinterpret_cast<const uint16_t*>(p);
  p += pad4((e.n_out + 2) >> 1);
  v.contrib = reinterpret_cast<const uint16_t*>(p);
  p += pad4((e.n_contrib + 1) >> 1);
  v.lrow_seg = reinterpret_cast<const uint16_t*>(p);
  p += pad4((e.n_rows + 2) >> 1);
  v.lcontrib = reinterpret_cast<const uint16_t*>(p);
  p += pad4((e.n_lcontrib + 1) >> 1);
  v.row_diag = reinterpret_cast<const uint32_t*>(p);
  p += pad4(e.n_rows);
  v.heavy = reinterpret_cast<const uint16_t*>(p);
  p += pad4((e.n_heavy + 1) >> 1);
  v.heavy_pos = reinterpret_cast<const uint32_t*>(p);
  return v;
}

template <typename T, int CONSUMERS, int ORDER, int SRC, bool HAS_MAT>
__global__ void __launch_bounds__(CONSUMERS + 32, (CONSUMERS == 128 ? TFEM_MIN_CTAS_128 : (CONSUMERS == 256 ? TFEM_MIN_CTAS : 2))) assemble_tiled_kernel(
    const int n_tiles, const int32_t* __restrict__ tile_list, const int32_t* __restrict__ e_off, const int32_t* __restrict__ e_blob,
    const int32_t* __restrict__ l_off, const int32_t* __restrict__ l_blob, const int max_vert,
    const int elem_stride, const int e_words, const int l_words, const T* __restrict__ coords,
    const QuadT<T> quad, const T alpha, const T beta, const SourceT<T> src, T* __restrict__ csr_val,
    T* __restrict__ load) {
  constexpr int NQV = NQ<ORDER>::value;
  constexpr bool HAS_LOAD = SRC != TFEM_SRC_NONE;
  constexpr bool SINSIN = SRC == TFEM_SRC_SINSIN;
  using V2 = typename Vec2<T>::type;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [ header | E blob x3 | L blob x2 | vertex coordinates x2 | sin/cos fields | sloc[9][elem_stride] ]
  uint64_t* e_bar = reinterpret_cast<uint64_t*>(smem_raw);  // [3] E blob landed
  uint64_t* l_bar = e_bar + kEStages;                       // [2] L blob landed
  uint64_t* full_bar = l_bar + kLStages;                    // [2] tile staged (coords + base point)
  uint64_t* done_bar = full_bar + 2;                        // [2] consumers finished the tile
  T* sbase = reinterpret_cast<T*>(smem_raw + 128);          // [2][6] bx, by, sin/cos(w bx), sin/cos(w by)
  int32_t* s_e = reinterpret_cast<int32_t*>(smem_raw + kSmemHeader);
  int32_t* s_l = s_e + kEStages * e_words;
  V2* vxy = reinterpret_cast<V2*>(s_l + kLStages * l_words);  // [2][max_vert]
  T* trig = reinterpret_cast<T*>(vxy + 2 * max_vert);         // [4][max_vert]
  T* sloc = trig + (SINSIN ? 4 : 0) * max_vert;               // [9][elem_stride]

  const int tid = threadIdx.x;
  const int n_local = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0) {
    for (int i = 0; i < kEStages + kLStages + 4; ++i) mbar_init(e_bar + i, 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= CONSUMERS) {
    // =================================== producer warp =======================================
    const int lane = tid - CONSUMERS;
    auto issue_e = [&](int it) {
      const int slot_index = (int)blockIdx.x + it * (int)gridDim.x;
      const int tile = tile_list ? __ldg(tile_list + slot_index) : slot_index;
      const int off0 = __ldg(e_off + tile);
      const uint32_t bytes = (uint32_t)(__ldg(e_off + tile + 1) - off0) * 4u;
      const int slot = it % kEStages;
      mbar_expect_tx(e_bar + slot, bytes);
      bulk_g2s(s_e + slot * e_words, e_blob + off0, bytes, e_bar + slot);
    };
    auto issue_l = [&](int it) {
      const int slot_index = (int)blockIdx.x + it * (int)gridDim.x;
      const int tile = tile_list ? __ldg(tile_list + slot_index) : slot_index;
      const int off0 = __ldg(l_off + tile);
      const uint32_t bytes = (uint32_t)(__ldg(l_off + tile + 1) - off0) * 4u;
      const int slot = it % kLStages;
      mbar_expect_tx(l_bar + slot, bytes);
      bulk_g2s(s_l + slot * l_words, l_blob + off0, bytes, l_bar + slot);
    };
    if (lane == 0) {
      issue_e(0);
      issue_l(0);
      if (n_local > 1) issue_e(1);
    }
    TFEM_T_DECL;
    for (int it = 0; it < n_local; ++it) {
      const int slot = it % kEStages, buf = it & 1;
      // stage tile `it` while the consumers work on tile it-1: vxy[buf] / sbase[buf] were last
      // read by tile it-2
      if (it >= 2) mbar_wait(done_bar + (it & 1), ((it - 2) >> 1) & 1);
      TFEM_T(0);
      mbar_wait(e_bar + slot, (it / kEStages) & 1);
      TFEM_T(1);
      const EView ev = view_e(s_e + slot * e_words);
      V2* dst = vxy + buf * max_vert;
      for (int i = lane; i < ev.n_vert; i += 32)
        cp_async<(int)sizeof(V2)>(dst + i, reinterpret_cast<const V2*>(coords) + ev.vert[i]);
      if constexpr (SINSIN) {
        T bx, by, sx, cx, sy, cy;
        load_xy(coords, ev.base_vertex, bx, by);
        sincos_full(src.p1 * bx, sx, cx);
        sincos_full(src.p2 * by, sy, cy);
        if (lane == 0) {
          T* sb = sbase + 6 * buf;
          sb[0] = bx; sb[1] = by; sb[2] = sx; sb[3] = cx; sb[4] = sy; sb[5] = cy;
        }
      }
      TFEM_T(2);
      cp_async_wait_all();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar + buf);
      TFEM_T(3);
      if (kLStages == 1 && it >= 1) {  // single L buffer: refill it for tile `it` once tile it-1 is done
        mbar_wait(done_bar + ((it - 1) & 1), ((it - 1) >> 1) & 1);
        if (lane == 0) issue_l(it);
      }
      // E slot of tile it+2 and L slot of tile it+1 are the ones tile it-1 used
      if (it + 1 < n_local) {
        if (it >= 1) mbar_wait(done_bar + ((it - 1) & 1), ((it - 1) >> 1) & 1);
        if (lane == 0) {
          if (kLStages >= 2) issue_l(it + 1);
          if (it + 2 < n_local) issue_e(it + 2);
        }
      }
      TFEM_T(4);
    }
#ifdef TFEM_DEBUG_TIMING
    if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 151))
      printf("cta %d producer (%d tiles): wait done(it-2) %lld | wait E %lld | issue gather + base sincos %lld | gather landed %lld | wait done(it-1)+issue %lld\n",
             (int)blockIdx.x, n_local, t_acc[0], t_acc[1], t_acc[2], t_acc[3], t_acc[4]);
#endif
    return;
  }

  // ===================================== consumer warps ========================================
  TFEM_T_DECL;
  for (int it = 0; it < n_local; ++it) {
    const int slot = it % kEStages, buf = it & 1;
    mbar_wait(full_bar + buf, (it >> 1) & 1);
    TFEM_T(0);
    mbar_wait(e_bar + slot, (it / kEStages) & 1);  // TMA writes visible to this thread too
    const EView ev = view_e(s_e + slot * e_words);
    const V2* xy = vxy + buf * max_vert;

    // ---- A: sin/cos of the source phase at every tile vertex, rotated from the base vertex ----
    if constexpr (SINSIN) {
      const T* sb = sbase + 6 * buf;
      const T bx = sb[0], by = sb[1], sbx = sb[2], cbx = sb[3], sby = sb[4], cby = sb[5];
      for (int i = tid; i < ((TFEM_DEBUG_SKIP & 1) ? 0 : ev.n_vert); i += CONSUMERS) {
        const V2 p = xy[i];
        T s, c;
        sincos_about(src.p1, p.x, bx, sbx, cbx, s, c);
        trig[0 * max_vert + i] = s;
        trig[1 * max_vert + i] = c;
        sincos_about(src.p2, p.y, by, sby, cby, s, c);
        trig[2 * max_vert + i] = s;
        trig[3 * max_vert + i] = c;
      }
      consumer_sync<CONSUMERS>();
    }

    TFEM_T(1);
    // ---- B: local matrices and loads, each tile element once ----------------------------------
    for (int el = tid; el < ((TFEM_DEBUG_SKIP & 2) ? 0 : ev.n_elem); el += CONSUMERS) {
      const uint32_t packed = ev.elem[el];
      const int a = packed & 1023u, b = (packed >> 10) & 1023u, c = (packed >> 20) & 1023u;
      const V2 p0 = xy[a], p1 = xy[b], p2 = xy[c];
      const T x0 = p0.x, y0 = p0.y;
      const T ax = p1.x - x0, ay = p1.y - y0;  // J = [[ax, bx], [ay, by]] (basis.py:87-88)
      const T bx = p2.x - x0, by = p2.y - y0;
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
          const T lim = T(0.02);
          const bool small = fabs(uax) < lim && fabs(ubx) < lim && fabs(uay) < lim && fabs(uby) < lim;
          const T sx0 = trig[0 * max_vert + a], cx0 = trig[1 * max_vert + a];
          const T sy0 = trig[2 * max_vert + a], cy0 = trig[3 * max_vert + a];
          const T amp = src.p0 * det;
          if (small) {  // one straight-line block: the 2*NQ Taylor chains are independent (ILP)
#pragma unroll
            for (int q = 0; q < NQV; ++q) {
              T s, cth;
              sincos_small(fma(quad.l1[q], uax, quad.l2[q] * ubx), s, cth);
              const T sx = fma(sx0, cth, cx0 * s);  // sin(w x0 + theta)
              sincos_small(fma(quad.l1[q], uay, quad.l2[q] * uby), s, cth);
              const T sy = fma(sy0, cth, cy0 * s);
              const T wf = (quad.w[q] * amp) * (sx * sy);
              b0 = fma(wf, quad.l0[q], b0);
              b1 = fma(wf, quad.l1[q], b1);
              b2 = fma(wf, quad.l2[q], b2);
            }
          } else {  // coarse element: evaluate the source directly
            for (int q = 0; q < NQV; ++q) {
              const T sx = sin(src.p1 * fma(quad.l1[q], ax, fma(quad.l2[q], bx, x0)));
              const T sy = sin(src.p2 * fma(quad.l1[q], ay, fma(quad.l2[q], by, y0)));
              const T wf = (quad.w[q] * amp) * (sx * sy);
              b0 = fma(wf, quad.l0[q], b0);
              b1 = fma(wf, quad.l1[q], b1);
              b2 = fma(wf, quad.l2[q], b2);
            }
          }
        } else {  // constant source
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
    consumer_sync<CONSUMERS>();

    TFEM_T(2);
    // ---- C: one thread per CSR entry / per load entry; contributions in increasing element id --
    mbar_wait(l_bar + (it % kLStages), (it / kLStages) & 1);
    TFEM_T(3);
    const LView lv = view_l(s_l + (it % kLStages) * l_words, ev);
    auto store = [&](int64_t pos, T value) {
      if (!(TFEM_DEBUG_SKIP & 8) || value == T(1.2345e30)) csr_val[pos] = value;
    };
    if constexpr (HAS_MAT) {
      // Entries with <= 2 contributions (every off-diagonal of a manifold mesh): one warp per run of
      // consecutive rows, lane i takes entries i, i+32, ... of the run, so a warp store covers 32
      // consecutive csr_val slots; two entries per lane are kept in flight.
      const int lane = tid & 31, warp = tid >> 5;
      for (int r = warp; r < ((TFEM_DEBUG_SKIP & 4) ? 0 : ev.n_runs); r += CONSUMERS / 32) {
        const int meta = lv.run_meta[r];
        const int base = meta & 0xffff, len = (meta >> 16) & 0xffff;
        const int64_t gpos = lv.run_start[r];
        for (int i = lane; i < len; i += 64) {
          const int i1 = i + 32;
          const bool two = i1 < len;
          const int oa = base + i, ob = base + (two ? i1 : i);
          const int sa = lv.ent_seg[oa], na = lv.ent_seg[oa + 1] - sa;
          const int sb = lv.ent_seg[ob], nb = lv.ent_seg[ob + 1] - sb;
          const int ca0 = lv.contrib[na > 0 ? sa : 0], ca1 = lv.contrib[na > 1 ? sa + 1 : 0];
          const int cb0 = lv.contrib[nb > 0 ? sb : 0], cb1 = lv.contrib[nb > 1 ? sb + 1 : 0];
          const T va0 = sloc[ca0], va1 = sloc[ca1], vb0 = sloc[cb0], vb1 = sloc[cb1];
          T acc_a = na > 0 ? va0 : T(0), acc_b = nb > 0 ? vb0 : T(0);
          if (na > 1) acc_a += va1;
          if (nb > 1) acc_b += vb1;
          if (na <= 2) store(gpos + i, acc_a);
          if (two && nb <= 2) store(gpos + i1, acc_b);
        }
      }
      TFEM_T(6);
      // the few other entries with > 2 contributions (non-manifold edges, degenerate elements)
      for (int h = tid; h < ((TFEM_DEBUG_SKIP & 4) ? 0 : ev.n_heavy); h += CONSUMERS) {
        const int o = lv.heavy[h];
        T acc = T(0);
        for (int s = lv.ent_seg[o]; s < lv.ent_seg[o + 1]; ++s) acc += sloc[lv.contrib[s]];
        store(lv.heavy_pos[h], acc);
      }
    }
    TFEM_T(7);
    // one thread per owned row: its load entry and its diagonal share one element list
    for (int j = tid; j < ((TFEM_DEBUG_SKIP & 4) ? 0 : ev.n_rows); j += CONSUMERS) {
      const int s0 = lv.lrow_seg[j], s1 = lv.lrow_seg[j + 1];
      T rhs = T(0), diag = T(0);
      for (int s = s0; s < s1; ++s) {
        const int code = lv.lcontrib[s];  // k*elem_stride + element
        if constexpr (HAS_LOAD) rhs += sloc[code + 6 * elem_stride];
        if constexpr (HAS_MAT) diag += sloc[code];
      }
      if constexpr (HAS_LOAD) {
        if (!(TFEM_DEBUG_SKIP & 8) || rhs == T(1.2345e30)) load[lv.row_id[j]] = rhs;
      }
      if constexpr (HAS_MAT) {
        const uint32_t pos = lv.row_diag[j];
        if (pos != 0xffffffffu) store(pos, diag);
      }
    }
    TFEM_T(4);
    consumer_sync<CONSUMERS>();  // sloc / trig / blob slots free for the next tile
    if (tid == 0) mbar_arrive(done_bar + buf);
    TFEM_T(5);
  }
#ifdef TFEM_DEBUG_TIMING
  if ((tid == 0 || tid == 255) && (blockIdx.x == 0 || blockIdx.x == 151))
    printf("cta %d consumer t%d (%d tiles): wait full %lld | A+sync %lld | B+sync %lld | wait L %lld | C light %lld heavy %lld rows %lld | end sync %lld\n",
           (int)blockIdx.x, tid, n_local, t_acc[0], t_acc[1], t_acc[2], t_acc[3], t_acc[6], t_acc[7], t_acc[4], t_acc[5]);
#endif
}

template <typename T>
size_t tiled_smem_bytes(const tfem_tile_plan* hp, int src_kind, int* elem_stride, int* e_words, int* l_words) {
  const int trig_fields = src_kind == TFEM_SRC_SINSIN ? 4 : 0;
  *elem_stride = hp->elem_stride;
  *e_words = (hp->max_e_words + 3) & ~3;
  *l_words = (hp->max_l_words + 3) & ~3;
  return kSmemHeader + 4 * ((size_t)kEStages * *e_words + (size_t)kLStages * *l_words) +
         sizeof(T) * ((size_t)(4 + trig_fields) * hp->max_vert + (size_t)9 * *elem_stride);
}

template <typename T, int CONSUMERS, int ORDER, int SRC, bool HAS_MAT>
int launch_tiled(const tfem_tile_plan* hp, const T* coords, const QuadT<T>& quad, T alpha, T beta,
                 const SourceT<T>& src, T* csr_val, T* load, cudaStream_t s) {
  int elem_stride = 0, e_words = 0, l_words = 0;
  const size_t smem = tiled_smem_bytes<T>(hp, SRC, &elem_stride, &e_words, &l_words);
  if (smem > 227 * 1024) return TFEM_ERR_TOO_LARGE;
  auto kern = assemble_tiled_kernel<T, CONSUMERS, ORDER, SRC, HAS_MAT>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return TFEM_ERR_LAUNCH;
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CONSUMERS + 32, smem) != cudaSuccess || per_sm < 1)
    return TFEM_ERR_LAUNCH;
  int64_t resident = (int64_t)sms * per_sm;  // persistent grid: every CTA is co-resident
  if (hp->reserve_ctas > 0 && resident > hp->reserve_ctas) resident -= hp->reserve_ctas;  // room for concurrent kernels
  const unsigned grid = (unsigned)(hp->n_tiles < resident ? hp->n_tiles : resident);
  kern<<<grid, CONSUMERS + 32, smem, s>>>((int)hp->n_tiles, hp->tile_list, hp->e_off, hp->e_blob, hp->l_off, hp->l_blob, hp->max_vert,
                                         elem_stride, e_words, l_words, coords, quad, alpha, beta, src, csr_val, load);
  return check_launch();
}

template <typename T, int CONSUMERS, int ORDER>
int dispatch_tiled(const tfem_tile_plan* hp, const T* coords, const QuadT<T>& quad, T alpha, T beta,
                   const SourceT<T>& src, T* csr_val, T* load, cudaStream_t s) {
  const int kind = load ? src.kind : TFEM_SRC_NONE;
  if (csr_val) {
    if (kind == TFEM_SRC_SINSIN) return launch_tiled<T, CONSUMERS, ORDER, TFEM_SRC_SINSIN, true>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
    if (kind == TFEM_SRC_CONST) return launch_tiled<T, CONSUMERS, ORDER, TFEM_SRC_CONST, true>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
    return launch_tiled<T, CONSUMERS, ORDER, TFEM_SRC_NONE, true>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
  }
  if (kind == TFEM_SRC_SINSIN) return launch_tiled<T, CONSUMERS, ORDER, TFEM_SRC_SINSIN, false>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
  if (kind == TFEM_SRC_CONST) return launch_tiled<T, CONSUMERS, ORDER, TFEM_SRC_CONST, false>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
  return TFEM_ERR_BAD_ARG;
}

template <typename T>
int assemble_tiled(const tfem_tile_plan* hp, const T* coords, int quad_order, const tfem_bilinear* form,
                   const tfem_source* source, T* csr_val, T* load, void* stream) {
  if (!hp || hp->n_tiles < 0) return TFEM_ERR_BAD_ARG;
  if (hp->n_tiles == 0) return TFEM_OK;
  if (!coords || (!csr_val && !load)) return TFEM_ERR_BAD_ARG;
  if (csr_val && !form) return TFEM_ERR_BAD_ARG;
  if (!hp->e_off || !hp->e_blob || !hp->l_off || !hp->l_blob) return TFEM_ERR_BAD_ARG;
  if (hp->max_vert > 1024 || hp->max_elem > 4096 || hp->elem_stride < hp->max_elem || 9 * hp->elem_stride > 65535)
    return TFEM_ERR_TOO_LARGE;
  if (hp->n_tiles > kMaxIndex) return TFEM_ERR_TOO_LARGE;
  if (tri_n_q(quad_order) == 0) return TFEM_ERR_UNSUPPORTED;
  SourceT<T> src = make_source<T>(source);
  if (load && (src.kind == TFEM_SRC_SAMPLED || src.kind < TFEM_SRC_NONE || src.kind > TFEM_SRC_SINSIN))
    return TFEM_ERR_BAD_ARG;  // sampled sources go through tfem_tri_p1_local_forms
  if (load && src.kind == TFEM_SRC_NONE) {  // f == 0: still write every row of the load vector
    src.kind = TFEM_SRC_CONST;
    src.p0 = T(0);
  }
  const QuadT<T> quad = make_quad<T>(quad_order);
  const T alpha = form ? T(form->alpha) : T(0), beta = form ? T(form->beta) : T(0);
  auto s = static_cast<cudaStream_t>(stream);
  // consumer threads per CTA: one tile element per thread where the tile allows it
  int consumers = hp->consumer_threads;
  if (consumers == 0) consumers = 256;  // measured best on B200 with ~160-row tiles (3 CTAs/SM); 128/384/512 selectable
#define TFEM_DISPATCH_ORDER(C)                                                                             \
  switch (quad_order) {                                                                                    \
    case 1: return dispatch_tiled<T, C, 1>(hp, coords, quad, alpha, beta, src, csr_val, load, s);         \
    case 2: return dispatch_tiled<T, C, 2>(hp, coords, quad, alpha, beta, src, csr_val, load, s);         \
    case 3: return dispatch_tiled<T, C, 3>(hp, coords, quad, alpha, beta, src, csr_val, load, s);         \
    default: return dispatch_tiled<T, C, 4>(hp, coords, quad, alpha, beta, src, csr_val, load, s);        \
  }
  if (consumers == 128) { TFEM_DISPATCH_ORDER(128) }
  if (consumers == 256) { TFEM_DISPATCH_ORDER(256) }
  if (consumers == 384) { TFEM_DISPATCH_ORDER(384) }
  if (consumers == 512) { TFEM_DISPATCH_ORDER(512) }
#undef TFEM_DISPATCH_ORDER
  return TFEM_ERR_BAD_ARG;
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
