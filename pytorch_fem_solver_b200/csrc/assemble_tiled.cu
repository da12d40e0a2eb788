// Fused P1 assembly to CSR + load vector: persistent, warp-specialised CTAs over row tiles
// (BASELINE.json config 2).
//
// Replaces, in one pass and without materialising any per-element tensor in HBM, the reference
// pipeline  Basis.__init__ geometry (basis/abstract_basis.py:42-63)  ->  user form evaluation
// (tests/test_assembly.py:68-84)  ->  (integrand*dx).sum(-3) (abstract_basis.py:83,104)  ->
// index_put_(accumulate=True) (abstract_basis.py:87-91,106-110).
//
// A tile owns a set of CSR rows; its index data are three contiguous, 16 B aligned blobs (layout in
// include/tfem_b200.h).  Each CTA is resident for the whole launch and walks tiles
// blockIdx.x, +gridDim.x, ...
//
//   producer warp (1 warp)                       consumer warps (8 warps)
//   ------------------------------------------   ------------------------------------------------
//   TMA bulk copies (cp.async.bulk + mbarrier     wait full[t]
//   complete_tx): E blob of tile t+2 when tile    A  rotate the tile's base sin/cos to every vertex
//   t-1 is done; LA blob (entries) of tile t      B  integrate every tile element ONCE: 6 matrix
//   as soon as every warp is through the             entries (symmetric form) + 3 load entries
//   entries of tile t-1; LB blob (rows) of           -> shared;  the last warp then evaluates the
//   tile t when tile t-1 is done                     full-range sin/cos at the base vertex of t+1
//   wait E blob t+1                               C  one lane per CSR ENTRY of the tile adds the
//   cp.async gather of the vertex coordinates        entry's (<= 2) contributions named by one
//   of tile t+1 (16 B per vertex) -> shared          packed word (no atomics, no read-modify-
//   arrive full[t+1]                                 write); a warp writes <= 32 consecutive
//                                                    csr_val slots (coalesced); one thread per
//                                                    owned row sums its load entry and diagonal
//                                                 every warp arrives done[t]
//
// so global-memory latency (blobs, coordinate gather) and the only library sin/cos of a tile are
// off the critical path.  f at the quadrature points is a short polynomial in the in-element phase
// around vertex 0 (coefficients from the vertex's sin/cos): no fp64 sin() per point.
// Elements on a tile border are recomputed by the neighbouring tile (halo ~15-18%).
// HBM traffic is coords + index blobs + outputs, each touched once.
#include "common.cuh"

// Profiling aid (never set in the shipped build): -DTFEM_DEBUG_SKIP=<mask> removes phases so their
// share of the kernel time can be measured; 1 = A (vertex sin/cos), 2 = B (integration), 4 = C,
// 8 = keep C's arithmetic but suppress its global stores.
#ifndef TFEM_DEBUG_SKIP
#define TFEM_DEBUG_SKIP 0
#endif
// tuning knobs of the shipped build (overridable for experiments)
#ifndef TFEM_MIN_CTAS
#define TFEM_MIN_CTAS 3  // resident CTAs per SM of the 256-consumer build (caps registers at 72)
#endif
#ifndef TFEM_LIGHT_UNROLL
#define TFEM_LIGHT_UNROLL 2  // segments a warp keeps in flight in the reduction phase
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
// Bounded wait: a protocol bug traps (after ~4 s of wall time) instead of hanging the GPU.  The
// suspend-time hint lets the hardware park the warp instead of polling.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t started = 0;
  for (uint32_t spin = 1; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    if ((spin & 0x3fffu) == 0) {  // every 16384 polls: look at the wall clock
      const uint64_t now = global_timer_ns();
      if (started == 0) started = now;
      else if (now - started > 4000000000ull) __trap();
    }
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
constexpr int kSmemHeader = 256;  // mbarriers (10 x 8 B) + two base-point records (2 x 6 values)

template <int ORDER> struct NQ;
template <> struct NQ<1> { static constexpr int value = 1; };
template <> struct NQ<2> { static constexpr int value = 3; };
template <> struct NQ<3> { static constexpr int value = 4; };
template <> struct NQ<4> { static constexpr int value = 6; };

__device__ __forceinline__ void sincos_full(double x, double& s, double& c) { sincos(x, &s, &c); }
__device__ __forceinline__ void sincos_full(float x, float& s, float& c) { sincosf(x, &s, &c); }

// Magnitude keys: order-preserving integer images of |v| (the high word for doubles), so that the
// largest of several magnitudes and a threshold test cost integer min/max instead of fp64 compares.
__device__ __forceinline__ int mag_key(double v) { return __double2hiint(v) & 0x7fffffff; }
__device__ __forceinline__ int mag_key(float v) { return __float_as_int(v) & 0x7fffffff; }
template <typename T> struct MagKey;
template <> struct MagKey<double> {  // high words of 0.008, 0.02, 0.03, 0.1 (low word dropped: slightly stricter)
  static constexpr int k008 = 0x3F80624D, k02 = 0x3F947AE1, k03 = 0x3F9EB851, k1 = 0x3FB99999;
};
template <> struct MagKey<float> {
  static constexpr int k008 = 0x3C03126E, k02 = 0x3CA3D70A, k03 = 0x3CF5C28F, k1 = 0x3DCCCCCC;
};

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
  const int key = mag_key(th);
  if (key < MagKey<T>::k1) {
    T st, ct;
    if (key < MagKey<T>::k02) sincos_small(th, st, ct);  // the usual case: a tile spans a few elements
    else sincos_medium(th, st, ct);
    s = fma(sb, ct, cb * st);
    c = fma(cb, ct, -(sb * st));
  } else {
    sincos_full(w * x, s, c);
  }
}

struct EView {  // "early" blob: header + vertices + connectivity
  int n_vert, n_elem, n_rows, n_runs, n_out, n_chunks, base_vertex, n_heavy_contrib, n_heavy;
  const int32_t* vert;
  const uint32_t* elem;
};

__device__ __forceinline__ EView view_e(const int32_t* b) {
  EView v;
  v.n_vert = b[0]; v.n_elem = b[1]; v.n_rows = b[2]; v.n_runs = b[3];
  v.n_out = b[4]; v.n_chunks = b[5]; v.base_vertex = b[6]; v.n_heavy_contrib = b[7]; v.n_heavy = b[8];
  v.vert = b + kBlobHeader;
  v.elem = reinterpret_cast<const uint32_t*>(v.vert + pad4(v.n_vert));
  return v;
}

struct LaView {  // "late" blob, entries: segments, one packed word per entry, heavy entries
  const int32_t* run_start;
  const int32_t* run_meta;
  const uint32_t* pair;
  const uint16_t* heavy_seg;
  const uint16_t* heavy_contrib;
  const uint32_t* heavy_pos;
};

__device__ __forceinline__ LaView view_la(const int32_t* b, const EView& e) {
  LaView v;
  v.run_start = b;
  v.run_meta = v.run_start + pad4(e.n_runs);
  const int32_t* p = v.run_meta + pad4(e.n_runs);
  v.pair = reinterpret_cast<const uint32_t*>(p);
  p += pad4(e.n_out);
  v.heavy_seg = reinterpret_cast<const uint16_t*>(p);
  p += pad4((e.n_heavy + 2) >> 1);
  v.heavy_contrib = reinterpret_cast<const uint16_t*>(p);
  p += pad4((e.n_heavy_contrib + 1) >> 1);
  v.heavy_pos = reinterpret_cast<const uint32_t*>(p);
  return v;
}

struct LbView {  // "late" blob, rows: row ids, per-row element chunks, diagonal positions
  const int32_t* row_id;
  const uint4* row_chunk;
  const uint32_t* row_diag;
};

__device__ __forceinline__ LbView view_lb(const int32_t* b, const EView& e) {
  LbView v;
  v.row_id = b;
  const int32_t* p = b + pad4(e.n_rows);
  v.row_chunk = reinterpret_cast<const uint4*>(p);  // 16 B aligned: every section is padded to 4 words
  p += 4 * e.n_chunks;
  v.row_diag = reinterpret_cast<const uint32_t*>(p);
  return v;
}

// sin(a + t) as a polynomial in t with coefficients k[] = {sin a, cos a, -sin a/2, -cos a/6, ...}
template <typename T, int DEG>
__device__ __forceinline__ void shifted_sine_coefficients(T s, T c, T (&k)[DEG + 1]) {
  constexpr double inv_fact[8] = {1.0, 1.0, -1.0 / 2.0, -1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, -1.0 / 720.0, -1.0 / 5040.0};
#pragma unroll
  for (int d = 0; d <= DEG; ++d) k[d] = d < 2 ? (d == 0 ? s : c) : T(inv_fact[d]) * ((d & 1) ? c : s);
}

// sum_q (w_q l_i(q)) sin(X0 + tx_q) sin(Y0 + ty_q), the phases tx, ty small: Horner of degree DEG
template <typename T, int ORDER, int DEG>
__device__ __forceinline__ void sinsin_moments(T sx0, T cx0, T sy0, T cy0, T uax, T ubx, T uay, T uby, T& m0, T& m1, T& m2) {
  const TriTable tt = tri_table(ORDER);  // folded at compile time
  T kx[DEG + 1], ky[DEG + 1];
  shifted_sine_coefficients<T, DEG>(sx0, cx0, kx);
  shifted_sine_coefficients<T, DEG>(sy0, cy0, ky);
  m0 = m1 = m2 = T(0);
#pragma unroll
  for (int q = 0; q < NQ<ORDER>::value; ++q) {
    const T l1 = T(tt.xi[q]), l2 = T(tt.eta[q]), l0 = T(1.0) - l1 - l2, w = T(0.5) * T(tt.w[q]);
    const T tx = fma(l1, uax, l2 * ubx), ty = fma(l1, uay, l2 * uby);
    T px = kx[DEG], py = ky[DEG];
#pragma unroll
    for (int d = DEG - 1; d >= 0; --d) {
      px = fma(px, tx, kx[d]);
      py = fma(py, ty, ky[d]);
    }
    const T f = px * py;
    m0 = fma(w * l0, f, m0);
    m1 = fma(w * l1, f, m1);
    m2 = fma(w * l2, f, m2);
  }
}

template <typename T, int CONSUMERS, int ORDER, int SRC, bool HAS_MAT>
__global__ void __launch_bounds__(CONSUMERS + 32, (CONSUMERS == 128 ? TFEM_MIN_CTAS_128 : (CONSUMERS == 192 ? 4 : (CONSUMERS == 256 ? TFEM_MIN_CTAS : 2)))) assemble_tiled_kernel(
    const int n_tiles, const int32_t* __restrict__ tile_list, const int32_t* __restrict__ e_off, const int32_t* __restrict__ e_blob,
    const int32_t* __restrict__ la_off, const int32_t* __restrict__ la_blob, const int32_t* __restrict__ lb_off,
    const int32_t* __restrict__ lb_blob, const int max_vert, const int elem_stride, const int e_words,
    const int la_words, const int lb_words, uint32_t* __restrict__ progress, const int n_progress,
    const T* __restrict__ coords,
    const QuadT<T> quad, const T alpha, const T beta, const SourceT<T> src, T* __restrict__ csr_val,
    T* __restrict__ load) {
  constexpr int NQV = NQ<ORDER>::value;
  constexpr bool HAS_LOAD = SRC != TFEM_SRC_NONE;
  constexpr bool SINSIN = SRC == TFEM_SRC_SINSIN;
  using V2 = typename Vec2<T>::type;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [ header | E blob x3 | LA blob | LB blob | vertex coordinates x2 | sin/cos fields | sloc[9][elem_stride] ]
  constexpr int kWarps = CONSUMERS / 32;
  uint64_t* e_bar = reinterpret_cast<uint64_t*>(smem_raw);  // [3] E blob landed
  uint64_t* la_bar = e_bar + kEStages;                      // LA blob landed
  uint64_t* lb_bar = la_bar + 1;                            // LB blob landed
  uint64_t* full_bar = lb_bar + 1;                          // [2] tile staged (coords + base point)
  uint64_t* done_bar = full_bar + 2;                        // [2] consumers finished the tile
  uint64_t* light_bar = done_bar + 2;                       // consumers are through the LA blob
  T* sbase = reinterpret_cast<T*>(smem_raw + 128);          // [2][6] bx, by, sin/cos(w bx), sin/cos(w by)
  int32_t* s_e = reinterpret_cast<int32_t*>(smem_raw + kSmemHeader);
  int32_t* s_la = s_e + kEStages * e_words;
  int32_t* s_lb = s_la + la_words;
  V2* vxy = reinterpret_cast<V2*>(s_lb + lb_words);  // [2][max_vert]
  T* trig = reinterpret_cast<T*>(vxy + 2 * max_vert);  // [4][max_vert]
  T* sloc = trig + (SINSIN ? 4 : 0) * max_vert;        // [9][elem_stride]

  const int tid = threadIdx.x;
  const int n_local = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0) {
    for (int i = 0; i < kEStages + 2; ++i) mbar_init(e_bar + i, 1);  // E, LA, LB: one expect_tx arrival each
    // full: the producer (coordinates) and, with a sin-sin source, the consumer warp that evaluates the
    // base point; done / light: one arrival per consumer warp
    for (int i = 0; i < 2; ++i) mbar_init(full_bar + i, SINSIN ? 2 : 1);
    for (int i = 0; i < 2; ++i) mbar_init(done_bar + i, kWarps);
    mbar_init(light_bar, kWarps);
    fence_mbar_init();
  }
  // the last column of the local-matrix table is never written by the integration phase: it is the
  // "no contribution" target of the packed index words
  if (tid < 9) sloc[tid * elem_stride + elem_stride - 1] = T(0);
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
    auto issue_l = [&](int it, const int32_t* off, const int32_t* blob, int32_t* dst, uint64_t* bar) {
      const int slot_index = (int)blockIdx.x + it * (int)gridDim.x;
      const int tile = tile_list ? __ldg(tile_list + slot_index) : slot_index;
      const int off0 = __ldg(off + tile);
      const uint32_t bytes = (uint32_t)(__ldg(off + tile + 1) - off0) * 4u;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(dst, blob + off0, bytes, bar);
    };
    if (lane == 0) {
      issue_e(0);
      issue_l(0, la_off, la_blob, s_la, la_bar);
      issue_l(0, lb_off, lb_blob, s_lb, lb_bar);
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
      TFEM_T(2);
      cp_async_wait_all();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar + buf);
      TFEM_T(3);
      if (it >= 1) {
        // the single LA buffer is free once every warp is through the entries of tile it-1, the LB
        // buffer and the E slot of tile it+2 once tile it-1 is finished
        mbar_wait(light_bar, (it - 1) & 1);
        if (lane == 0) issue_l(it, la_off, la_blob, s_la, la_bar);
        mbar_wait(done_bar + ((it - 1) & 1), ((it - 1) >> 1) & 1);
        if (lane == 0) issue_l(it, lb_off, lb_blob, s_lb, lb_bar);
      }
      if (lane == 0 && it + 2 < n_local) issue_e(it + 2);
      TFEM_T(4);
    }
#ifdef TFEM_DEBUG_TIMING
    if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 151))
      printf("cta %d producer (%d tiles): wait done(it-2) %lld | wait E %lld | issue gather %lld | gather landed %lld | wait light/done(it-1)+issue %lld\n",
             (int)blockIdx.x, n_local, t_acc[0], t_acc[1], t_acc[2], t_acc[3], t_acc[4]);
#endif
    return;
  }

  // ===================================== consumer warps ========================================
  // The tile's only full-range sin/cos (at its base vertex) is evaluated one tile ahead by the last
  // consumer warp: lane 0 takes the x phase, lane 1 the y phase.
  auto stage_base = [&](int nt) {
    if constexpr (SINSIN) {
      const int lane = tid & 31;
      mbar_wait(e_bar + nt % kEStages, (nt / kEStages) & 1);
      T bx, by, s, c;
      load_xy(coords, s_e[(nt % kEStages) * e_words + 6], bx, by);
      sincos_full(lane == 0 ? src.p1 * bx : src.p2 * by, s, c);
      T* sb = sbase + 6 * (nt & 1);
      if (lane == 0) {
        sb[0] = bx; sb[1] = by; sb[2] = s; sb[3] = c;
      } else if (lane == 1) {
        sb[4] = s; sb[5] = c;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar + (nt & 1));
    }
  };
  if (tid >= CONSUMERS - 32 && n_local > 0) stage_base(0);
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
    const TriTable tt = tri_table(ORDER);  // folded at compile time
    for (int el = tid; el < ((TFEM_DEBUG_SKIP & 2) ? 0 : ev.n_elem); el += CONSUMERS) {
      const uint32_t packed = ev.elem[el];
      const int a = packed & 1023u, b = (packed >> 10) & 1023u, c = (packed >> 20) & 1023u;
      const V2 p0 = xy[a], p1 = xy[b], p2 = xy[c];
      const T x0 = p0.x, y0 = p0.y;
      const T ax = p1.x - x0, ay = p1.y - y0;  // J = [[ax, bx], [ay, by]] (basis.py:87-88)
      const T bx = p2.x - x0, by = p2.y - y0;
      const T det = ax * by - bx * ay;  // signed (element_tri.py:139)
      if constexpr (HAS_MAT) {
        // grad(phi_i) = e_i / det with e_1 = (by, -bx), e_2 = (-ay, ax), e_0 = -e_1 - e_2, so
        // sum_q dx grad(phi_i).grad(phi_j) = (wsum / det) e_i.e_j ; mass = det * reference mass
        const T kc = (alpha * quad.wsum) / det;
        const T mb = beta * det;
        const T md = mb * quad.mref[0], mo = mb * quad.mref[1];
        const T e0x = ay - by, e0y = bx - ax;
        sloc[0 * elem_stride + el] = fma(kc, fma(e0x, e0x, e0y * e0y), md);
        sloc[1 * elem_stride + el] = fma(kc, fma(by, by, bx * bx), md);
        sloc[2 * elem_stride + el] = fma(kc, fma(ay, ay, ax * ax), md);
        sloc[3 * elem_stride + el] = fma(kc, fma(e0x, by, -(e0y * bx)), mo);
        sloc[4 * elem_stride + el] = fma(-kc, fma(by, ay, bx * ax), mo);
        sloc[5 * elem_stride + el] = fma(kc, fma(e0y, ax, -(e0x * ay)), mo);
      }
      if constexpr (HAS_LOAD) {
        T b0, b1, b2;
        if constexpr (SINSIN) {
          // phase of the source relative to vertex 0: w*(x_q - x0) = xi*(w ax) + eta*(w bx)
          const T uax = src.p1 * ax, ubx = src.p1 * bx, uay = src.p2 * ay, uby = src.p2 * by;
          const int reach = max(max(mag_key(uax), mag_key(ubx)), max(mag_key(uay), mag_key(uby)));  // >= every |phase|
          const T sx0 = trig[0 * max_vert + a], cx0 = trig[1 * max_vert + a];
          const T sy0 = trig[2 * max_vert + a], cy0 = trig[3 * max_vert + a];
          if (reach < MagKey<T>::k008) {  // |phase| < 0.008: truncation reach^6/720 < 4e-16
            sinsin_moments<T, ORDER, 5>(sx0, cx0, sy0, cy0, uax, ubx, uay, uby, b0, b1, b2);
          } else if (reach < MagKey<T>::k03) {  // |phase| < 0.03: reach^8/40320 < 2e-17
            sinsin_moments<T, ORDER, 7>(sx0, cx0, sy0, cy0, uax, ubx, uay, uby, b0, b1, b2);
          } else {  // coarse element: evaluate the source directly
            b0 = b1 = b2 = T(0);
#pragma unroll 1
            for (int q = 0; q < NQV; ++q) {
              const T sx = sin(src.p1 * fma(quad.l1[q], ax, fma(quad.l2[q], bx, x0)));
              const T sy = sin(src.p2 * fma(quad.l1[q], ay, fma(quad.l2[q], by, y0)));
              const T wf = quad.w[q] * (sx * sy);
              b0 = fma(wf, quad.l0[q], b0);
              b1 = fma(wf, quad.l1[q], b1);
              b2 = fma(wf, quad.l2[q], b2);
            }
          }
        } else {  // constant source: the moments of the basis functions are compile-time constants
          b0 = b1 = b2 = T(0);
#pragma unroll
          for (int q = 0; q < NQV; ++q) {
            const T l1 = T(tt.xi[q]), l2 = T(tt.eta[q]), l0 = T(1.0) - l1 - l2, w = T(0.5) * T(tt.w[q]);
            b0 += w * l0;
            b1 += w * l1;
            b2 += w * l2;
          }
        }
        const T amp = src.p0 * det;
        sloc[6 * elem_stride + el] = amp * b0;
        sloc[7 * elem_stride + el] = amp * b1;
        sloc[8 * elem_stride + el] = amp * b2;
      }
    }
    // the last warp integrates the fewest elements (tile elements rarely fill the last round): it
    // evaluates the next tile's base point while the others finish
    if (tid >= CONSUMERS - 32 && it + 1 < n_local) stage_base(it + 1);
    consumer_sync<CONSUMERS>();

    TFEM_T(2);
    // ---- C: one thread per CSR entry / per load entry; contributions in increasing element id --
    mbar_wait(la_bar, it & 1);
    TFEM_T(3);
    const LaView lv = view_la(s_la, ev);
    auto store = [&](int64_t pos, T value) {
      if (!(TFEM_DEBUG_SKIP & 8) || value == T(1.2345e30)) csr_val[pos] = value;
    };
    if constexpr (HAS_MAT) {
      // Entries with <= 2 contributions (every off-diagonal of a manifold mesh): one lane per entry,
      // one warp per segment of consecutive csr_val slots (coalesced stores).  One word per entry
      // names both contributions; a missing one points at the zero column, so there is no count
      // and no inner loop.
      // Runs are cut into segments of <= 32 entries, one warp pass each.
      const int lane = tid & 31, warp = tid >> 5;
      constexpr int kLightUnroll = TFEM_LIGHT_UNROLL;
#pragma unroll kLightUnroll
      for (int sg = warp; sg < ((TFEM_DEBUG_SKIP & 4) ? 0 : ev.n_runs); sg += kWarps) {
        const uint32_t meta = (uint32_t)lv.run_meta[sg];
        if (lane < (int)(meta >> 16)) {
          const uint32_t word = lv.pair[(meta & 0xffffu) + lane];
          if (word != 0xffffffffu) store((uint32_t)lv.run_start[sg] + lane, sloc[word & 0xffffu] + sloc[word >> 16]);
        }
      }
      TFEM_T(6);
      // the few other entries with > 2 contributions (non-manifold edges, degenerate elements)
      for (int h = tid; h < ((TFEM_DEBUG_SKIP & 4) ? 0 : ev.n_heavy); h += CONSUMERS) {
        T acc = T(0);
        for (int s = lv.heavy_seg[h]; s < lv.heavy_seg[h + 1]; ++s) acc += sloc[lv.heavy_contrib[s]];
        store(lv.heavy_pos[h], acc);
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(light_bar);  // this warp no longer reads the LA buffer
    mbar_wait(lb_bar, it & 1);
    const LbView lr = view_lb(s_lb, ev);
    TFEM_T(7);
    // one thread per owned row: its load entry and its diagonal share one element list, read as
    // chunks of 7 codes + link (one 16 B shared-memory load per chunk)
    for (int j = tid; j < ((TFEM_DEBUG_SKIP & 4) ? 0 : ev.n_rows); j += CONSUMERS) {
      T rhs = T(0), diag = T(0);
      uint32_t chunk = (uint32_t)j;
      do {
        const uint4 w = lr.row_chunk[chunk];
        const uint32_t code[7] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16, w.z & 0xffffu, w.z >> 16, w.w & 0xffffu};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          if constexpr (HAS_LOAD) rhs += sloc[code[k] + 6 * elem_stride];
          if constexpr (HAS_MAT) diag += sloc[code[k]];
        }
        chunk = w.w >> 16;
      } while (chunk != 0);
      if constexpr (HAS_LOAD) {
        if (!(TFEM_DEBUG_SKIP & 8) || rhs == T(1.2345e30)) load[lr.row_id[j]] = rhs;
      }
      if constexpr (HAS_MAT) {
        const uint32_t pos = lr.row_diag[j];
        if (pos != 0xffffffffu) store(pos, diag);
      }
    }
    TFEM_T(4);
    // sloc / blob slots / vxy[buf] are free once every warp is here.  With a sin-sin source the next
    // tile's phase A only writes trig (not read in phase C) and ends in a block barrier before sloc is
    // written again, so each warp just reports to the producer; otherwise the block barrier stays.
    if constexpr (!SINSIN) consumer_sync<CONSUMERS>();
    if (progress != nullptr && (int)blockIdx.x + it * (int)gridDim.x < n_progress) {
      // the first tiles of the call hold the multi-GPU interface rows: tell the exchange kernels
      // waiting on the counter (tfem_iface_pack_after) that this warp's stores are out
      __threadfence();
      __syncwarp();
      if ((tid & 31) == 0) atomicAdd(progress, 1u);
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(done_bar + buf);
    TFEM_T(5);
  }
#ifdef TFEM_DEBUG_TIMING
  if ((tid == 0 || tid == 255) && (blockIdx.x == 0 || blockIdx.x == 151))
    printf("cta %d consumer t%d (%d tiles): wait full %lld | A+sync %lld | B+sync %lld | wait L %lld | C light %lld heavy %lld rows %lld | end sync %lld\n",
           (int)blockIdx.x, tid, n_local, t_acc[0], t_acc[1], t_acc[2], t_acc[3], t_acc[6], t_acc[7], t_acc[4], t_acc[5]);
#endif
}

template <typename T>
size_t tiled_smem_bytes(const tfem_tile_plan* hp, int src_kind, int* elem_stride, int* e_words, int* la_words, int* lb_words) {
  const int trig_fields = src_kind == TFEM_SRC_SINSIN ? 4 : 0;
  *elem_stride = hp->elem_stride;
  *e_words = (hp->max_e_words + 3) & ~3;
  *la_words = (hp->max_la_words + 3) & ~3;
  *lb_words = (hp->max_lb_words + 3) & ~3;
  return kSmemHeader + 4 * ((size_t)kEStages * *e_words + (size_t)*la_words + (size_t)*lb_words) +
         sizeof(T) * ((size_t)(4 + trig_fields) * hp->max_vert + (size_t)9 * *elem_stride);
}

template <typename T, int CONSUMERS, int ORDER, int SRC, bool HAS_MAT>
int launch_tiled(const tfem_tile_plan* hp, const T* coords, const QuadT<T>& quad, T alpha, T beta,
                 const SourceT<T>& src, T* csr_val, T* load, cudaStream_t s) {
  int elem_stride = 0, e_words = 0, la_words = 0, lb_words = 0;
  const size_t smem = tiled_smem_bytes<T>(hp, SRC, &elem_stride, &e_words, &la_words, &lb_words);
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
  kern<<<grid, CONSUMERS + 32, smem, s>>>((int)hp->n_tiles, hp->tile_list, hp->e_off, hp->e_blob, hp->la_off, hp->la_blob, hp->lb_off,
                                         hp->lb_blob, hp->max_vert, elem_stride, e_words, la_words, lb_words,
                                         hp->n_progress_tiles > 0 ? hp->progress : nullptr, hp->n_progress_tiles, coords, quad, alpha, beta, src, csr_val, load);
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
  if (!hp->e_off || !hp->e_blob || !hp->la_off || !hp->la_blob || !hp->lb_off || !hp->lb_blob) return TFEM_ERR_BAD_ARG;
  if (hp->max_vert > 1024 || hp->max_elem > 4096 || hp->elem_stride <= hp->max_elem || 9 * hp->elem_stride > 65535)
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
  if (consumers == 0) consumers = 256;  // with ~192-row tiles (3 CTAs/SM); 128/192/384/512 selectable, all within 5 % on B200
#define TFEM_DISPATCH_ORDER(C)                                                                             \
  switch (quad_order) {                                                                                    \
    case 1: return dispatch_tiled<T, C, 1>(hp, coords, quad, alpha, beta, src, csr_val, load, s);         \
    case 2: return dispatch_tiled<T, C, 2>(hp, coords, quad, alpha, beta, src, csr_val, load, s);         \
    case 3: return dispatch_tiled<T, C, 3>(hp, coords, quad, alpha, beta, src, csr_val, load, s);         \
    default: return dispatch_tiled<T, C, 4>(hp, coords, quad, alpha, beta, src, csr_val, load, s);        \
  }
  if (consumers == 128) { TFEM_DISPATCH_ORDER(128) }
  if (consumers == 192) { TFEM_DISPATCH_ORDER(192) }
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
