// Fused P1 assembly to CSR + load vector: persistent, warp-specialised CTAs over row tiles
// (BASELINE.json config 2), tile plan format v2 (include/tfem_b200.h).
//
// Replaces, in one pass and without materialising any per-element tensor in HBM, the reference
// pipeline  Basis.__init__ geometry (basis/abstract_basis.py:42-63)  ->  user form evaluation
// (tests/test_assembly.py:68-84)  ->  (integrand*dx).sum(-3) (abstract_basis.py:83,104)  ->
// index_put_(accumulate=True) (abstract_basis.py:87-91,106-110).
//
// A tile owns a set of CSR rows.  Its index data are a TEMPLATE (tile-local connectivity and the
// (element, slot) contributions of every CSR entry / row; shared by all congruent tiles and kept
// RESIDENT in shared memory) and a small per-tile INSTANCE (vertex ids, CSR offsets of the entry
// segments, row ids).  Each CTA is resident for the whole launch and walks a contiguous block of the
// tile list, so on a lattice-numbered mesh it fetches a template once and then streams instances.
//
//   producer warp (1 warp)                        consumer warps (CONSUMERS / 32)
//   -------------------------------------------   -----------------------------------------------
//   TMA bulk copy (cp.async.bulk + mbarrier        wait full[t]
//   complete_tx) of the instance of tile t+1       B  integrate every tile element ONCE: 6 matrix
//   cp.async gather of the vertex coordinates         entries (symmetric form) + 3 load entries
//   of tile t (16 B per vertex) -> shared             -> shared table [element][9]
//   template changed?  TB part after phase B of    block barrier
//   tile t-1, TC part after its phase C            C  one lane per CSR ENTRY of the tile adds the
//   sin/cos of the source at the tile's base          entry's (<= 2) contributions named by one
//   vertex (the only library sincos of a tile)        packed word (no atomics, no read-modify-
//   arrive full[t]                                    write); a warp writes <= 32 consecutive
//                                                     csr_val slots (coalesced); one thread per
//                                                     owned row sums its load entry and diagonal
//                                                  block barrier; every warp arrives done[t]
//
// f at the quadrature points: sin/cos at the element's centroid by rotating the base vertex's
// values through a short polynomial, then a degree-4/6/8 expansion about the centroid (chosen per warp
// from the element size) -- no fp64 sin() per point, no per-vertex pass.
// Elements on a tile border are recomputed by the neighbouring tile (halo ~11%).
// HBM traffic is coords + instances + outputs, each touched once.
#include <type_traits>

#include "common.cuh"

// Profiling aid (never set in the shipped build): -DTFEM_DEBUG_SKIP=<mask> removes phases so their
// share of the kernel time can be measured; 2 = B (integration), 4 = C (reduction),
// 8 = keep C's arithmetic but suppress its global stores, 16 = B without the load vector's source.
#ifndef TFEM_DEBUG_SKIP
#define TFEM_DEBUG_SKIP 0
#endif
#ifndef TFEM_WS_INST_STAGES
#define TFEM_WS_INST_STAGES 6  // instance blobs in flight in the role-specialised kernel (stages + prefetch distance)
#endif
#ifndef TFEM_SEG_UNROLL
#define TFEM_SEG_UNROLL 2  // segments a warp keeps in flight in the reduction phase
#endif

// -DTFEM_DEBUG_TIMING: consumer thread 0 and the producer's lane 0 of every CTA accumulate the cycles
// they spend per phase into a device array read back with tfem_debug_timing_read().  Profiling aid only.
#ifdef TFEM_DEBUG_TIMING
#define TFEM_T_DECL long long t_mark = clock64(), t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TFEM_T(i)                      \
  do {                                 \
    const long long t_now = clock64(); \
    t_acc[i] += t_now - t_mark;        \
    t_mark = t_now;                    \
  } while (0)
__device__ long long tfem_timing_table[2][8];
__device__ int tfem_timing_ctas;
#else
#define TFEM_T_DECL
#define TFEM_T(i)
#endif

#ifdef TFEM_STAGGER_NS
__device__ int tfem_stagger_slots[1024];
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
#ifndef TFEM_WAIT_SLEEP_NS
#define TFEM_WAIT_SLEEP_NS 0
#endif
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if TFEM_WAIT_SLEEP_NS > 0
  // poll at a fixed interval instead of the hardware's wake-on-every-arrival (a barrier with N arrivals per phase wakes
  // each waiting warp N times: issue slots the integration warps need)
  if (mbar_test(bar, parity)) return;
  uint64_t started_at = 0;
  for (uint32_t spin = 1;; ++spin) {
    __nanosleep(TFEM_WAIT_SLEEP_NS);
    if (mbar_test(bar, parity)) return;
    if ((spin & 0x3fffu) == 0) {
      const uint64_t now = global_timer_ns();
      if (started_at == 0) started_at = now;
      else if (now - started_at > 4000000000ull) __trap();
    }
  }
#endif
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
#ifndef TFEM_DEBUG_NO_BARRIER  // (timing experiment only: results are wrong without the barriers)
  asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
#endif
}

__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }
constexpr int kInstHeader = 4;
constexpr int kTbHeader = 8;
constexpr int kInstStages = 4;    // instance blobs in flight: tile t+1 prefetched, t being staged, t-1 and t-2 with the consumers
constexpr int kDlSlots = 6;       // one "dl" table row per tile element: d0 b0 d1 b1 d2 b2 (include/tfem_b200.h)
constexpr int kStages = 3;        // tiles staged ahead of the consumers (coordinates, record, base point)
constexpr int kSmemHeader = 512;  // mbarriers, stage records, base-point records

// ---- shared-memory accesses by 32-bit address (no generic-pointer arithmetic in the hot loops) --
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 lds_v4i(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void lds_2(uint32_t a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ void lds_2(uint32_t a, float& x, float& y) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
}
__device__ __forceinline__ void lds_1(uint32_t a, double& x) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a)); }
__device__ __forceinline__ void lds_1(uint32_t a, float& x) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); }
__device__ __forceinline__ void sts_2(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void sts_2(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void sts_1(uint32_t a, double x) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory"); }
__device__ __forceinline__ void sts_1(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }
// plan codes are byte offsets into a table of fp64 slots; the fp32 table is half as wide
template <typename T>
__device__ __forceinline__ uint32_t code_offset(uint32_t code) { return sizeof(T) == 8 ? code : code >> 1; }

constexpr int kSrcResidual = 100;  // internal source kind of tfem_weak_residual_tiled (not part of tfem_source_kind)

template <int ORDER> struct NQ;
template <> struct NQ<1> { static constexpr int value = 1; };
template <> struct NQ<2> { static constexpr int value = 3; };
template <> struct NQ<3> { static constexpr int value = 4; };
template <> struct NQ<4> { static constexpr int value = 6; };

__device__ __forceinline__ void sincos_full(double x, double& s, double& c) { sincos(x, &s, &c); }
__device__ __forceinline__ void sincos_full(float x, float& s, float& c) { sincosf(x, &s, &c); }

// Magnitude keys: order-preserving integer images of |v| (the high word for doubles), so that
// threshold tests cost integer max / compare instead of fp64 instructions.
__device__ __forceinline__ int hi_word(double v) { return __double2hiint(v); }
__device__ __forceinline__ int hi_word(float v) { return __float_as_int(v); }
__device__ __forceinline__ int mag_key(double v) { return __double2hiint(v) & 0x7fffffff; }
__device__ __forceinline__ int mag_key(float v) { return __float_as_int(v) & 0x7fffffff; }
inline int host_mag_key(double v, double) {
  uint64_t bits;
  memcpy(&bits, &v, 8);
  return (int)((bits >> 32) & 0x7fffffffu);
}
inline int host_mag_key(double v, float) {
  const float f = (float)v;
  uint32_t bits;
  memcpy(&bits, &f, 4);
  return (int)(bits & 0x7fffffffu);
}

// 1/d without the range checks of the library division: MUFU seed (relative error < 2^-20) and one cubic
// Newton step r (1 + e + e^2): truncation e^3 < 1e-18, i.e. the result is within ~2 ulp (inf / NaN for d == 0
// like the reference's division).
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  e = fma(e, e, e);
  return fma(r, e, r);
}
__device__ __forceinline__ float fast_rcp(float d) { return __frcp_rn(d); }

// Every floating-point constant of the hot loop, passed as a kernel parameter: operands then come
// from the constant bank (one load feeds two constants) instead of two immediate moves per use.
template <typename T>
struct TiledConst {
  T w1_3, w2_3;                 // source frequencies / 3 (phase of the centroid)
  T s3, s5, s7, s9;             // sin(t) = t + t^3 (s3 + t^2 (s5 + t^2 (s7 + t^2 s9)))
  T c2, c4, c6, c8, c10;        // cos(t) = 1 + t^2 (c2 + t^2 (c4 + ...))
  T f[9];                       // signed inverse factorials of sin(a + t) = sum_d f[d] t^d {sin a | cos a}
  T kcw, md, mo;                // alpha * sum_q w_q ; beta * reference mass (diagonal, off-diagonal)
  T c1[kMaxQ], c2q[kMaxQ];      // barycentric offsets of the quadrature points from the centroid
  T wl0[kMaxQ], wl1[kMaxQ], wl2[kMaxQ];  // w_q * l_i(q)
  T wq[kMaxQ];                           // w_q
  T m0, m1, m2;                 // sum_q w_q l_i(q)  (constant source)
  T o3_w1, o3_w2;               // 4-point rule: source frequencies * 2/15 (offsets of its outer points from the centroid)
  T o3_a, o3_b, o3_c;           // 4-point rule: w_c / 3, w_o / 5, 2 w_o / 5
};

template <typename T>
TiledConst<T> make_tiled_const(int order, const QuadT<T>& quad, T alpha, T beta, const SourceT<T>& src) {
  TiledConst<T> c{};
  c.w1_3 = T(double(src.p1) / 3.0);
  c.w2_3 = T(double(src.p2) / 3.0);
  c.s3 = T(-1.0 / 6.0); c.s5 = T(1.0 / 120.0); c.s7 = T(-1.0 / 5040.0); c.s9 = T(1.0 / 362880.0);
  c.c2 = T(-0.5); c.c4 = T(1.0 / 24.0); c.c6 = T(-1.0 / 720.0); c.c8 = T(1.0 / 40320.0); c.c10 = T(-1.0 / 3628800.0);
  const double inv_fact[9] = {1.0, 1.0, -1.0 / 2.0, -1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, -1.0 / 720.0, -1.0 / 5040.0, 1.0 / 40320.0};
  for (int d = 0; d < 9; ++d) c.f[d] = T(inv_fact[d]);
  c.kcw = alpha * quad.wsum;
  c.md = beta * quad.mref[0];
  c.mo = beta * quad.mref[1];
  const TriTable tt = tri_table(order);
  double m0 = 0, m1 = 0, m2 = 0;
  for (int q = 0; q < tt.n_q; ++q) {
    const double l1 = tt.xi[q], l2 = tt.eta[q], l0 = 1.0 - l1 - l2, w = 0.5 * tt.w[q];
    c.c1[q] = T(l1 - 1.0 / 3.0);
    c.c2q[q] = T(l2 - 1.0 / 3.0);
    c.wl0[q] = T(w * l0); c.wl1[q] = T(w * l1); c.wl2[q] = T(w * l2);
    c.wq[q] = T(w);
    m0 += w * l0; m1 += w * l1; m2 += w * l2;
  }
  c.m0 = T(m0); c.m1 = T(m1); c.m2 = T(m2);
  c.o3_w1 = T(double(src.p1) * 2.0 / 15.0);
  c.o3_w2 = T(double(src.p2) * 2.0 / 15.0);
  c.o3_a = T(0.5 * (-9.0 / 16.0) / 3.0);
  c.o3_b = T(0.5 * (25.0 / 48.0) / 5.0);
  c.o3_c = T(0.5 * (25.0 / 48.0) * 2.0 / 5.0);
  return c;
}

// sin(t), cos(t) for |t| <= 0.04 (truncation t^7/5040 < 3.3e-14 and t^8/40320 < 2e-16, relative to 1)
template <typename T>
__device__ __forceinline__ void sincos_small(const TiledConst<T>& k, T t, T& s, T& c) {
  const T z = t * t;
  const T ps = fma(z, k.s5, k.s3);
  s = fma(t * z, ps, t);
  T pc = fma(z, k.c6, k.c4);
  pc = fma(z, pc, k.c2);
  c = fma(z, pc, T(1));
}

// sin(t), cos(t) for |t| <= 0.2 (truncation < 6e-16 relative to 1)
template <typename T>
__device__ __forceinline__ void sincos_medium(const TiledConst<T>& k, T t, T& s, T& c) {
  const T z = t * t;
  T ps = fma(z, k.s9, k.s7);
  ps = fma(z, ps, k.s5);
  ps = fma(z, ps, k.s3);
  s = fma(t * z, ps, t);
  T pc = fma(z, k.c10, k.c8);
  pc = fma(z, pc, k.c6);
  pc = fma(z, pc, k.c4);
  pc = fma(z, pc, k.c2);
  c = fma(z, pc, T(1));
}

// m_i = sum_q (w_q l_i(q)) sin(Xc + tx_q) sin(Yc + ty_q): the source about the element's CENTROID, whose
// sin/cos (sx, cx, sy, cy) are given; u** = frequency * edge vector.  Horner of degree DEG per point
// (sin(a + t) = sum_d f[d] t^d {sin a | cos a}); a point AT the centroid (the first point of the 4-point
// rule) needs no polynomial at all.
template <typename T, int ORDER, int DEG>
__device__ __forceinline__ void sinsin_moments(const TiledConst<T>& k, T sx, T cx, T sy, T cy, T uax, T ubx, T uay, T uby, T& m0, T& m1, T& m2) {
  const TriTable tt = tri_table(ORDER);  // folded at compile time (only used to spot a point at the centroid)
  T kx[DEG + 1], ky[DEG + 1];
#pragma unroll
  for (int d = 0; d <= DEG; ++d) {
    kx[d] = d < 2 ? (d == 0 ? sx : cx) : k.f[d] * ((d & 1) ? cx : sx);
    ky[d] = d < 2 ? (d == 0 ? sy : cy) : k.f[d] * ((d & 1) ? cy : sy);
  }
  m0 = m1 = m2 = T(0);
#pragma unroll
  for (int q = 0; q < NQ<ORDER>::value; ++q) {
    const double o1 = tt.xi[q] - 1.0 / 3.0, o2 = tt.eta[q] - 1.0 / 3.0;
    T f;
    if (o1 * o1 + o2 * o2 < 1e-24) {
      f = sx * sy;
    } else {
      const T tx = fma(k.c1[q], uax, k.c2q[q] * ubx), ty = fma(k.c1[q], uay, k.c2q[q] * uby);
      T px = kx[DEG], py = ky[DEG];
#pragma unroll
      for (int d = DEG - 1; d >= 0; --d) {
        px = fma(px, tx, kx[d]);
        py = fma(py, ty, ky[d]);
      }
      f = px * py;
    }
    m0 = fma(k.wl0[q], f, m0);
    m1 = fma(k.wl1[q], f, m1);
    m2 = fma(k.wl2[q], f, m2);
  }
}

// The same moments for the 4-point rule (element_tri.py:99-108: centroid, weight -9/16, and the three points
// centroid + 2/5 (vertex - centroid), weight 25/48) with its structure spelled out: the outer points' phases are
// 2 va - vb, 2 vb - va, -(va + vb) with v = 2/15 frequency * edge, and m_i = a f_c + b (f_1 + f_2 + f_3) + c f_i'.
// 17 fp64 operations and 20 constants fewer than the table-driven form; degree 4 (|phase| < 4e-3).
template <typename T>
__device__ __forceinline__ void sinsin_moments_o3(const TiledConst<T>& k, T sx, T cx, T sy, T cy, T ax, T bx, T ay, T by, T& m0, T& m1, T& m2) {
  const T vax = k.o3_w1 * ax, vbx = k.o3_w1 * bx, vay = k.o3_w2 * ay, vby = k.o3_w2 * by;
  const T kx2 = k.f[2] * sx, kx3 = k.f[3] * cx, kx4 = k.f[4] * sx;
  const T ky2 = k.f[2] * sy, ky3 = k.f[3] * cy, ky4 = k.f[4] * sy;
  auto shifted = [](T t, T k0, T k1, T k2, T k3, T k4) { return fma(fma(fma(fma(k4, t, k3), t, k2), t, k1), t, k0); };
  const T f1 = shifted(fma(T(2), vax, -vbx), sx, cx, kx2, kx3, kx4) * shifted(fma(T(2), vay, -vby), sy, cy, ky2, ky3, ky4);  // l = (.2, .6, .2)
  const T f2 = shifted(fma(T(2), vbx, -vax), sx, cx, kx2, kx3, kx4) * shifted(fma(T(2), vby, -vay), sy, cy, ky2, ky3, ky4);  // l = (.2, .2, .6)
  const T f3 = shifted(-(vax + vbx), sx, cx, kx2, kx3, kx4) * shifted(-(vay + vby), sy, cy, ky2, ky3, ky4);                  // l = (.6, .2, .2)
  const T common = fma(k.o3_b, (f1 + f2) + f3, k.o3_a * (sx * sy));
  m0 = fma(k.o3_c, f3, common);
  m1 = fma(k.o3_c, f1, common);
  m2 = fma(k.o3_c, f2, common);
}

// largest |barycentric offset| sum over the rule's points: |phase about the centroid| <= this * reach
inline double centroid_spread(int order) {
  const TriTable tt = tri_table(order);
  double worst = 0.0;
  for (int q = 0; q < tt.n_q; ++q) {
    const double c1 = tt.xi[q] - 1.0 / 3.0, c2 = tt.eta[q] - 1.0 / 3.0;
    const double s = (c1 < 0 ? -c1 : c1) + (c2 < 0 ? -c2 : c2);
    worst = s > worst ? s : worst;
  }
  return worst > 0.05 ? worst : 0.05;
}

template <typename T>
struct TiledArgs {
  int n_tiles, n_strided;
  const int32_t* tile_list;
  const int4* tile_desc;
  const int32_t* inst_blob;
  const int4* tpl_desc;
  const int32_t* tpl_blob;
  int max_vert, max_elem, inst_words, tb_words, tc_words;
  int od_base[3];  // byte offsets (in units of T: already scaled) of the off-diagonal arrays inside the table
  int table_stride;  // bytes of one table (role-specialised kernel: two of them)
  uint32_t* progress;
  const T* coords;
  const T* f_q;          // [n_el, n_q] source at the quadrature points (TFEM_SRC_SAMPLED), via the instances' elem_id section
  const T* grad_u;       // [n_el, n_q, d] (weak residual): grad u at the quadrature points, d = 3 on fractures else 2
  const T* frac_inv;     // [n_mesh, 2, 3] J_f^+ (weak residual on fractures)
  const T* frac_metric;  // [n_mesh, 4] (a00, a01, a11, det J_f) or NULL; element e lies on fracture e / n_el_per_mesh
  int n_el_per_mesh;
  SourceT<T> src;
  T* csr_val;
  T* load;
  // fast path of the source evaluation (sincos_small + degree 4) when  |phase base -> centroid| < key_rot
  // and  max squared edge length < key_len2  (both as integer images of the thresholds); else the
  // general path picks per warp: rotation small / medium / library, expansion degree 4 / 6 / 8 / sin() per point
  int key_rot, key_rot_medium, key_len2;
  int key_deg4, key_deg6, key_deg8;
};

// ---- phase B of one tile element: its 6 matrix entries (symmetric form) and 3 load entries -> table row `row` ----
// `slot` = the element's index in the tile (connectivity word, element id), `row` = its table row (the scratch row for
// the padding lanes of a warp: every lane computes, the votes below need them all).
template <typename T, int ORDER, int SRC, bool HAS_MAT, bool FRAC>
__device__ __forceinline__ void integrate_tile_element(const TiledArgs<T>& args, const TiledConst<T>& cst, const QuadT<T>& quad,
                                                       uint32_t a_xy, uint32_t a_sb, uint32_t a_eid, uint32_t a_elem, uint32_t a_tab,
                                                       uint32_t a_od0, uint32_t a_od1, uint32_t a_od2, uint32_t slot, uint32_t row) {
  constexpr int NQV = NQ<ORDER>::value;
  constexpr bool HAS_LOAD = SRC != TFEM_SRC_NONE;
  constexpr bool SINSIN = SRC == TFEM_SRC_SINSIN;
  constexpr bool SAMPLED = SRC == TFEM_SRC_SAMPLED;
  constexpr bool RESIDUAL = SRC == kSrcResidual;  // weak residual: f v - grad v . grad u, both given at the quadrature points
  constexpr bool NEED_IDS = SAMPLED || FRAC || RESIDUAL;
  constexpr uint32_t kRowBytes = kDlSlots * sizeof(T);
  constexpr uint32_t kS = sizeof(T);
  const uint32_t packed = lds_u32(a_elem + 4u * slot);
  const uint32_t out = a_tab + row * kRowBytes;
  // sampled source / fracture metric of this element, fetched first so the loads overlap the geometry
  T fq[NQV];
  T ru0 = T(0), ru1 = T(0), ru2 = T(0);
  T a00 = T(1), a01 = T(0), a11 = T(1), detf = T(1);
  (void)ru2;
  if constexpr (NEED_IDS) {
    const int64_t eid = (int64_t)lds_u32(a_eid + 4u * slot);
    if constexpr (SAMPLED) {
#pragma unroll
      for (int q = 0; q < NQV; ++q) fq[q] = __ldg(args.f_q + eid * NQV + q);
    }
    if constexpr (RESIDUAL) {
      // U = sum_q w_q grad u(x_q) (the basis gradients are constant on the element) and the samples of f
      constexpr int D = FRAC ? 3 : 2;
      const T* gu = args.grad_u + eid * (NQV * D);
      ru0 = ru1 = ru2 = T(0);
#pragma unroll
      for (int q = 0; q < NQV; ++q) {
        const T w = cst.wq[q];
        ru0 = fma(w, __ldg(gu + q * D), ru0);
        ru1 = fma(w, __ldg(gu + q * D + 1), ru1);
        if constexpr (FRAC) ru2 = fma(w, __ldg(gu + q * D + 2), ru2);
        fq[q] = args.f_q ? __ldg(args.f_q + eid * NQV + q) : T(0);
      }
      if constexpr (FRAC) {  // pull the 3-D vector back to the fracture's plane: J_f^+ U
        const T* ji = args.frac_inv + 6 * (eid / args.n_el_per_mesh);
        const T p0 = fma(__ldg(ji), ru0, fma(__ldg(ji + 1), ru1, __ldg(ji + 2) * ru2));
        const T p1 = fma(__ldg(ji + 3), ru0, fma(__ldg(ji + 4), ru1, __ldg(ji + 5) * ru2));
        ru0 = p0;
        ru1 = p1;
      }
    }
    if constexpr (FRAC) {
      const T* metric = args.frac_metric + 4 * (eid / args.n_el_per_mesh);
      a00 = __ldg(metric); a01 = __ldg(metric + 1); a11 = __ldg(metric + 2); detf = __ldg(metric + 3);
    }
  }
  T x0, y0, x1, y1, x2, y2;
  lds_2(a_xy + (packed & 1023u) * (2 * kS), x0, y0);
  lds_2(a_xy + ((packed >> 10) & 1023u) * (2 * kS), x1, y1);
  lds_2(a_xy + (packed >> 20) * (2 * kS), x2, y2);
  const T ax = x1 - x0, ay = y1 - y0;  // J = [[ax, bx], [ay, by]] (basis.py:87-88)
  const T bx = x2 - x0, by = y2 - y0;
  const T det = ax * by - bx * ay;  // signed (element_tri.py:139)
  // e_1 = (by, -bx), e_2 = (-ay, ax): e_i . e_j, on a fracture in the metric a = J_f^+ J_f^+^T of the plane
  T s11, s22;
  if constexpr (FRAC) {
    s11 = fma(a00 * by, by, fma(a11 * bx, bx, T(-2) * (a01 * by) * bx));
    s22 = fma(a00 * ay, ay, fma(a11 * ax, ax, T(-2) * (a01 * ay) * ax));
  } else {
    s11 = fma(by, by, bx * bx);
    s22 = fma(ay, ay, ax * ax);  // squared edge lengths
  }
  T k00 = T(0), k11 = T(0), k22 = T(0), k01 = T(0), k12 = T(0), k20 = T(0), b0 = T(0), b1 = T(0), b2 = T(0);
  if constexpr (HAS_MAT) {
    // grad(phi_i) = e_i / det with e_1 = (by, -bx), e_2 = (-ay, ax), e_0 = -e_1 - e_2, so
    // sum_q dx grad(phi_i).grad(phi_j) = (wsum / det) e_i.e_j ; mass = det * reference mass
    const T kc = (FRAC ? cst.kcw * detf : cst.kcw) * fast_rcp(det);
    const T area = FRAC ? det * detf : det;
    const T md = area * cst.md, mo = area * cst.mo;
    T s12;
    if constexpr (FRAC) s12 = fma(a01, fma(by, ax, bx * ay), -fma(a00 * by, ay, (a11 * bx) * ax));
    else s12 = -fma(by, ay, bx * ax);
    const T s01 = -s11 - s12, s02 = -s22 - s12, s00 = -s01 - s02;
    k00 = fma(kc, s00, md);
    k11 = fma(kc, s11, md);
    k22 = fma(kc, s22, md);
    k01 = fma(kc, s01, mo);
    k12 = fma(kc, s12, mo);
    k20 = fma(kc, s02, mo);
  }
  if constexpr (HAS_LOAD) {
    if constexpr (SINSIN) {
      if (TFEM_DEBUG_SKIP & 16) {
        b0 = ax; b1 = bx; b2 = ay + by;
      } else {
        // sin/cos of the source phase at the centroid: rotate the base vertex's values
        T pbx, pby, sbx, cbx, sby, cby;
        lds_2(a_sb, pbx, pby);
        lds_2(a_sb + 2 * kS, sbx, cbx);
        lds_2(a_sb + 4 * kS, sby, cby);
        const T dx = fma(cst.w1_3, (x0 + x1) + x2, -pbx);
        const T dy = fma(cst.w2_3, (y0 + y1) + y2, -pby);
        const int rot = max(mag_key(dx), mag_key(dy));
        const int len2 = max(hi_word(s11), hi_word(s22));  // non-negative: the high words order like the values
        T sx, cx, sy, cy;
        if (!__any_sync(0xffffffffu, (rot >= args.key_rot) | (len2 >= args.key_len2))) {
          // every lane of the warp: small rotation, small element
          T st, ct, su, cu;
          sincos_small(cst, dx, st, ct);
          sincos_small(cst, dy, su, cu);
          sx = fma(sbx, ct, cbx * st);
          cx = fma(cbx, ct, -(sbx * st));
          sy = fma(sby, cu, cby * su);
          cy = fma(cby, cu, -(sby * su));
          if constexpr (ORDER == 3) sinsin_moments_o3<T>(cst, sx, cx, sy, cy, ax, bx, ay, by, b0, b1, b2);
          else sinsin_moments<T, ORDER, 4>(cst, sx, cx, sy, cy, args.src.p1 * ax, args.src.p1 * bx, args.src.p2 * ay, args.src.p2 * by, b0, b1, b2);
        } else {
          const int rot_w = __reduce_max_sync(0xffffffffu, rot);
          if (rot_w >= args.key_rot_medium) {
            sincos_full(dx + pbx, sx, cx);
            sincos_full(dy + pby, sy, cy);
          } else {
            T st, ct, su, cu;
            sincos_medium(cst, dx, st, ct);
            sincos_medium(cst, dy, su, cu);
            sx = fma(sbx, ct, cbx * st);
            cx = fma(cbx, ct, -(sbx * st));
            sy = fma(sby, cu, cby * su);
            cy = fma(cby, cu, -(sby * su));
          }
          // expansion about the centroid, degree by the largest phase any lane of the warp needs (the frequency-scaled
          // edge vectors are formed here, behind a scheduling fence: the fast path above has no use for them)
          T fax = ax, fbx = bx, fay = ay, fby = by;
          if constexpr (sizeof(T) == 8) asm volatile("" : "+d"(fax), "+d"(fbx), "+d"(fay), "+d"(fby));
          const T uax = args.src.p1 * fax, ubx = args.src.p1 * fbx, uay = args.src.p2 * fay, uby = args.src.p2 * fby;
          const int reach = __reduce_max_sync(0xffffffffu, max(max(mag_key(uax), mag_key(ubx)), max(mag_key(uay), mag_key(uby))));
          if (reach < args.key_deg4) {
            sinsin_moments<T, ORDER, 4>(cst, sx, cx, sy, cy, uax, ubx, uay, uby, b0, b1, b2);
          } else if (reach < args.key_deg6) {
            sinsin_moments<T, ORDER, 6>(cst, sx, cx, sy, cy, uax, ubx, uay, uby, b0, b1, b2);
          } else if (reach < args.key_deg8) {
            sinsin_moments<T, ORDER, 8>(cst, sx, cx, sy, cy, uax, ubx, uay, uby, b0, b1, b2);
          } else {  // coarse element: evaluate the source directly
#pragma unroll 1
            for (int q = 0; q < NQV; ++q) {
              const T fx = sin(args.src.p1 * fma(quad.l1[q], ax, fma(quad.l2[q], bx, x0)));
              const T fy = sin(args.src.p2 * fma(quad.l1[q], ay, fma(quad.l2[q], by, y0)));
              const T wf = quad.w[q] * (fx * fy);
              b0 = fma(wf, quad.l0[q], b0);
              b1 = fma(wf, quad.l1[q], b1);
              b2 = fma(wf, quad.l2[q], b2);
            }
          }
        }
      }
    } else if constexpr (RESIDUAL) {
      // r_i = det_f (det sum_q w_q l_i f_q  -  e_i . U),  grad phi_i = e_i / det, e_1 = (by, -bx), e_2 = (-ay, ax)
#pragma unroll
      for (int q = 0; q < NQV; ++q) {
        b0 = fma(cst.wl0[q], fq[q], b0);
        b1 = fma(cst.wl1[q], fq[q], b1);
        b2 = fma(cst.wl2[q], fq[q], b2);
      }
      const T e1u = fma(by, ru0, -(bx * ru1)), e2u = fma(ax, ru1, -(ay * ru0));
      b0 = fma(det, b0, e1u + e2u);
      b1 = fma(det, b1, -e1u);
      b2 = fma(det, b2, -e2u);
    } else if constexpr (SAMPLED) {  // f given at the element's quadrature points
#pragma unroll
      for (int q = 0; q < NQV; ++q) {
        b0 = fma(cst.wl0[q], fq[q], b0);
        b1 = fma(cst.wl1[q], fq[q], b1);
        b2 = fma(cst.wl2[q], fq[q], b2);
      }
    } else {  // constant source: the moments of the basis functions are constants
      b0 = cst.m0; b1 = cst.m1; b2 = cst.m2;
    }
    const T amp = RESIDUAL ? detf : (SAMPLED ? T(1) : args.src.p0) * (FRAC ? det * detf : det);
    b0 *= amp; b1 *= amp; b2 *= amp;
  }
  sts_2(out, k00, b0);
  sts_2(out + 2 * kS, k11, b1);
  sts_2(out + 4 * kS, k22, b2);
  if constexpr (HAS_MAT) {
    sts_1(a_od0 + row * kS, k01);
    sts_1(a_od1 + row * kS, k12);
    sts_1(a_od2 + row * kS, k20);
  }
}

// ---- phase C of one tile: one lane per CSR entry / one thread per owned row, contributions in increasing element id ----
// `inst` = the tile's instance, `s_tc` = the reduction part of its template, `a_tab` = the table phase B filled; executed by
// NTHREADS threads (tid 0 .. NTHREADS-1, whole warps).
template <typename T, int NTHREADS, bool HAS_LOAD, bool HAS_MAT>
__device__ __forceinline__ void reduce_tile(const TiledArgs<T>& args, uint32_t a_tab, const int32_t* inst, const int32_t* s_tc, int n_vert,
                                            int n_rows, int n_segs, int n_chunks, int n_heavy, int n_heavy_contrib, int tid, int warp,
                                            int lane) {
  constexpr int kWarps = NTHREADS / 32;
  const uint32_t a_seg = smem_u32(inst + kInstHeader + pad4(n_vert));  // seg_start[n_segs]
  const uint32_t a_row = a_seg + 4u * pad4(n_segs);                      // row_id[n_rows]
  const uint32_t a_pair = smem_u32(s_tc);
  const uint32_t a_chunk = a_pair + 128u * n_segs;
  const uint32_t a_diag = a_chunk + 16u * n_chunks;
  const uint16_t* heavy_seg = reinterpret_cast<const uint16_t*>(s_tc + 32 * n_segs + 4 * n_chunks + pad4((n_rows + 1) >> 1));
  const uint16_t* heavy_contrib = heavy_seg + 2 * pad4((n_heavy + 2) >> 1);
  const uint16_t* heavy_pos = heavy_contrib + 2 * pad4((n_heavy_contrib + 1) >> 1);
  auto store = [&](uint32_t code, T value) {
    const uint32_t pos = lds_u32(a_seg + 4u * (code >> 5)) + (code & 31u);
    if (!(TFEM_DEBUG_SKIP & 8) || value == T(1.2345e30)) args.csr_val[pos] = value;
  };
  if constexpr (HAS_MAT) {
    // Entries with <= 2 contributions (every off-diagonal of a manifold mesh): one lane per entry,
    // one warp per segment of <= 32 consecutive csr_val slots (coalesced stores).  One word per lane
    // names both contributions; a missing one is code 0 (the zero row), so there is no count and no
    // inner loop.
    T* const csr_lane = args.csr_val + lane;
    uint32_t a_w = a_pair + 128u * warp + 4u * lane, a_s = a_seg + 4u * warp;
    constexpr int kSegUnroll = TFEM_SEG_UNROLL;
#pragma unroll kSegUnroll
    for (int sg = warp; sg < ((TFEM_DEBUG_SKIP & 4) ? 0 : n_segs); sg += kWarps, a_w += 128u * kWarps, a_s += 4u * kWarps) {
      const uint32_t word = lds_u32(a_w);
      const uint32_t start = lds_u32(a_s);
      if (word != 0xffffffffu) {  // lanes without an entry issue no table load (they would only add bank conflicts)
        T first, second;
        lds_1(a_tab + code_offset<T>(word & 0xffffu), first);
        lds_1(a_tab + code_offset<T>(word >> 16), second);
        const T value = first + second;
        if (!(TFEM_DEBUG_SKIP & 8) || value == T(1.2345e30)) csr_lane[start] = value;
      }
    }
    // the few other entries with > 2 contributions (non-manifold edges, degenerate elements)
    for (int h = tid; h < ((TFEM_DEBUG_SKIP & 4) ? 0 : n_heavy); h += NTHREADS) {
      T acc = T(0);
      for (int s = heavy_seg[h]; s < heavy_seg[h + 1]; ++s) {
        T v;
        lds_1(a_tab + code_offset<T>(heavy_contrib[s]), v);
        acc += v;
      }
      store(heavy_pos[h], acc);
    }
  }
  // one thread per owned row: its load entry and its diagonal share one element list, read as
  // chunks of 7 codes + link (one 16 B shared-memory load per chunk); the diagonal and the load term
  // of an element sit side by side in the table (one 16 B load per contribution)
  for (int j = tid; j < ((TFEM_DEBUG_SKIP & 4) ? 0 : n_rows); j += NTHREADS) {
    T rhs = T(0), diag = T(0);
    uint32_t chunk = (uint32_t)j;
    do {
      const uint4 w = lds_v4(a_chunk + 16u * chunk);
      const uint32_t code[7] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16, w.z & 0xffffu, w.z >> 16, w.w & 0xffffu};
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        T d, b;
        lds_2(a_tab + code_offset<T>(code[k]), d, b);
        diag += d;
        rhs += b;
      }
      chunk = w.w >> 16;
    } while (chunk != 0);
    if constexpr (HAS_LOAD) {
      if (!(TFEM_DEBUG_SKIP & 8) || rhs == T(1.2345e30)) args.load[lds_u32(a_row + 4u * j)] = rhs;
    }
    if constexpr (HAS_MAT) {
      const uint32_t code = lds_u16(a_diag + 2u * j);
      if (code != 0xffffu) store(code, diag);
    }
  }
}

// Shared-memory views of a CTA of the tiled kernels.
template <typename T>
struct TileSmem {
  uint64_t *inst_bar, *tb_bar, *tc_bar, *full_bar, *bdone_bar, *done_bar;
  int32_t* s_rec;  // [S][8] record of a staged tile: template generation, n_vert, n_elem, n_rows, n_segs, n_chunks, n_heavy, n_heavy_contrib
  T* sbase;        // [S][8] w1*bx, w2*by, sin/cos(w1 bx), sin/cos(w2 by) of the tile's base vertex
  int32_t *s_inst, *s_tb, *s_tc;
  typename Vec2<T>::type* vxy;  // [S][max_vert]
  int4* s_desc = nullptr;                   // [64] ring of tile descriptors fetched 32 at a time (role-specialised kernel), else NULL
  typename Vec2<T>::type* s_bxy = nullptr;  // [64] the tiles' base-vertex coordinates
};

// the it-th tile of a CTA: its share of the leading (progress-reporting) tiles, dealt round-robin, then one contiguous
// block of the rest
template <typename T>
__device__ __forceinline__ int tile_at(const TiledArgs<T>& args, int n_lead, int cta, int grid, int blk_lo, int it) {
  return __ldg(args.tile_list + (it < n_lead ? cta + it * grid : blk_lo + (it - n_lead)));
}

// =================================== producer warp =======================================
// Stages tile after tile for the consumers, S tiles ahead (KI instance blobs in flight).  Never blocks on data it
// fetched itself: the coordinate gather completes on the tile's `full` barrier asynchronously, the base vertex's
// coordinates are read one tile ahead.
// ROLE 0: everything; the role-specialised kernel splits the work over two warps -- ROLE 1: instance / template copies,
// base point and tile record (never waits for an instance: the consumers then read n_vert from the instance header),
// ROLE 2: the coordinate gather alone (waits for the instance, arrives on `full` like ROLE 1).
template <typename T, bool SINSIN, int S, int KI, int ROLE = 0>
__device__ __forceinline__ void run_producer(const TiledArgs<T>& args, const TileSmem<T>& sm, int lane, int n_local, int n_lead, int cta,
                                             int grid, int blk_lo) {
  if (n_local == 0) return;
  using V2 = typename Vec2<T>::type;
  const V2* coords2 = reinterpret_cast<const V2*>(args.coords);
  if constexpr (ROLE == 2) {
    for (int it = 0; it < n_local; ++it) {
      const int stage = it % S, round = it / S;
      if (it >= S) mbar_wait(sm.done_bar + stage, (round - 1) & 1);  // vxy[stage] is free
      mbar_wait(sm.inst_bar + it % KI, (it / KI) & 1);
      const uint32_t a_vert = smem_u32(sm.s_inst + (it % KI) * args.inst_words);
      const int n_vert = (int)lds_u32(a_vert);
      const uint32_t a_dst = smem_u32(sm.vxy + stage * args.max_vert);
#pragma unroll 4
      for (int i = lane; i < n_vert; i += 32) {
        const uint32_t v = lds_u32(a_vert + 4u * (kInstHeader + i));
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(a_dst + (uint32_t)sizeof(V2) * i), "l"(coords2 + v), "n"((int)sizeof(V2)) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(sm.full_bar + stage)) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.full_bar + stage);  // orders the instance (this warp observed it land) before the consumers
    }
    return;
  }
  auto issue_inst = [&](int it, const int4& d) {
    const int slot = it % KI;
    mbar_expect_tx(sm.inst_bar + slot, (uint32_t)d.y * 4u);
    bulk_g2s(sm.s_inst + slot * args.inst_words, args.inst_blob + d.x, (uint32_t)d.y * 4u, sm.inst_bar + slot);
  };
  constexpr int PF = KI - S;  // instance blobs are fetched PF tiles ahead of the tile being staged
  // tile descriptors and base vertices: three dependent global loads per tile (tile list -> descriptor -> coordinates).
  // With a descriptor ring the 32 lanes fetch them for 32 tiles at once, so their latency is paid once per 32 tiles
  // instead of on every tile's critical path.
  const bool ring = sm.s_desc != nullptr;
  auto refill = [&](int first) {
    const int j = first + lane;
    if (j < n_local) {
      const int4 dj = __ldg(args.tile_desc + tile_at(args, n_lead, cta, grid, blk_lo, j));
      sm.s_desc[j & 63] = dj;
      sm.s_bxy[j & 63] = __ldg(coords2 + dj.w);
    }
  };
  auto desc_of = [&](int j) { return ring ? sm.s_desc[j & 63] : __ldg(args.tile_desc + tile_at(args, n_lead, cta, grid, blk_lo, j)); };
  if (ring) {
    refill(0);
    refill(32);
    __syncwarp();
  }
  int4 d_next = desc_of(0);
  V2 base_next = ring ? sm.s_bxy[0] : __ldg(coords2 + d_next.w);
  if (lane == 0) {
    issue_inst(0, d_next);
    for (int k = 1; k < PF && k < n_local; ++k) issue_inst(k, desc_of(k));
  }
  int cur_tpl = -1, gen = 0;
  int4 head_a = make_int4(0, 0, 0, 0), head_b = make_int4(0, 0, 0, 0);  // TB header of the resident template
  TFEM_T_DECL;
  for (int it = 0; it < n_local; ++it) {
    const int stage = it % S, round = it / S;
    const int4 d = d_next;
    const V2 base = base_next;
    // the coordinates and the instance slot of tile it-3 are free once that tile is done
    if (it >= S) mbar_wait(sm.done_bar + stage, (round - 1) & 1);
    TFEM_T(0);
    if (ring && it > 0 && (it & 31) == 0) {  // the ring holds tiles [it, it + 32): fetch the next 32
      refill(it + 32);
      __syncwarp();
    }
    if (PF > 1 && lane == 0 && it + PF < n_local) issue_inst(it + PF, desc_of(it + PF));
    if (it + 1 < n_local) {
      d_next = desc_of(it + 1);
      if (PF == 1 && lane == 0) issue_inst(it + 1, d_next);
      base_next = ring ? sm.s_bxy[(it + 1) & 63] : __ldg(coords2 + d_next.w);
    }
    int n_vert = 0;
    if constexpr (ROLE == 0) {
      mbar_wait(sm.inst_bar + it % KI, (it / KI) & 1);
      TFEM_T(1);
      const uint32_t a_vert = smem_u32(sm.s_inst + (it % KI) * args.inst_words);
      n_vert = (int)lds_u32(a_vert);
      const uint32_t a_dst = smem_u32(sm.vxy + stage * args.max_vert);
#pragma unroll 4
      for (int i = lane; i < (((TFEM_DEBUG_SKIP & 32) && it >= S) ? 0 : n_vert); i += 32) {  // (32: timing experiment, stale coordinates)
        const uint32_t v = lds_u32(a_vert + 4u * (kInstHeader + i));
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(a_dst + (uint32_t)sizeof(V2) * i), "l"(coords2 + v), "n"((int)sizeof(V2)) : "memory");
      }
      // this lane's copies arrive on `full` when they land
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(sm.full_bar + stage)) : "memory");
    }
    TFEM_T(2);
    if (d.z != cur_tpl) {
      // the single template buffer: TB is free after the integration phase of tile it-1, TC after
      // its reduction phase.  Rare on lattice meshes (a CTA's tiles are congruent); on an
      // unstructured mesh every tile brings its own template and this is the steady state.
      const int4 td = __ldg(args.tpl_desc + d.z);
      if (it >= 1) mbar_wait(sm.bdone_bar + (it - 1) % S, ((it - 1) / S) & 1);
      if (lane == 0) {
        mbar_expect_tx(sm.tb_bar, (uint32_t)td.y * 4u);
        bulk_g2s(sm.s_tb, args.tpl_blob + td.x, (uint32_t)td.y * 4u, sm.tb_bar);
      }
      if (it >= 1) mbar_wait(sm.done_bar + (it - 1) % S, ((it - 1) / S) & 1);
      if (lane == 0) {
        mbar_expect_tx(sm.tc_bar, (uint32_t)td.w * 4u);
        bulk_g2s(sm.s_tc, args.tpl_blob + td.z, (uint32_t)td.w * 4u, sm.tc_bar);
      }
      // the consumers read the template's header from the tile record: wait for TB here (they then need
      // no wait of their own for it: this thread's arrival on `full` orders the TMA writes before them)
      mbar_wait(sm.tb_bar, gen & 1);
      head_a = lds_v4i(smem_u32(sm.s_tb));
      head_b = lds_v4i(smem_u32(sm.s_tb) + 16u);
      ++gen;
      cur_tpl = d.z;
    }
    TFEM_T(3);
    if constexpr (SINSIN) {
      // the tile's only full-range sin/cos, at its base vertex (lane 0: x phase, lane 1: y phase)
      const T pbx = args.src.p1 * base.x, pby = args.src.p2 * base.y;
      if (lane < 2) {
        T s, c;
        if (TFEM_DEBUG_SKIP & 64) { s = T(0.5); c = T(0.8660254037844386); }  // (64: timing experiment, no library sincos)
        else sincos_full(lane == 0 ? pbx : pby, s, c);
        T* sb = sm.sbase + 8 * stage;
        if (lane == 0) {
          sb[0] = pbx; sb[1] = pby; sb[2] = s; sb[3] = c;
        } else {
          sb[4] = s; sb[5] = c;
        }
      }
    }
    if (lane == 0) {
      int4* rec = reinterpret_cast<int4*>(sm.s_rec + 8 * stage);
      rec[0] = make_int4(gen, n_vert, head_a.y, head_a.z);          // generation, n_vert, n_elem, n_rows
      rec[1] = make_int4(head_a.w, head_b.x, head_b.y, head_b.z);  // n_segs, n_chunks, n_heavy, n_heavy_contrib
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(sm.full_bar + stage);
    TFEM_T(4);
  }
#ifdef TFEM_DEBUG_TIMING
  if (lane == 0)
    for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&tfem_timing_table[0][i]), (unsigned long long)t_acc[i]);
#endif
}

template <int CONSUMERS>
struct MinCtas { static constexpr int value = CONSUMERS <= 256 ? 3 : 2; };

template <typename T, int CONSUMERS, int ORDER, int SRC, bool HAS_MAT, bool FRAC>
__global__ void __launch_bounds__(CONSUMERS + 32, MinCtas<CONSUMERS>::value)
assemble_tiled_kernel(const TiledArgs<T> args, const TiledConst<T> cst, const QuadT<T> quad) {
  constexpr int NQV = NQ<ORDER>::value;
  constexpr bool HAS_LOAD = SRC != TFEM_SRC_NONE;
  constexpr bool SINSIN = SRC == TFEM_SRC_SINSIN;
  constexpr bool SAMPLED = SRC == TFEM_SRC_SAMPLED;
  constexpr bool RESIDUAL = SRC == kSrcResidual;  // weak residual: f v - grad v . grad u, both given at the quadrature points
  constexpr bool NEED_IDS = SAMPLED || FRAC || RESIDUAL;
  constexpr int kWarps = CONSUMERS / 32;
  constexpr uint32_t kRowBytes = kDlSlots * sizeof(T);
  constexpr uint32_t kS = sizeof(T);
  using V2 = typename Vec2<T>::type;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [ header | instance x4 | TB | TC | vertex coordinates x3 | table: dl rows [2 + max_elem][6], 3 off-diagonal arrays ]
  uint64_t* inst_bar = reinterpret_cast<uint64_t*>(smem_raw);  // [4] instance landed
  uint64_t* tb_bar = inst_bar + kInstStages;                   // template part TB landed
  uint64_t* tc_bar = tb_bar + 1;                               // template part TC landed
  uint64_t* full_bar = tc_bar + 1;                             // [3] tile staged: 32 coordinate-gather arrivals + the record
  uint64_t* bdone_bar = full_bar + kStages;                    // [3] consumers are through the integration phase
  uint64_t* done_bar = bdone_bar + kStages;                    // [3] consumers finished the tile
  // [3][8] record of a staged tile, written by the producer: template generation, n_vert, n_elem, n_rows,
  // n_segs, n_chunks, n_heavy, n_heavy_contrib -- everything the consumers need, in two 16 B loads
  int32_t* s_rec = reinterpret_cast<int32_t*>(smem_raw + 128);
  T* sbase = reinterpret_cast<T*>(smem_raw + 256);             // [3][8] w1*bx, w2*by, sin/cos(w1 bx), sin/cos(w2 by)
  int32_t* s_inst = reinterpret_cast<int32_t*>(smem_raw + kSmemHeader);
  int32_t* s_tb = s_inst + kInstStages * args.inst_words;
  int32_t* s_tc = s_tb + args.tb_words;
  V2* vxy = reinterpret_cast<V2*>(s_tc + args.tc_words);  // [3][max_vert]
  T* sloc = reinterpret_cast<T*>(vxy + kStages * args.max_vert);  // row 0 = zeros, rows 1.. = tile elements, last row = scratch

  const int tid = threadIdx.x;
  // this CTA's tiles: its share of the leading (progress-reporting) tiles, dealt round-robin, then one
  // contiguous block of the rest
  const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
  const int n_lead = cta < args.n_strided ? (args.n_strided - cta + grid - 1) / grid : 0;
  const int rest = args.n_tiles - args.n_strided;
  const int blk_lo = args.n_strided + (int)(((int64_t)rest * cta) / grid);
  const int blk_hi = args.n_strided + (int)(((int64_t)rest * (cta + 1)) / grid);
  const int n_local = n_lead + (blk_hi - blk_lo);

  if (tid == 0) {
    for (int i = 0; i < kInstStages + 2; ++i) mbar_init(inst_bar + i, 1);  // one expect_tx arrival each
    for (int i = 0; i < kStages; ++i) mbar_init(full_bar + i, 33);
    for (int i = 0; i < kStages; ++i) mbar_init(bdone_bar + i, kWarps);
    for (int i = 0; i < kStages; ++i) mbar_init(done_bar + i, kWarps);
    fence_mbar_init();
  }
  if (tid < kDlSlots) sloc[tid] = T(0);  // row 0 of the table: the "no contribution" target of the packed codes
  __syncthreads();

  if (tid >= CONSUMERS) {
    run_producer<T, SINSIN, kStages, kInstStages>(args, TileSmem<T>{inst_bar, tb_bar, tc_bar, full_bar, bdone_bar, done_bar, s_rec, sbase, s_inst,
                                                               s_tb, s_tc, vxy}, tid - CONSUMERS, n_local, n_lead, cta, grid, blk_lo);
    return;
  }

  // ===================================== consumer warps ========================================
  const int lane = tid & 31, warp = tid >> 5;
#ifdef TFEM_STAGGER_NS
  // experiment: the second CTA to arrive on an SM starts late, so that one CTA integrates (fp64 pipe)
  // while the other reduces (shared-memory pipe)
  {
    __shared__ int s_arrival;
    if (tid == 0) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      s_arrival = atomicAdd(&tfem_stagger_slots[smid], 1);
    }
    consumer_sync<CONSUMERS>();
    if (s_arrival & 1) {
      const uint64_t t0 = global_timer_ns();
      while (global_timer_ns() - t0 < TFEM_STAGGER_NS) __nanosleep(100);
    }
  }
#endif
  const uint32_t a_tab = smem_u32(sloc);
  const uint32_t a_od0 = a_tab + (uint32_t)args.od_base[0], a_od1 = a_tab + (uint32_t)args.od_base[1], a_od2 = a_tab + (uint32_t)args.od_base[2];
  const uint32_t a_elem = smem_u32(s_tb + kTbHeader);
  int seen_tc = 0;
  TFEM_T_DECL;
  for (int it = 0; it < n_local; ++it) {
    const int stage = it % kStages, round = it / kStages;
    mbar_wait(full_bar + stage, round & 1);
    TFEM_T(0);
    // everything about the tile in two 16 B loads (the producer's arrival on `full` also orders the TMA
    // writes of the instance and of a new TB part before this point)
    const int4 rec_a = lds_v4i(smem_u32(s_rec + 8 * stage)), rec_b = lds_v4i(smem_u32(s_rec + 8 * stage) + 16u);
    const int tc_gen = rec_a.x, n_vert = rec_a.y, n_elem = rec_a.z, n_rows = rec_a.w;
    const int n_segs = rec_b.x, n_chunks = rec_b.y, n_heavy = rec_b.z, n_heavy_contrib = rec_b.w;
    const uint32_t a_xy = smem_u32(vxy + stage * args.max_vert);
    // elem_id[n_elem] closes the instance (include/tfem_b200.h)
    const uint32_t a_eid = smem_u32(s_inst + (it % kInstStages) * args.inst_words) + 4u * (uint32_t)(kInstHeader + pad4(n_vert) + pad4(n_segs) + pad4(n_rows));
    TFEM_T(1);

    // ---- B: local matrices and loads, each tile element once ----------------------------------
    const uint32_t a_sb = smem_u32(sbase + 8 * stage);
    for (int base = warp * 32; base < ((TFEM_DEBUG_SKIP & 2) ? 0 : n_elem); base += CONSUMERS) {
      // lanes past the end recompute the last element into a scratch row (no divergence: the votes
      // below need every lane)
      const int el = base + lane;
      const bool valid = el < n_elem;
      const uint32_t slot = (uint32_t)(valid ? el : n_elem - 1);
      const uint32_t row = (uint32_t)(valid ? el + 1 : args.max_elem + 1);
      integrate_tile_element<T, ORDER, SRC, HAS_MAT, FRAC>(args, cst, quad, a_xy, a_sb, a_eid, a_elem, a_tab, a_od0, a_od1, a_od2, slot, row);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bdone_bar + stage);  // the producer may replace the TB part
    consumer_sync<CONSUMERS>();
    TFEM_T(2);

    // ---- C: one lane per CSR entry / one thread per row; contributions in increasing element id --
    if (tc_gen != seen_tc) {  // a new template: its TC part may still be in flight
      mbar_wait(tc_bar, (tc_gen - 1) & 1);
      seen_tc = tc_gen;
    }
    TFEM_T(3);
    reduce_tile<T, CONSUMERS, HAS_LOAD, HAS_MAT>(args, a_tab, s_inst + (it % kInstStages) * args.inst_words, s_tc, n_vert, n_rows, n_segs,
                                                 n_chunks, n_heavy, n_heavy_contrib, tid, warp, lane);
    TFEM_T(5);
    if (args.progress != nullptr && it < n_lead) {
      // the first tiles of the call hold the multi-GPU interface rows: tell the exchange kernels
      // waiting on the counter (tfem_iface_pack_after) that this warp's stores are out
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(args.progress, 1u);
    }
    // the table, the instance slot and vxy[stage] are free once every warp is here
    consumer_sync<CONSUMERS>();
    if (lane == 0) mbar_arrive(done_bar + stage);
    TFEM_T(6);
  }
#ifdef TFEM_DEBUG_TIMING
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&tfem_timing_table[1][i]), (unsigned long long)t_acc[i]);
    atomicAdd(&tfem_timing_ctas, 1);
  }
#endif
}

template <typename T, int CONSUMERS, int ORDER, int SRC, bool HAS_MAT, bool FRAC = false>
int launch_tiled(const tfem_tile_plan* hp, TiledArgs<T>& args, const TiledConst<T>& cst, const QuadT<T>& quad, cudaStream_t s) {
  args.inst_words = pad4(hp->max_inst_words);
  args.tb_words = pad4(hp->max_tb_words);
  args.tc_words = pad4(hp->max_tc_words);
  const size_t smem = kSmemHeader + 4 * ((size_t)kInstStages * args.inst_words + (size_t)args.tb_words + (size_t)args.tc_words) +
                      sizeof(T) * ((size_t)2 * kStages * hp->max_vert) + (size_t)hp->table_bytes / 8 * sizeof(T);
  if (smem > 227 * 1024) return TFEM_ERR_TOO_LARGE;
  auto kern = assemble_tiled_kernel<T, CONSUMERS, ORDER, SRC, HAS_MAT, FRAC>;
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
  kern<<<grid, CONSUMERS + 32, smem, s>>>(args, cst, quad);
  return check_launch();
}

// ================================================================================================
// Role-specialised variant: ONE CTA per SM whose consumer warps are split into an INTEGRATION group
// (NB warps: phase B only, the fp64 pipe) and a REDUCTION group (NC warps: phase C only, the
// shared-memory / store pipes) working one tile apart on two tables.  The two phases of consecutive
// tiles therefore always overlap on an SM -- in the kernel above two co-resident CTAs fall into step
// and time-share first the fp64 pipe, then the shared-memory pipe -- and there is no block barrier:
//
//   producer (1 warp)        integration warps (NB)                 reduction warps (NC)
//   stage tile t+1..t+3      wait full[t], wait done[t-2]           wait bdone[t]
//   (as above)               B: chunks of 32 elements, dealt        C: entries, heavy entries, rows of
//                            round-robin ACROSS tiles (no           table[t & 1] -> csr_val / load
//                            quantisation loss per tile)            arrive done[t]
//                            -> table[t & 1]; arrive bdone[t]
// ================================================================================================
constexpr int kWsStages = 4;      // tiles staged ahead of the reduction group (it runs one tile behind the integration group)
constexpr int kWsInstStages = TFEM_WS_INST_STAGES;
constexpr int kWsHeader = 2688;   // <= 21 mbarriers | [4][8] records at 192 | [4][8] base-point values at 320 | descriptor ring at 640 | base vertices at 1664

template <typename T, int NB, int NC, int ORDER, int SRC, bool HAS_MAT, bool FRAC>
__global__ void __launch_bounds__(32 * (NB + NC + 2), 1)
assemble_tiled_ws_kernel(const TiledArgs<T> args, const TiledConst<T> cst, const QuadT<T> quad) {
  constexpr bool HAS_LOAD = SRC != TFEM_SRC_NONE;
  constexpr bool SINSIN = SRC == TFEM_SRC_SINSIN;
  constexpr bool NEED_IDS = SRC == TFEM_SRC_SAMPLED || FRAC || SRC == kSrcResidual;
  constexpr int S = kWsStages, KI = kWsInstStages;
  using V2 = typename Vec2<T>::type;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [ header | instance x5 | TB | TC | vertex coordinates x4 | table x2 ]
  TileSmem<T> sm;
  sm.inst_bar = reinterpret_cast<uint64_t*>(smem_raw);
  sm.tb_bar = sm.inst_bar + KI;
  sm.tc_bar = sm.tb_bar + 1;
  sm.full_bar = sm.tc_bar + 1;
  sm.bdone_bar = sm.full_bar + S;
  sm.done_bar = sm.bdone_bar + S;
  sm.s_rec = reinterpret_cast<int32_t*>(smem_raw + 192);
  sm.sbase = reinterpret_cast<T*>(smem_raw + 320);
  sm.s_desc = reinterpret_cast<int4*>(smem_raw + 640);
  sm.s_bxy = reinterpret_cast<V2*>(smem_raw + 1664);
  sm.s_inst = reinterpret_cast<int32_t*>(smem_raw + kWsHeader);
  sm.s_tb = sm.s_inst + KI * args.inst_words;
  sm.s_tc = sm.s_tb + args.tb_words;
  sm.vxy = reinterpret_cast<V2*>(sm.s_tc + args.tc_words);
  unsigned char* tables = reinterpret_cast<unsigned char*>(sm.vxy + S * args.max_vert);

  const int tid = threadIdx.x;
  const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
  const int n_lead = cta < args.n_strided ? (args.n_strided - cta + grid - 1) / grid : 0;
  const int rest = args.n_tiles - args.n_strided;
  const int blk_lo = args.n_strided + (int)(((int64_t)rest * cta) / grid);
  const int blk_hi = args.n_strided + (int)(((int64_t)rest * (cta + 1)) / grid);
  const int n_local = n_lead + (blk_hi - blk_lo);

  if (tid == 0) {
    for (int i = 0; i < KI + 2; ++i) mbar_init(sm.inst_bar + i, 1);  // one expect_tx arrival each
    for (int i = 0; i < S; ++i) mbar_init(sm.full_bar + i, 34);  // 32 coordinate-gather completions + one arrival per producer warp
    for (int i = 0; i < S; ++i) mbar_init(sm.bdone_bar + i, NB);
    for (int i = 0; i < S; ++i) mbar_init(sm.done_bar + i, NC);
    fence_mbar_init();
  }
  if (tid < 2 * kDlSlots)  // row 0 of both tables: the "no contribution" target of the packed codes
    reinterpret_cast<T*>(tables + (tid / kDlSlots) * args.table_stride)[tid % kDlSlots] = T(0);
  __syncthreads();

  const int lane = tid & 31, warp = tid >> 5;
  if (warp >= NB + NC) {
    if (warp == NB + NC) run_producer<T, SINSIN, S, KI, 1>(args, sm, lane, n_local, n_lead, cta, grid, blk_lo);
    else run_producer<T, SINSIN, S, KI, 2>(args, sm, lane, n_local, n_lead, cta, grid, blk_lo);
    return;
  }
  const uint32_t a_tables = smem_u32(tables);

  if (warp < NB) {
    // =============================== integration warps: phase B ===============================
    const uint32_t a_elem = smem_u32(sm.s_tb + kTbHeader);
    int chunk_base = 0;  // chunks dealt so far, modulo NB: chunk c of this tile goes to warp (chunk_base + c) % NB
    TFEM_T_DECL;
    for (int it = 0; it < n_local; ++it) {
      const int stage = it % S, round = it / S;
      mbar_wait(sm.full_bar + stage, round & 1);
      TFEM_T(0);
      const int4 rec_a = lds_v4i(smem_u32(sm.s_rec + 8 * stage));
      const int n_elem = rec_a.z;
      const uint32_t a_xy = smem_u32(sm.vxy + stage * args.max_vert);
      const uint32_t a_sb = smem_u32(sm.sbase + 8 * stage);
      uint32_t a_eid = 0;
      if constexpr (NEED_IDS) {  // elem_id[n_elem] closes the instance (include/tfem_b200.h)
        const int4 rec_b = lds_v4i(smem_u32(sm.s_rec + 8 * stage) + 16u);
        const uint32_t a_inst = smem_u32(sm.s_inst + (it % KI) * args.inst_words);
        a_eid = a_inst + 4u * (uint32_t)(kInstHeader + pad4((int)lds_u32(a_inst)) + pad4(rec_b.x) + pad4(rec_a.w));
      }
      // this tile's table was read by the reduction group two tiles ago
      if (it >= 2) mbar_wait(sm.done_bar + (it - 2) % S, ((it - 2) / S) & 1);
      TFEM_T(1);
      const uint32_t a_tab = a_tables + (uint32_t)(it & 1) * (uint32_t)args.table_stride;
      const uint32_t a_od0 = a_tab + (uint32_t)args.od_base[0], a_od1 = a_tab + (uint32_t)args.od_base[1], a_od2 = a_tab + (uint32_t)args.od_base[2];
      const int n_chunks = (n_elem + 31) >> 5;
      int first = warp - chunk_base;
      if (first < 0) first += NB;
      for (int c = first; c < ((TFEM_DEBUG_SKIP & 2) ? 0 : n_chunks); c += NB) {
        // lanes past the end recompute the last element into a scratch row (no divergence: the votes of the
        // source evaluation need every lane)
        const int el = c * 32 + lane;
        const bool valid = el < n_elem;
        const uint32_t slot = (uint32_t)(valid ? el : n_elem - 1);
        const uint32_t row = (uint32_t)(valid ? el + 1 : args.max_elem + 1);
        integrate_tile_element<T, ORDER, SRC, HAS_MAT, FRAC>(args, cst, quad, a_xy, a_sb, a_eid, a_elem, a_tab, a_od0, a_od1, a_od2, slot, row);
      }
      chunk_base = (chunk_base + n_chunks) % NB;
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.bdone_bar + stage);  // (release: this warp's table rows are visible to whoever acquires the phase)
      TFEM_T(2);
    }
#ifdef TFEM_DEBUG_TIMING
    if (tid == 0) {
      for (int i = 0; i < 3; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&tfem_timing_table[1][i]), (unsigned long long)t_acc[i]);
      atomicAdd(&tfem_timing_ctas, 1);
    }
#endif
    return;
  }

  // ================================= reduction warps: phase C =================================
  const int ctid = tid - 32 * NB, cwarp = warp - NB;
  int seen_tc = 0;
  TFEM_T_DECL;
  for (int it = 0; it < n_local; ++it) {
    const int stage = it % S, round = it / S;
    mbar_wait(sm.bdone_bar + stage, round & 1);
    TFEM_T(4);
    const int4 rec_a = lds_v4i(smem_u32(sm.s_rec + 8 * stage)), rec_b = lds_v4i(smem_u32(sm.s_rec + 8 * stage) + 16u);
    const int tc_gen = rec_a.x, n_rows = rec_a.w;
    const int n_vert = (int)lds_u32(smem_u32(sm.s_inst + (it % KI) * args.inst_words));  // (the record's copy is not filled in here)
    const int n_segs = rec_b.x, n_chunks = rec_b.y, n_heavy = rec_b.z, n_heavy_contrib = rec_b.w;
    if (tc_gen != seen_tc) {  // a new template: its TC part may still be in flight
      mbar_wait(sm.tc_bar, (tc_gen - 1) & 1);
      seen_tc = tc_gen;
    }
    const uint32_t a_tab = a_tables + (uint32_t)(it & 1) * (uint32_t)args.table_stride;
    reduce_tile<T, 32 * NC, HAS_LOAD, HAS_MAT>(args, a_tab, sm.s_inst + (it % KI) * args.inst_words, sm.s_tc, n_vert, n_rows, n_segs, n_chunks,
                                               n_heavy, n_heavy_contrib, ctid, cwarp, lane);
    if (args.progress != nullptr && it < n_lead) {
      // the first tiles of the call hold the multi-GPU interface rows: tell the exchange kernels
      // waiting on the counter (tfem_iface_pack_after) that this warp's stores are out
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(args.progress, 1u);
    }
    // the table, the instance slot and vxy[stage] are free once every reduction warp is here
    __syncwarp();
    if (lane == 0) mbar_arrive(sm.done_bar + stage);
    TFEM_T(5);
  }
#ifdef TFEM_DEBUG_TIMING
  if (ctid == 0)
    for (int i = 4; i < 6; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&tfem_timing_table[1][i]), (unsigned long long)t_acc[i]);
#endif
}

template <typename T, int NB, int NC, int ORDER, int SRC, bool HAS_MAT, bool FRAC = false>
int launch_tiled_ws(const tfem_tile_plan* hp, TiledArgs<T>& args, const TiledConst<T>& cst, const QuadT<T>& quad, cudaStream_t s) {
  args.inst_words = pad4(hp->max_inst_words);
  args.tb_words = pad4(hp->max_tb_words);
  args.tc_words = pad4(hp->max_tc_words);
  args.table_stride = (int)((size_t)hp->table_bytes / 8 * sizeof(T));
  const size_t smem = kWsHeader + 4 * ((size_t)kWsInstStages * args.inst_words + (size_t)args.tb_words + (size_t)args.tc_words) +
                      sizeof(T) * ((size_t)2 * kWsStages * hp->max_vert) + (size_t)2 * args.table_stride;
  if (smem > 227 * 1024) return TFEM_ERR_TOO_LARGE;
  auto kern = assemble_tiled_ws_kernel<T, NB, NC, ORDER, SRC, HAS_MAT, FRAC>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TFEM_ERR_LAUNCH;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
    return TFEM_ERR_LAUNCH;
  int64_t resident = sms;  // persistent grid, one CTA per SM
  // room for concurrent kernels: reserve_ctas counts CTA slots of the classic kernel (two per SM)
  if (hp->reserve_ctas > 0 && resident > (hp->reserve_ctas + 1) / 2) resident -= (hp->reserve_ctas + 1) / 2;
  const unsigned grid = (unsigned)(hp->n_tiles < resident ? hp->n_tiles : resident);
  kern<<<grid, 32 * (NB + NC + 2), smem, s>>>(args, cst, quad);
  return check_launch();
}

// how a (source kind, matrix?, fracture?) instantiation is launched: the classic kernel with CONSUMERS threads, or the
// role-specialised one with NB integration and NC reduction warps
template <typename T, int CONSUMERS, int ORDER>
struct LaunchClassic {
  template <int SRC, bool HAS_MAT, bool FRAC = false>
  static int run(const tfem_tile_plan* hp, TiledArgs<T>& args, const TiledConst<T>& cst, const QuadT<T>& quad, cudaStream_t s) {
    return launch_tiled<T, CONSUMERS, ORDER, SRC, HAS_MAT, FRAC>(hp, args, cst, quad, s);
  }
};
template <typename T, int NB, int NC, int ORDER>
struct LaunchWs {
  template <int SRC, bool HAS_MAT, bool FRAC = false>
  static int run(const tfem_tile_plan* hp, TiledArgs<T>& args, const TiledConst<T>& cst, const QuadT<T>& quad, cudaStream_t s) {
    return launch_tiled_ws<T, NB, NC, ORDER, SRC, HAS_MAT, FRAC>(hp, args, cst, quad, s);
  }
};

template <typename T, int ORDER, typename L>
int dispatch_tiled(const tfem_tile_plan* hp, TiledArgs<T>& args, const TiledConst<T>& cst, const QuadT<T>& quad, cudaStream_t s) {
  const int kind = args.load ? args.src.kind : TFEM_SRC_NONE;
#ifdef TFEM_FAST_BUILD  // experiment builds: only the headline instantiation (fp64, 4-point rule, K+M and sin-sin load)
  if constexpr (sizeof(T) == 8 && ORDER == 3) {
    if (args.csr_val && kind == TFEM_SRC_SINSIN) return L::template run<TFEM_SRC_SINSIN, true>(hp, args, cst, quad, s);
  }
  return TFEM_ERR_UNSUPPORTED;
#else
  if (args.frac_metric && kind != kSrcResidual) {  // fracture network: tangential forms; analytic sources live in 3-D and come sampled
    if (args.csr_val) {
      if (kind == TFEM_SRC_SAMPLED) return L::template run<TFEM_SRC_SAMPLED, true, true>(hp, args, cst, quad, s);
      if (kind == TFEM_SRC_CONST) return L::template run<TFEM_SRC_CONST, true, true>(hp, args, cst, quad, s);
      if (kind == TFEM_SRC_NONE) return L::template run<TFEM_SRC_NONE, true, true>(hp, args, cst, quad, s);
      return TFEM_ERR_UNSUPPORTED;
    }
    if (kind == TFEM_SRC_SAMPLED) return L::template run<TFEM_SRC_SAMPLED, false, true>(hp, args, cst, quad, s);
    if (kind == TFEM_SRC_CONST) return L::template run<TFEM_SRC_CONST, false, true>(hp, args, cst, quad, s);
    return TFEM_ERR_UNSUPPORTED;
  }
  if (kind == kSrcResidual) {
    if (args.frac_metric) return L::template run<kSrcResidual, false, true>(hp, args, cst, quad, s);
    return L::template run<kSrcResidual, false, false>(hp, args, cst, quad, s);
  }
  if (kind == TFEM_SRC_SAMPLED) {
    if (args.csr_val) return L::template run<TFEM_SRC_SAMPLED, true>(hp, args, cst, quad, s);
    return L::template run<TFEM_SRC_SAMPLED, false>(hp, args, cst, quad, s);
  }
  if (args.csr_val) {
    if (kind == TFEM_SRC_SINSIN) return L::template run<TFEM_SRC_SINSIN, true>(hp, args, cst, quad, s);
    if (kind == TFEM_SRC_CONST) return L::template run<TFEM_SRC_CONST, true>(hp, args, cst, quad, s);
    return L::template run<TFEM_SRC_NONE, true>(hp, args, cst, quad, s);
  }
  if (kind == TFEM_SRC_SINSIN) return L::template run<TFEM_SRC_SINSIN, false>(hp, args, cst, quad, s);
  if (kind == TFEM_SRC_CONST) return L::template run<TFEM_SRC_CONST, false>(hp, args, cst, quad, s);
  return TFEM_ERR_BAD_ARG;
#endif
}

// consumer_threads of the plan selects the kernel: 0 = default (role-specialised, 12 + 12 warps), 384 = the classic kernel
// with that many consumer threads (two CTAs per SM), 100 * NB + NC >= 1000 = the role-specialised kernel (one CTA per SM)
#ifdef TFEM_FAST_BUILD
#define TFEM_WS_CONFIGS(X) X(16, 8) X(12, 12) X(14, 10) X(14, 12) X(16, 10) X(16, 12) X(12, 14) X(10, 14)
#else
#define TFEM_WS_CONFIGS(X) X(12, 12)
#endif
constexpr int kDefaultConsumers = 1212;

template <typename T, typename F>
int with_order(int quad_order, F&& f) {
  switch (quad_order) {
    case 1: return f(std::integral_constant<int, 1>{});
    case 2: return f(std::integral_constant<int, 2>{});
    case 3: return f(std::integral_constant<int, 3>{});
    default: return f(std::integral_constant<int, 4>{});
  }
}

template <typename T>
int dispatch_tiled_order(int consumers, int quad_order, const tfem_tile_plan* hp, TiledArgs<T>& args, const TiledConst<T>& cst,
                         const QuadT<T>& quad, cudaStream_t s) {
  if (consumers == 0) {
    // default: the role-specialised kernel; plans whose tiles do not fit its two tables and deeper staging in shared
    // memory (scattered numberings: large instances and templates) run on the classic kernel -- both report progress
    // with 12 warps per tile
    const int status = dispatch_tiled_order<T>(kDefaultConsumers, quad_order, hp, args, cst, quad, s);
    if (status != TFEM_ERR_TOO_LARGE) return status;
    consumers = 384;
  }
#define TFEM_WS_CASE(NB, NC)                                                                                               \
  if (consumers == 100 * NB + NC)                                                                                          \
    return with_order<T>(quad_order, [&](auto o) { return dispatch_tiled<T, decltype(o)::value, LaunchWs<T, NB, NC, decltype(o)::value>>(hp, args, cst, quad, s); });
  TFEM_WS_CONFIGS(TFEM_WS_CASE)
#undef TFEM_WS_CASE
  if (consumers == 384)
    return with_order<T>(quad_order, [&](auto o) { return dispatch_tiled<T, decltype(o)::value, LaunchClassic<T, 384, decltype(o)::value>>(hp, args, cst, quad, s); });
  return TFEM_ERR_BAD_ARG;
}

template <typename T>
int assemble_tiled(const tfem_tile_plan* hp, const T* coords, int quad_order, const tfem_bilinear* form,
                   const tfem_source* source, const T* f_q, int64_t n_el_per_mesh, const T* frac_metric, T* csr_val, T* load,
                   void* stream) {
  if (!hp || hp->n_tiles < 0) return TFEM_ERR_BAD_ARG;
  if (hp->n_tiles == 0) return TFEM_OK;
  if (!coords || (!csr_val && !load)) return TFEM_ERR_BAD_ARG;
  if (csr_val && !form) return TFEM_ERR_BAD_ARG;
  if (!hp->tile_list || !hp->tile_desc || !hp->inst_blob || !hp->tpl_desc || !hp->tpl_blob) return TFEM_ERR_BAD_ARG;
  if (hp->max_vert > 1024 || hp->max_elem > 880 || hp->max_vert < 0 || hp->max_elem < 0) return TFEM_ERR_TOO_LARGE;
  if (hp->table_bytes > 65536 || hp->table_bytes % 16 != 0) return TFEM_ERR_BAD_ARG;
  for (int k = 0; k < 3; ++k)
    if (hp->od_base[k] < 48 * (hp->max_elem + 2) || hp->od_base[k] % 8 != 0 || hp->od_base[k] + 8 * (hp->max_elem + 2) > hp->table_bytes)
      return TFEM_ERR_BAD_ARG;
  if (hp->n_tiles > kMaxIndex) return TFEM_ERR_TOO_LARGE;
  if (tri_n_q(quad_order) == 0) return TFEM_ERR_UNSUPPORTED;
  TiledArgs<T> args{};
  args.src = make_source<T>(source);
  if (load && (args.src.kind < TFEM_SRC_NONE || args.src.kind > TFEM_SRC_SINSIN)) return TFEM_ERR_BAD_ARG;
  const bool sampled = load && args.src.kind == TFEM_SRC_SAMPLED;
  if (sampled && !f_q) return TFEM_ERR_BAD_ARG;
  if ((sampled || frac_metric) && !hp->has_elem_ids) return TFEM_ERR_BAD_ARG;  // the plan must carry element ids
  if (frac_metric && n_el_per_mesh <= 0) return TFEM_ERR_BAD_ARG;
  args.f_q = f_q;
  args.frac_metric = frac_metric;
  args.n_el_per_mesh = (int)(n_el_per_mesh > 0 ? n_el_per_mesh : 1);
  if (load && args.src.kind == TFEM_SRC_NONE) {  // f == 0: still write every row of the load vector
    args.src.kind = TFEM_SRC_CONST;
    args.src.p0 = T(0);
  }
  const QuadT<T> quad = make_quad<T>(quad_order);
  args.n_tiles = (int)hp->n_tiles;
  const bool reports = hp->progress != nullptr && hp->n_progress_tiles > 0;
  args.n_strided = reports ? (hp->n_progress_tiles < args.n_tiles ? hp->n_progress_tiles : args.n_tiles) : 0;
  args.progress = reports ? hp->progress : nullptr;
  args.tile_list = hp->tile_list;
  args.tile_desc = reinterpret_cast<const int4*>(hp->tile_desc);
  args.inst_blob = hp->inst_blob;
  args.tpl_desc = reinterpret_cast<const int4*>(hp->tpl_desc);
  args.tpl_blob = hp->tpl_blob;
  args.max_vert = hp->max_vert;
  args.max_elem = hp->max_elem;
  for (int k = 0; k < 3; ++k) args.od_base[k] = hp->od_base[k] / 8 * (int)sizeof(T);
  args.coords = coords;
  const TiledConst<T> cst = make_tiled_const<T>(quad_order, quad, form ? T(form->alpha) : T(0), form ? T(form->beta) : T(0), args.src);
  args.csr_val = csr_val;
  args.load = load;
  // accuracy thresholds of the source evaluation (see sincos_small / sincos_medium / sinsin_moments)
  const double spread = centroid_spread(quad_order);
  const double deg4_reach = 4.0e-3 / spread;  // |phase about the centroid| < 4e-3: t^5/120 < 1e-14
  double w_max = args.src.p1 < 0 ? -(double)args.src.p1 : (double)args.src.p1;
  const double w2_abs = args.src.p2 < 0 ? -(double)args.src.p2 : (double)args.src.p2;
  w_max = w2_abs > w_max ? w2_abs : w_max;
  args.key_rot = host_mag_key(0.04, T(0));
  args.key_rot_medium = host_mag_key(0.2, T(0));
  // fast-path test on squared edge lengths: frequency * edge length < deg4_reach
  args.key_len2 = host_mag_key(w_max > 0 ? (deg4_reach / w_max) * (deg4_reach / w_max) : 1e300, T(0));
  args.key_deg4 = host_mag_key(deg4_reach, T(0));
  args.key_deg6 = host_mag_key(3.0e-2 / spread, T(0));  // t^7/5040 < 5e-15
  args.key_deg8 = host_mag_key(1.0e-1 / spread, T(0));  // t^9/362880 < 3e-15
  auto s = static_cast<cudaStream_t>(stream);
  return dispatch_tiled_order<T>(hp->consumer_threads, quad_order, hp, args, cst, quad, s);
}

}  // namespace tfem

namespace tfem {
// Weak residual r = sum over elements of  sum_q dx (f phi_i - grad phi_i . grad u)  straight to the DOF vector in ONE launch
// of the tiled kernel (load-vector path: rows sum their elements' terms in increasing element order).
template <typename T>
int weak_residual_tiled(const tfem_tile_plan* hp, const T* coords, int quad_order, const T* f_q, const T* grad_u,
                        int64_t n_el_per_mesh, const T* frac_inv, const T* frac_metric, T* r, void* stream) {
  if (!hp || hp->n_tiles < 0) return TFEM_ERR_BAD_ARG;
  if (hp->n_tiles == 0) return TFEM_OK;
  if (!coords || !grad_u || !r || !hp->has_elem_ids) return TFEM_ERR_BAD_ARG;
  if ((frac_inv == nullptr) != (frac_metric == nullptr)) return TFEM_ERR_BAD_ARG;
  if (frac_inv && n_el_per_mesh <= 0) return TFEM_ERR_BAD_ARG;
  if (!hp->tile_list || !hp->tile_desc || !hp->inst_blob || !hp->tpl_desc || !hp->tpl_blob) return TFEM_ERR_BAD_ARG;
  if (hp->max_vert > 1024 || hp->max_elem > 880 || hp->table_bytes > 65536 || hp->table_bytes % 16 != 0) return TFEM_ERR_TOO_LARGE;
  if (tri_n_q(quad_order) == 0) return TFEM_ERR_UNSUPPORTED;
  TiledArgs<T> args{};
  args.src.kind = kSrcResidual;
  args.src.p0 = T(1);
  const QuadT<T> quad = make_quad<T>(quad_order);
  args.n_tiles = (int)hp->n_tiles;
  args.n_strided = 0;
  args.progress = nullptr;
  args.tile_list = hp->tile_list;
  args.tile_desc = reinterpret_cast<const int4*>(hp->tile_desc);
  args.inst_blob = hp->inst_blob;
  args.tpl_desc = reinterpret_cast<const int4*>(hp->tpl_desc);
  args.tpl_blob = hp->tpl_blob;
  args.max_vert = hp->max_vert;
  args.max_elem = hp->max_elem;
  for (int k = 0; k < 3; ++k) args.od_base[k] = hp->od_base[k] / 8 * (int)sizeof(T);
  args.coords = coords;
  args.f_q = f_q;
  args.grad_u = grad_u;
  args.frac_inv = frac_inv;
  args.frac_metric = frac_metric;
  args.n_el_per_mesh = (int)(n_el_per_mesh > 0 ? n_el_per_mesh : 1);
  const TiledConst<T> cst = make_tiled_const<T>(quad_order, quad, T(0), T(0), args.src);
  args.csr_val = nullptr;
  args.load = r;
  auto s = static_cast<cudaStream_t>(stream);
  // the residual's integration phase streams grad u from global memory (latency-bound, not fp64-bound): the classic
  // kernel's 24 integrating warps per SM keep more loads in flight than the 12 of the role-specialised one (C4 forward:
  // 89 us against 145 us)
  return dispatch_tiled_order<T>(hp->consumer_threads ? hp->consumer_threads : 384, quad_order, hp, args, cst, quad, s);
}
}  // namespace tfem

#define TFEM_TILED_API(T, SUF)                                                                                        \
  extern "C" int tfem_weak_residual_tiled_##SUF(const tfem_tile_plan* host_plan, const T* coords, int quad_order,     \
                                                const T* f_q, const T* grad_u, int64_t n_el_per_mesh,                 \
                                                const T* frac_inv, const T* frac_metric, T* r, void* stream) {        \
    return tfem::weak_residual_tiled<T>(host_plan, coords, quad_order, f_q, grad_u, n_el_per_mesh, frac_inv,          \
                                        frac_metric, r, stream);                                                      \
  }                                                                                                                   \
  extern "C" int tfem_tri_p1_assemble_csr_##SUF(const tfem_tile_plan* host_plan, const T* coords, int quad_order,     \
                                                const tfem_bilinear* host_form, const tfem_source* host_source,       \
                                                T* csr_val, T* load, void* stream) {                                  \
    return tfem::assemble_tiled<T>(host_plan, coords, quad_order, host_form, host_source, nullptr, 0, nullptr,        \
                                   csr_val, load, stream);                                                            \
  }                                                                                                                   \
  extern "C" int tfem_tri_p1_assemble_csr_ex_##SUF(const tfem_tile_plan* host_plan, const T* coords, int quad_order,  \
                                                   const tfem_bilinear* host_form, const tfem_source* host_source,    \
                                                   const T* f_q, int64_t n_el_per_mesh, const T* frac_metric,         \
                                                   T* csr_val, T* load, void* stream) {                               \
    return tfem::assemble_tiled<T>(host_plan, coords, quad_order, host_form, host_source, f_q, n_el_per_mesh,         \
                                   frac_metric, csr_val, load, stream);                                               \
  }
TFEM_TILED_API(double, f64)
TFEM_TILED_API(float, f32)

#ifdef TFEM_DEBUG_TIMING
// profiling builds only: cycles per phase summed over all CTAs, [producer | consumer thread 0][8], then zeroed
extern "C" int tfem_debug_timing_read(long long* host_out, int* host_ctas) {
  long long zero[2][8] = {};
  int zero_ctas = 0;
  if (cudaMemcpyFromSymbol(host_out, tfem_timing_table, sizeof(zero)) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(host_ctas, tfem_timing_ctas, sizeof(int)) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(tfem_timing_table, zero, sizeof(zero));
  cudaMemcpyToSymbol(tfem_timing_ctas, &zero_ctas, sizeof(int));
  return 0;
}
#endif
