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
#include "common.cuh"

// Profiling aid (never set in the shipped build): -DTFEM_DEBUG_SKIP=<mask> removes phases so their
// share of the kernel time can be measured; 2 = B (integration), 4 = C (reduction),
// 8 = keep C's arithmetic but suppress its global stores, 16 = B without the load vector's source.
#ifndef TFEM_DEBUG_SKIP
#define TFEM_DEBUG_SKIP 0
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
constexpr int kInstHeader = 4;
constexpr int kTbHeader = 8;
constexpr int kInstStages = 3;   // instance blobs in flight
constexpr int kSlots = 9;        // K00 K11 K22 K01 K12 K20 b0 b1 b2
constexpr int kSmemHeader = 384; // mbarriers, two stage records, two base-point records

template <int ORDER> struct NQ;
template <> struct NQ<1> { static constexpr int value = 1; };
template <> struct NQ<2> { static constexpr int value = 3; };
template <> struct NQ<3> { static constexpr int value = 4; };
template <> struct NQ<4> { static constexpr int value = 6; };

__device__ __forceinline__ void sincos_full(double x, double& s, double& c) { sincos(x, &s, &c); }
__device__ __forceinline__ void sincos_full(float x, float& s, float& c) { sincosf(x, &s, &c); }

// Magnitude keys: order-preserving integer images of |v| (the high word for doubles), so that the
// largest of several magnitudes and a threshold test cost integer max / one warp reduction.
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

// 1/d without the range checks of the library division: MUFU seed, one cubic and one quadratic
// Newton step (relative error ~1e-16; inf / NaN for d == 0 like the reference's division).
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  e = fma(e, e, e);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ float fast_rcp(float d) { return __frcp_rn(d); }

// sin(t), cos(t) for |t| <= 0.04 (truncation < 2e-16 relative to 1)
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

// sin(t), cos(t) for |t| <= 0.2 (truncation < 6e-16 relative to 1)
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

// sin(a + t) as a polynomial in t with coefficients k[] = {sin a, cos a, -sin a/2, -cos a/6, ...}
template <typename T, int DEG>
__device__ __forceinline__ void shifted_sine_coefficients(T s, T c, T (&k)[DEG + 1]) {
  constexpr double inv_fact[9] = {1.0, 1.0, -1.0 / 2.0, -1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, -1.0 / 720.0, -1.0 / 5040.0, 1.0 / 40320.0};
#pragma unroll
  for (int d = 0; d <= DEG; ++d) k[d] = d < 2 ? (d == 0 ? s : c) : T(inv_fact[d]) * ((d & 1) ? c : s);
}

// m_i = sum_q (w_q l_i(q)) sin(Xc + tx_q) sin(Yc + ty_q): the source about the element's CENTROID, whose
// sin/cos (sx, cx, sy, cy) are given; u** = frequency * edge vector.  Horner of degree DEG per point; a
// point AT the centroid (the first point of the 4-point rule) needs no polynomial at all.
template <typename T, int ORDER, int DEG>
__device__ __forceinline__ void sinsin_moments(T sx, T cx, T sy, T cy, T uax, T ubx, T uay, T uby, T& m0, T& m1, T& m2) {
  const TriTable tt = tri_table(ORDER);  // folded at compile time
  T kx[DEG + 1], ky[DEG + 1];
  shifted_sine_coefficients<T, DEG>(sx, cx, kx);
  shifted_sine_coefficients<T, DEG>(sy, cy, ky);
  m0 = m1 = m2 = T(0);
#pragma unroll
  for (int q = 0; q < NQ<ORDER>::value; ++q) {
    const double l1 = tt.xi[q], l2 = tt.eta[q], l0 = 1.0 - l1 - l2, w = 0.5 * tt.w[q];
    const double c1 = l1 - 1.0 / 3.0, c2 = l2 - 1.0 / 3.0;  // barycentric offset from the centroid
    T f;
    if (c1 * c1 + c2 * c2 < 1e-24) {
      f = sx * sy;
    } else {
      const T tx = fma(T(c1), uax, T(c2) * ubx), ty = fma(T(c1), uay, T(c2) * uby);
      T px = kx[DEG], py = ky[DEG];
#pragma unroll
      for (int d = DEG - 1; d >= 0; --d) {
        px = fma(px, tx, kx[d]);
        py = fma(py, ty, ky[d]);
      }
      f = px * py;
    }
    m0 = fma(T(w * l0), f, m0);
    m1 = fma(T(w * l1), f, m1);
    m2 = fma(T(w * l2), f, m2);
  }
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
  int max_vert, inst_words, tb_words, tc_words;
  uint32_t* progress;
  const T* coords;
  T alpha, beta;
  SourceT<T> src;
  T* csr_val;
  T* load;
  int key_rot_small, key_rot_medium;  // |phase base -> centroid| below: sincos_small / sincos_medium, else library
  int key_deg4, key_deg6, key_deg8;   // element reach below: expansion degree 4 / 6 / 8, else sin() per point
};

template <int CONSUMERS>
struct MinCtas { static constexpr int value = CONSUMERS <= 256 ? 3 : 2; };

template <typename T, int CONSUMERS, int ORDER, int SRC, bool HAS_MAT>
__global__ void __launch_bounds__(CONSUMERS + 32, MinCtas<CONSUMERS>::value)
assemble_tiled_kernel(const TiledArgs<T> args, const QuadT<T> quad) {
  constexpr int NQV = NQ<ORDER>::value;
  constexpr bool HAS_LOAD = SRC != TFEM_SRC_NONE;
  constexpr bool SINSIN = SRC == TFEM_SRC_SINSIN;
  constexpr int kWarps = CONSUMERS / 32;
  using V2 = typename Vec2<T>::type;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [ header | instance x3 | TB | TC | vertex coordinates x2 | table[1 + max_elem][9] ]
  uint64_t* inst_bar = reinterpret_cast<uint64_t*>(smem_raw);  // [3] instance landed
  uint64_t* tb_bar = inst_bar + kInstStages;                   // template part TB landed
  uint64_t* tc_bar = tb_bar + 1;                               // template part TC landed
  uint64_t* full_bar = tc_bar + 1;                             // [2] tile staged (coords, base point, template issued)
  uint64_t* bdone_bar = full_bar + 2;                          // [2] consumers are through the integration phase
  uint64_t* done_bar = bdone_bar + 2;                          // [2] consumers finished the tile
  int32_t* s_rec = reinterpret_cast<int32_t*>(smem_raw + 128); // [2][4] tb generation, tc generation, rotation mode
  T* sbase = reinterpret_cast<T*>(smem_raw + 192);             // [2][8] w1*bx, w2*by, sin/cos(w1 bx), sin/cos(w2 by)
  int32_t* s_inst = reinterpret_cast<int32_t*>(smem_raw + kSmemHeader);
  int32_t* s_tb = s_inst + kInstStages * args.inst_words;
  int32_t* s_tc = s_tb + args.tb_words;
  V2* vxy = reinterpret_cast<V2*>(s_tc + args.tc_words);  // [2][max_vert]
  T* sloc = reinterpret_cast<T*>(vxy + 2 * args.max_vert);  // [1 + max_elem][9]

  const int tid = threadIdx.x;
  // this CTA's tiles: its share of the leading (progress-reporting) tiles, dealt round-robin, then one
  // contiguous block of the rest
  const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
  const int n_lead = cta < args.n_strided ? (args.n_strided - cta + grid - 1) / grid : 0;
  const int rest = args.n_tiles - args.n_strided;
  const int blk_lo = args.n_strided + (int)(((int64_t)rest * cta) / grid);
  const int blk_hi = args.n_strided + (int)(((int64_t)rest * (cta + 1)) / grid);
  const int n_local = n_lead + (blk_hi - blk_lo);
  auto tile_at = [&](int it) { return __ldg(args.tile_list + (it < n_lead ? cta + it * grid : blk_lo + (it - n_lead))); };

  if (tid == 0) {
    for (int i = 0; i < kInstStages + 2; ++i) mbar_init(inst_bar + i, 1);  // one expect_tx arrival each
    for (int i = 0; i < 2; ++i) mbar_init(full_bar + i, 1);
    for (int i = 0; i < 2; ++i) mbar_init(bdone_bar + i, kWarps);
    for (int i = 0; i < 2; ++i) mbar_init(done_bar + i, kWarps);
    fence_mbar_init();
  }
  if (tid < kSlots) sloc[tid] = T(0);  // row 0 of the table: the "no contribution" target of the packed codes
  __syncthreads();

  if (tid >= CONSUMERS) {
    // =================================== producer warp =======================================
    const int lane = tid - CONSUMERS;
    if (n_local == 0) return;
    auto issue_inst = [&](int it, const int4& d) {
      const int slot = it % kInstStages;
      mbar_expect_tx(inst_bar + slot, (uint32_t)d.y * 4u);
      bulk_g2s(s_inst + slot * args.inst_words, args.inst_blob + d.x, (uint32_t)d.y * 4u, inst_bar + slot);
    };
    int4 d_next = __ldg(args.tile_desc + tile_at(0));
    if (lane == 0) issue_inst(0, d_next);
    int cur_tpl = -1, tb_gen = 0, tc_gen = 0;
    TFEM_T_DECL;
    for (int it = 0; it < n_local; ++it) {
      const int stage = it & 1;
      const int4 d = d_next;
      // coordinates / instance slots of tile it-2 are free once its consumers are done
      if (it >= 2) mbar_wait(done_bar + stage, ((it - 2) >> 1) & 1);
      TFEM_T(0);
      if (it + 1 < n_local) {
        d_next = __ldg(args.tile_desc + tile_at(it + 1));
        if (lane == 0) issue_inst(it + 1, d_next);
      }
      mbar_wait(inst_bar + it % kInstStages, (it / kInstStages) & 1);
      TFEM_T(1);
      const int32_t* inst = s_inst + (it % kInstStages) * args.inst_words;
      const int n_vert = inst[0];
      V2* dst = vxy + stage * args.max_vert;
      for (int i = lane; i < n_vert; i += 32)
        cp_async<(int)sizeof(V2)>(dst + i, reinterpret_cast<const V2*>(args.coords) + inst[kInstHeader + i]);
      TFEM_T(2);
      if (d.z != cur_tpl) {
        // the single template buffer: TB is free after the integration phase of tile it-1, TC after
        // its reduction phase.  Rare on lattice meshes (a CTA's tiles are congruent); on an
        // unstructured mesh every tile brings its own template and this is the steady state.
        const int4 td = __ldg(args.tpl_desc + d.z);
        if (it >= 1) mbar_wait(bdone_bar + ((it - 1) & 1), ((it - 1) >> 1) & 1);
        if (lane == 0) {
          mbar_expect_tx(tb_bar, (uint32_t)td.y * 4u);
          bulk_g2s(s_tb, args.tpl_blob + td.x, (uint32_t)td.y * 4u, tb_bar);
        }
        if (it >= 1) mbar_wait(done_bar + ((it - 1) & 1), ((it - 1) >> 1) & 1);
        if (lane == 0) {
          mbar_expect_tx(tc_bar, (uint32_t)td.w * 4u);
          bulk_g2s(s_tc, args.tpl_blob + td.z, (uint32_t)td.w * 4u, tc_bar);
        }
        ++tb_gen;
        ++tc_gen;
        cur_tpl = d.z;
      }
      TFEM_T(3);
      cp_async_wait_all();
      __syncwarp();
      int rot_mode = 0;
      if constexpr (SINSIN) {
        // the tile's only full-range sin/cos, at its base vertex (lane 0: x phase, lane 1: y phase), and
        // the largest phase distance of a tile vertex from it (bounds base -> centroid for every element)
        const V2 b = dst[inst[3]];
        const T pbx = args.src.p1 * b.x, pby = args.src.p2 * b.y;
        int key = 0;
        for (int i = lane; i < n_vert; i += 32) {
          const V2 p = dst[i];
          key = max(key, max(mag_key(fma(args.src.p1, p.x, -pbx)), mag_key(fma(args.src.p2, p.y, -pby))));
        }
        key = __reduce_max_sync(0xffffffffu, key);
        rot_mode = key < args.key_rot_small ? 0 : (key < args.key_rot_medium ? 1 : 2);
        T s, c;
        sincos_full(lane == 0 ? pbx : pby, s, c);
        T* sb = sbase + 8 * stage;
        if (lane == 0) {
          sb[0] = pbx; sb[1] = pby; sb[2] = s; sb[3] = c;
        } else if (lane == 1) {
          sb[4] = s; sb[5] = c;
        }
      }
      if (lane == 0) {
        s_rec[4 * stage + 0] = tb_gen;
        s_rec[4 * stage + 1] = tc_gen;
        s_rec[4 * stage + 2] = rot_mode;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar + stage);
      TFEM_T(4);
    }
#ifdef TFEM_DEBUG_TIMING
    if (lane == 0)
      for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&tfem_timing_table[0][i]), (unsigned long long)t_acc[i]);
#endif
    return;
  }

  // ===================================== consumer warps ========================================
  const int lane = tid & 31, warp = tid >> 5;
  int seen_tb = 0, seen_tc = 0;
  TFEM_T_DECL;
  for (int it = 0; it < n_local; ++it) {
    const int stage = it & 1;
    mbar_wait(full_bar + stage, (it >> 1) & 1);
    TFEM_T(0);
    const int tb_gen = s_rec[4 * stage + 0], tc_gen = s_rec[4 * stage + 1], rot_mode = s_rec[4 * stage + 2];
    if (tb_gen != seen_tb) {  // a new template: TMA writes become visible to the threads that wait
      mbar_wait(tb_bar, (tb_gen - 1) & 1);
      seen_tb = tb_gen;
    }
    const int n_elem = s_tb[1], n_rows = s_tb[2], n_segs = s_tb[3], n_chunks = s_tb[4], n_heavy = s_tb[5], n_heavy_contrib = s_tb[6];
    const uint32_t* elem = reinterpret_cast<const uint32_t*>(s_tb + kTbHeader);
    const V2* xy = vxy + stage * args.max_vert;
    TFEM_T(1);

    // ---- B: local matrices and loads, each tile element once ----------------------------------
    const TriTable tt = tri_table(ORDER);  // folded at compile time
    const T* sb = sbase + 8 * stage;
    for (int base = warp * 32; base < ((TFEM_DEBUG_SKIP & 2) ? 0 : n_elem); base += CONSUMERS) {
      // lanes past the end recompute the last element (warp-wide reductions below need every lane)
      const int el = min(base + lane, n_elem - 1);
      const bool valid = base + lane < n_elem;
      const uint32_t packed = elem[el];
      const V2 p0 = xy[packed & 1023u], p1 = xy[(packed >> 10) & 1023u], p2 = xy[packed >> 20];
      const T ax = p1.x - p0.x, ay = p1.y - p0.y;  // J = [[ax, bx], [ay, by]] (basis.py:87-88)
      const T bx = p2.x - p0.x, by = p2.y - p0.y;
      const T det = ax * by - bx * ay;  // signed (element_tri.py:139)
      T* out = sloc + (el + 1) * kSlots;
      if constexpr (HAS_MAT) {
        // grad(phi_i) = e_i / det with e_1 = (by, -bx), e_2 = (-ay, ax), e_0 = -e_1 - e_2, so
        // sum_q dx grad(phi_i).grad(phi_j) = (wsum / det) e_i.e_j ; mass = det * reference mass
        const T kc = (args.alpha * quad.wsum) * fast_rcp(det);
        const T md = det * (args.beta * quad.mref[0]), mo = det * (args.beta * quad.mref[1]);
        const T s11 = fma(by, by, bx * bx), s22 = fma(ay, ay, ax * ax), s12 = -fma(by, ay, bx * ax);
        const T s01 = -s11 - s12, s02 = -s22 - s12, s00 = -s01 - s02;
        if (valid) {
          out[0] = fma(kc, s00, md);
          out[1] = fma(kc, s11, md);
          out[2] = fma(kc, s22, md);
          out[3] = fma(kc, s01, mo);
          out[4] = fma(kc, s12, mo);
          out[5] = fma(kc, s02, mo);
        }
      }
      if constexpr (HAS_LOAD) {
        T b0, b1, b2;
        if constexpr (SINSIN) {
          const T uax = args.src.p1 * ax, ubx = args.src.p1 * bx, uay = args.src.p2 * ay, uby = args.src.p2 * by;
          if (TFEM_DEBUG_SKIP & 16) {
            b0 = uax; b1 = ubx; b2 = uay + uby;
          } else {
          // sin/cos of the source phase at the centroid: rotate the base vertex's values
          const T dx = fma(args.src.p1 * T(1.0 / 3.0), (p0.x + p1.x) + p2.x, -sb[0]);
          const T dy = fma(args.src.p2 * T(1.0 / 3.0), (p0.y + p1.y) + p2.y, -sb[1]);
          T sx, cx, sy, cy;
          if (rot_mode == 2) {
            sincos_full(dx + sb[0], sx, cx);
            sincos_full(dy + sb[1], sy, cy);
          } else {
            T st, ct, su, cu;
            if (rot_mode == 0) {
              sincos_small(dx, st, ct);
              sincos_small(dy, su, cu);
            } else {
              sincos_medium(dx, st, ct);
              sincos_medium(dy, su, cu);
            }
            const T sbx = sb[2], cbx = sb[3], sby = sb[4], cby = sb[5];
            sx = fma(sbx, ct, cbx * st);
            cx = fma(cbx, ct, -(sbx * st));
            sy = fma(sby, cu, cby * su);
            cy = fma(cby, cu, -(sby * su));
          }
          // expansion about the centroid, degree by the largest phase any lane of the warp needs
          const int reach = __reduce_max_sync(0xffffffffu, max(max(mag_key(uax), mag_key(ubx)), max(mag_key(uay), mag_key(uby))));
          if (reach < args.key_deg4) {
            sinsin_moments<T, ORDER, 4>(sx, cx, sy, cy, uax, ubx, uay, uby, b0, b1, b2);
          } else if (reach < args.key_deg6) {
            sinsin_moments<T, ORDER, 6>(sx, cx, sy, cy, uax, ubx, uay, uby, b0, b1, b2);
          } else if (reach < args.key_deg8) {
            sinsin_moments<T, ORDER, 8>(sx, cx, sy, cy, uax, ubx, uay, uby, b0, b1, b2);
          } else {  // coarse element: evaluate the source directly
            b0 = b1 = b2 = T(0);
#pragma unroll 1
            for (int q = 0; q < NQV; ++q) {
              const T fx = sin(args.src.p1 * fma(quad.l1[q], ax, fma(quad.l2[q], bx, p0.x)));
              const T fy = sin(args.src.p2 * fma(quad.l1[q], ay, fma(quad.l2[q], by, p0.y)));
              const T wf = quad.w[q] * (fx * fy);
              b0 = fma(wf, quad.l0[q], b0);
              b1 = fma(wf, quad.l1[q], b1);
              b2 = fma(wf, quad.l2[q], b2);
            }
          }
          }
        } else {  // constant source: the moments of the basis functions are compile-time constants
          double c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
          for (int q = 0; q < NQV; ++q) {
            const double l1 = tt.xi[q], l2 = tt.eta[q], l0 = 1.0 - l1 - l2, w = 0.5 * tt.w[q];
            c0 += w * l0;
            c1 += w * l1;
            c2 += w * l2;
          }
          b0 = T(c0); b1 = T(c1); b2 = T(c2);
        }
        const T amp = args.src.p0 * det;
        if (valid) {
          out[6] = amp * b0;
          out[7] = amp * b1;
          out[8] = amp * b2;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bdone_bar + stage);  // the producer may replace the TB part
    consumer_sync<CONSUMERS>();
    TFEM_T(2);

    // ---- C: one lane per CSR entry / one thread per row; contributions in increasing element id --
    if (tc_gen != seen_tc) {
      mbar_wait(tc_bar, (tc_gen - 1) & 1);
      seen_tc = tc_gen;
    }
    mbar_wait(inst_bar + it % kInstStages, (it / kInstStages) & 1);  // the instance's TMA writes, for this thread
    TFEM_T(3);
    const int32_t* inst = s_inst + (it % kInstStages) * args.inst_words;
    const int32_t* seg_start = inst + kInstHeader + pad4(inst[0]);
    const int32_t* row_id = seg_start + pad4(n_segs);
    const uint32_t* pair = reinterpret_cast<const uint32_t*>(s_tc);
    const uint4* row_chunk = reinterpret_cast<const uint4*>(pair + 32 * n_segs);
    const uint16_t* row_diag = reinterpret_cast<const uint16_t*>(row_chunk + n_chunks);
    const uint16_t* heavy_seg = row_diag + 2 * pad4((n_rows + 1) >> 1);
    const uint16_t* heavy_contrib = heavy_seg + 2 * pad4((n_heavy + 2) >> 1);
    const uint16_t* heavy_pos = heavy_contrib + 2 * pad4((n_heavy_contrib + 1) >> 1);
    auto store = [&](uint32_t code, T value) {
      const uint32_t pos = (uint32_t)seg_start[code >> 5] + (code & 31u);
      if (!(TFEM_DEBUG_SKIP & 8) || value == T(1.2345e30)) args.csr_val[pos] = value;
    };
    if constexpr (HAS_MAT) {
      // Entries with <= 2 contributions (every off-diagonal of a manifold mesh): one lane per entry,
      // one warp per segment of <= 32 consecutive csr_val slots (coalesced stores).  One word per lane
      // names both contributions; a missing one is code 0 (the zero row), so there is no count and no
      // inner loop.
      constexpr int kSegUnroll = TFEM_SEG_UNROLL;
#pragma unroll kSegUnroll
      for (int sg = warp; sg < ((TFEM_DEBUG_SKIP & 4) ? 0 : n_segs); sg += kWarps) {
        const uint32_t word = pair[sg * 32 + lane];
        if (word != 0xffffffffu) {
          const T value = sloc[word & 0xffffu] + sloc[word >> 16];
          if (!(TFEM_DEBUG_SKIP & 8) || value == T(1.2345e30)) args.csr_val[(uint32_t)seg_start[sg] + lane] = value;
        }
      }
      TFEM_T(4);
      // the few other entries with > 2 contributions (non-manifold edges, degenerate elements)
      for (int h = tid; h < ((TFEM_DEBUG_SKIP & 4) ? 0 : n_heavy); h += CONSUMERS) {
        T acc = T(0);
        for (int s = heavy_seg[h]; s < heavy_seg[h + 1]; ++s) acc += sloc[heavy_contrib[s]];
        store(heavy_pos[h], acc);
      }
    }
    // one thread per owned row: its load entry and its diagonal share one element list, read as
    // chunks of 7 codes + link (one 16 B shared-memory load per chunk)
    for (int j = tid; j < ((TFEM_DEBUG_SKIP & 4) ? 0 : n_rows); j += CONSUMERS) {
      T rhs = T(0), diag = T(0);
      uint32_t chunk = (uint32_t)j;
      do {
        const uint4 w = row_chunk[chunk];
        const uint32_t code[7] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16, w.z & 0xffffu, w.z >> 16, w.w & 0xffffu};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          if constexpr (HAS_LOAD) rhs += sloc[code[k] + 6];
          if constexpr (HAS_MAT) diag += sloc[code[k]];
        }
        chunk = w.w >> 16;
      } while (chunk != 0);
      if constexpr (HAS_LOAD) {
        if (!(TFEM_DEBUG_SKIP & 8) || rhs == T(1.2345e30)) args.load[row_id[j]] = rhs;
      }
      if constexpr (HAS_MAT) {
        const uint32_t code = row_diag[j];
        if (code != 0xffffu) store(code, diag);
      }
    }
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

template <typename T, int CONSUMERS, int ORDER, int SRC, bool HAS_MAT>
int launch_tiled(const tfem_tile_plan* hp, TiledArgs<T>& args, const QuadT<T>& quad, cudaStream_t s) {
  args.inst_words = pad4(hp->max_inst_words);
  args.tb_words = pad4(hp->max_tb_words);
  args.tc_words = pad4(hp->max_tc_words);
  const size_t smem = kSmemHeader + 4 * ((size_t)kInstStages * args.inst_words + (size_t)args.tb_words + (size_t)args.tc_words) +
                      sizeof(T) * ((size_t)4 * hp->max_vert + (size_t)kSlots * (hp->max_elem + 1));
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
  kern<<<grid, CONSUMERS + 32, smem, s>>>(args, quad);
  return check_launch();
}

template <typename T, int CONSUMERS, int ORDER>
int dispatch_tiled(const tfem_tile_plan* hp, TiledArgs<T>& args, const QuadT<T>& quad, cudaStream_t s) {
  const int kind = args.load ? args.src.kind : TFEM_SRC_NONE;
#ifdef TFEM_FAST_BUILD  // experiment builds: only the headline instantiation (fp64, 4-point rule, K+M and sin-sin load)
  if constexpr (sizeof(T) == 8 && ORDER == 3) {
    if (args.csr_val && kind == TFEM_SRC_SINSIN) return launch_tiled<T, CONSUMERS, ORDER, TFEM_SRC_SINSIN, true>(hp, args, quad, s);
  }
  return TFEM_ERR_UNSUPPORTED;
#else
  if (args.csr_val) {
    if (kind == TFEM_SRC_SINSIN) return launch_tiled<T, CONSUMERS, ORDER, TFEM_SRC_SINSIN, true>(hp, args, quad, s);
    if (kind == TFEM_SRC_CONST) return launch_tiled<T, CONSUMERS, ORDER, TFEM_SRC_CONST, true>(hp, args, quad, s);
    return launch_tiled<T, CONSUMERS, ORDER, TFEM_SRC_NONE, true>(hp, args, quad, s);
  }
  if (kind == TFEM_SRC_SINSIN) return launch_tiled<T, CONSUMERS, ORDER, TFEM_SRC_SINSIN, false>(hp, args, quad, s);
  if (kind == TFEM_SRC_CONST) return launch_tiled<T, CONSUMERS, ORDER, TFEM_SRC_CONST, false>(hp, args, quad, s);
  return TFEM_ERR_BAD_ARG;
#endif
}

template <typename T>
int assemble_tiled(const tfem_tile_plan* hp, const T* coords, int quad_order, const tfem_bilinear* form,
                   const tfem_source* source, T* csr_val, T* load, void* stream) {
  if (!hp || hp->n_tiles < 0) return TFEM_ERR_BAD_ARG;
  if (hp->n_tiles == 0) return TFEM_OK;
  if (!coords || (!csr_val && !load)) return TFEM_ERR_BAD_ARG;
  if (csr_val && !form) return TFEM_ERR_BAD_ARG;
  if (!hp->tile_list || !hp->tile_desc || !hp->inst_blob || !hp->tpl_desc || !hp->tpl_blob) return TFEM_ERR_BAD_ARG;
  if (hp->max_vert > 1024 || hp->max_elem > 7000 || hp->max_vert < 0 || hp->max_elem < 0) return TFEM_ERR_TOO_LARGE;
  if (hp->n_tiles > kMaxIndex) return TFEM_ERR_TOO_LARGE;
  if (tri_n_q(quad_order) == 0) return TFEM_ERR_UNSUPPORTED;
  TiledArgs<T> args{};
  args.src = make_source<T>(source);
  if (load && (args.src.kind == TFEM_SRC_SAMPLED || args.src.kind < TFEM_SRC_NONE || args.src.kind > TFEM_SRC_SINSIN))
    return TFEM_ERR_BAD_ARG;  // sampled sources go through tfem_tri_p1_local_forms
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
  args.coords = coords;
  args.alpha = form ? T(form->alpha) : T(0);
  args.beta = form ? T(form->beta) : T(0);
  args.csr_val = csr_val;
  args.load = load;
  // accuracy thresholds of the source evaluation (see sincos_small / sincos_medium / sinsin_moments)
  const double spread = centroid_spread(quad_order);
  args.key_rot_small = host_mag_key(0.04, T(0));
  args.key_rot_medium = host_mag_key(0.2, T(0));
  args.key_deg4 = host_mag_key(4.0e-3 / spread, T(0));  // t^5/120 < 1e-14
  args.key_deg6 = host_mag_key(3.0e-2 / spread, T(0));  // t^7/5040 < 5e-15
  args.key_deg8 = host_mag_key(1.0e-1 / spread, T(0));  // t^9/362880 < 3e-15
  auto s = static_cast<cudaStream_t>(stream);
  int consumers = hp->consumer_threads;
  if (consumers == 0) consumers = 384;
#define TFEM_DISPATCH_ORDER(C)                                         \
  switch (quad_order) {                                                \
    case 1: return dispatch_tiled<T, C, 1>(hp, args, quad, s);         \
    case 2: return dispatch_tiled<T, C, 2>(hp, args, quad, s);         \
    case 3: return dispatch_tiled<T, C, 3>(hp, args, quad, s);         \
    default: return dispatch_tiled<T, C, 4>(hp, args, quad, s);        \
  }
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
