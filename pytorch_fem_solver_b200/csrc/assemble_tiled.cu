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
//   TMA bulk copy of the blob of tile t+2        wait full[t]
//   (cp.async.bulk + mbarrier complete_tx)       A  rotate the tile's base sin/cos to every vertex
//   wait blob t+1                                B  integrate every tile element ONCE: 6 matrix
//   cp.async gather of the vertex coordinates       entries (symmetric form) + 3 load entries
//   of tile t+1 (16 B per vertex) -> shared      C  one thread per owned row sums its incident
//   full-range sin/cos at ONE base vertex           elements in increasing element order into a
//   arrive full[t+1]                                 shared image of the row (no atomics)
//                                                D  stream the image to csr_val in runs of
//                                                   consecutive rows (coalesced 8 B stores)
//                                                arrive done[t]
//
// so global-memory latency (blob, coordinate gather) and the only library sin/cos of a tile are
// off the consumers' critical path.  f at the quadrature points is obtained by rotating the
// vertex-0 sin/cos by the in-element phase with a short Taylor series: no fp64 sin() per point.
// Elements on a tile border are recomputed by the neighbouring tile (halo ~15%).
// HBM traffic is coords + index blob + outputs, each touched once.
#include "common.cuh"

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
constexpr int kBlobHeader = 8;
constexpr int kBlobStages = 3;
constexpr int kSmemHeader = 256;  // mbarriers (7 x 8 B) + two base-point records (2 x 6 values)

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

struct BlobView {
  int n_vert, n_elem, n_rows, n_runs, n_corner, n_out, base_vertex;
  const int32_t* vert;
  const uint32_t* elem;
  const int32_t* row_id;
  const int32_t* row_meta;
  const int32_t* row_cptr;
  const uint32_t* corner;
  const int32_t* run_start;
  const int32_t* run_meta;
};

__device__ __forceinline__ BlobView view_blob(const int32_t* b) {
  BlobView v;
  v.n_vert = b[0]; v.n_elem = b[1]; v.n_rows = b[2]; v.n_runs = b[3]; v.n_corner = b[4]; v.n_out = b[5];
  v.base_vertex = b[6];
  v.vert = b + kBlobHeader;
  v.elem = reinterpret_cast<const uint32_t*>(v.vert + pad4(v.n_vert));
  v.row_id = reinterpret_cast<const int32_t*>(v.elem + pad4(v.n_elem));
  v.row_meta = v.row_id + pad4(v.n_rows);
  v.row_cptr = v.row_meta + pad4(v.n_rows);
  v.corner = reinterpret_cast<const uint32_t*>(v.row_cptr + pad4(v.n_rows + 1));
  v.run_start = reinterpret_cast<const int32_t*>(v.corner + pad4(v.n_corner));
  v.run_meta = v.run_start + pad4(v.n_runs);
  return v;
}

template <typename T, int CONSUMERS, int ORDER, int SRC, bool HAS_MAT>
__global__ void __launch_bounds__(CONSUMERS + 32, 2) assemble_tiled_kernel(
    const int n_tiles, const int32_t* __restrict__ tile_off, const int32_t* __restrict__ blob,
    const int max_vert, const int elem_stride, const int max_out, const int max_blob_words,
    const T* __restrict__ coords, const QuadT<T> quad, const T alpha, const T beta, const SourceT<T> src,
    T* __restrict__ csr_val, T* __restrict__ load) {
  constexpr int NQV = NQ<ORDER>::value;
  constexpr bool HAS_LOAD = SRC != TFEM_SRC_NONE;
  constexpr bool SINSIN = SRC == TFEM_SRC_SINSIN;
  using V2 = typename Vec2<T>::type;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [ header | blob x3 | vertex coordinates x2 | sin/cos fields | sloc[9][elem_stride] | sout ]
  uint64_t* blob_bar = reinterpret_cast<uint64_t*>(smem_raw);  // [3] TMA landed
  uint64_t* full_bar = blob_bar + kBlobStages;                 // [2] tile staged (coords + base point)
  uint64_t* done_bar = full_bar + 2;                           // [2] consumers finished the tile
  T* sbase = reinterpret_cast<T*>(smem_raw + 64);              // [2][6] bx, by, sin/cos(w bx), sin/cos(w by)
  int32_t* sblob = reinterpret_cast<int32_t*>(smem_raw + kSmemHeader);
  V2* vxy = reinterpret_cast<V2*>(sblob + kBlobStages * max_blob_words);  // [2][max_vert]
  T* trig = reinterpret_cast<T*>(vxy + 2 * max_vert);                     // [4][max_vert]
  T* sloc = trig + (SINSIN ? 4 : 0) * max_vert;
  T* sout = sloc + 9 * elem_stride;
  (void)max_out;

  const int tid = threadIdx.x;
  const int n_local = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0) {
    for (int i = 0; i < kBlobStages; ++i) mbar_init(blob_bar + i, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(done_bar + i, 1);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= CONSUMERS) {
    // =================================== producer warp =======================================
    const int lane = tid - CONSUMERS;
    auto issue_blob = [&](int it) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int off0 = __ldg(tile_off + tile);
      const uint32_t bytes = (uint32_t)(__ldg(tile_off + tile + 1) - off0) * 4u;
      const int slot = it % kBlobStages;
      mbar_expect_tx(blob_bar + slot, bytes);
      bulk_g2s(sblob + slot * max_blob_words, blob + off0, bytes, blob_bar + slot);
    };
    if (lane == 0) {
      issue_blob(0);
      if (n_local > 1) issue_blob(1);
    }
    for (int it = 0; it < n_local; ++it) {
      const int slot = it % kBlobStages, buf = it & 1;
      // stage tile `it` while the consumers work on tile it-1: vxy[buf] / sbase[buf] were last
      // read by tile it-2
      if (it >= 2) mbar_wait(done_bar + (it & 1), ((it - 2) >> 1) & 1);
      mbar_wait(blob_bar + slot, (it / kBlobStages) & 1);
      const BlobView bv = view_blob(sblob + slot * max_blob_words);
      V2* dst = vxy + buf * max_vert;
      for (int i = lane; i < bv.n_vert; i += 32)
        cp_async<(int)sizeof(V2)>(dst + i, reinterpret_cast<const V2*>(coords) + bv.vert[i]);
      if constexpr (SINSIN) {
        T bx, by, sx, cx, sy, cy;
        load_xy(coords, bv.base_vertex, bx, by);
        sincos_full(src.p1 * bx, sx, cx);
        sincos_full(src.p2 * by, sy, cy);
        if (lane == 0) {
          T* sb = sbase + 6 * buf;
          sb[0] = bx; sb[1] = by; sb[2] = sx; sb[3] = cx; sb[4] = sy; sb[5] = cy;
        }
      }
      cp_async_wait_all();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar + buf);
      // the blob slot of tile it+2 is the one tile it-1 used: refill it once that tile is done
      if (it + 2 < n_local) {
        if (it >= 1) mbar_wait(done_bar + ((it - 1) & 1), ((it - 1) >> 1) & 1);
        if (lane == 0) issue_blob(it + 2);
      }
    }
    return;
  }

  // ===================================== consumer warps ========================================
  for (int it = 0; it < n_local; ++it) {
    const int slot = it % kBlobStages, buf = it & 1;
    mbar_wait(full_bar + buf, (it >> 1) & 1);
    mbar_wait(blob_bar + slot, (it / kBlobStages) & 1);  // TMA writes visible to this thread too
    const BlobView bv = view_blob(sblob + slot * max_blob_words);
    const V2* xy = vxy + buf * max_vert;

    // ---- A: sin/cos of the source phase at every tile vertex, rotated from the base vertex ----
    if constexpr (SINSIN) {
      const T* sb = sbase + 6 * buf;
      const T bx = sb[0], by = sb[1], sbx = sb[2], cbx = sb[3], sby = sb[4], cby = sb[5];
      for (int i = tid; i < bv.n_vert; i += CONSUMERS) {
        const V2 p = xy[i];
        T s, c;
        sincos_about(src.p1, p.x, bx, sbx, cbx, s, c);
        trig[0 * max_vert + i] = s;
        trig[1 * max_vert + i] = c;
        sincos_about(src.p2, p.y, by, sby, cby, s, c);
        trig[2 * max_vert + i] = s;
        trig[3 * max_vert + i] = c;
      }
    }
    if constexpr (HAS_MAT) {
      for (int i = tid; i < bv.n_out; i += CONSUMERS) sout[i] = T(0);
    }
    if constexpr (SINSIN || HAS_MAT) consumer_sync<CONSUMERS>();

    // ---- B: local matrices and loads, each tile element once ----------------------------------
    for (int el = tid; el < bv.n_elem; el += CONSUMERS) {
      const uint32_t packed = bv.elem[el];
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

    // ---- C: one thread per owned row gathers its corners in increasing element order --------
    for (int j = tid; j < bv.n_rows; j += CONSUMERS) {
      const int meta = bv.row_meta[j];
      const int base = meta & 0xffff, pos_diag = (meta >> 16) & 0xff;
      const int c0 = bv.row_cptr[j], c1 = bv.row_cptr[j + 1];
      T diag = T(0), rhs = T(0);
      for (int cidx = c0; cidx < c1; ++cidx) {
        const uint32_t cw = bv.corner[cidx];
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
      if constexpr (HAS_LOAD) load[bv.row_id[j]] = rhs;
    }

    // ---- D: stream the row images out, one warp per run of consecutive rows ------------------
    if constexpr (HAS_MAT) {
      consumer_sync<CONSUMERS>();
      const int lane = tid & 31, warp = tid >> 5;
      for (int r = warp; r < bv.n_runs; r += CONSUMERS / 32) {
        const int gstart = bv.run_start[r];
        const int meta = bv.run_meta[r];
        const int base = meta & 0xffff, len = (meta >> 16) & 0xffff;
        for (int i = lane; i < len; i += 32) csr_val[(int64_t)gstart + i] = sout[base + i];
      }
    }
    consumer_sync<CONSUMERS>();  // sloc / sout / trig / blob slot free for the next tile
    if (tid == 0) mbar_arrive(done_bar + buf);
  }
}

template <typename T>
size_t tiled_smem_bytes(const tfem_tile_plan* hp, int src_kind, int* elem_stride, int* blob_words) {
  const int trig_fields = src_kind == TFEM_SRC_SINSIN ? 4 : 0;
  *elem_stride = (hp->max_elem + 31) & ~31;
  *blob_words = (hp->max_blob_words + 3) & ~3;
  return kSmemHeader + 4 * (size_t)kBlobStages * *blob_words +
         sizeof(T) * ((size_t)(4 + trig_fields) * hp->max_vert + (size_t)9 * *elem_stride + (size_t)hp->max_out);
}

template <typename T, int CONSUMERS, int ORDER, int SRC, bool HAS_MAT>
int launch_tiled(const tfem_tile_plan* hp, const T* coords, const QuadT<T>& quad, T alpha, T beta,
                 const SourceT<T>& src, T* csr_val, T* load, cudaStream_t s) {
  int elem_stride = 0, blob_words = 0;
  const size_t smem = tiled_smem_bytes<T>(hp, SRC, &elem_stride, &blob_words);
  if (smem > 227 * 1024) return TFEM_ERR_TOO_LARGE;
  auto kern = assemble_tiled_kernel<T, CONSUMERS, ORDER, SRC, HAS_MAT>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return TFEM_ERR_LAUNCH;
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CONSUMERS + 32, smem) != cudaSuccess || per_sm < 1)
    return TFEM_ERR_LAUNCH;
  const int64_t resident = (int64_t)sms * per_sm;  // persistent grid: every CTA is co-resident
  const unsigned grid = (unsigned)(hp->n_tiles < resident ? hp->n_tiles : resident);
  kern<<<grid, CONSUMERS + 32, smem, s>>>((int)hp->n_tiles, hp->tile_off, hp->blob, hp->max_vert, elem_stride,
                                         hp->max_out, blob_words, coords, quad, alpha, beta, src, csr_val, load);
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
  if (!hp->tile_off || !hp->blob) return TFEM_ERR_BAD_ARG;
  if (hp->max_vert > 1024 || hp->max_elem > 4096 || hp->max_out > 65535) return TFEM_ERR_TOO_LARGE;
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
  constexpr int CONSUMERS = 256;
  switch (quad_order) {
    case 1: return dispatch_tiled<T, CONSUMERS, 1>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
    case 2: return dispatch_tiled<T, CONSUMERS, 2>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
    case 3: return dispatch_tiled<T, CONSUMERS, 3>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
    default: return dispatch_tiled<T, CONSUMERS, 4>(hp, coords, quad, alpha, beta, src, csr_val, load, s);
  }
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
