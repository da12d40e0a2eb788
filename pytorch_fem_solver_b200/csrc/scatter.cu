// Local-to-global scatter as a deterministic segmented reduction (no float atomics), the
// multi-GPU interface pack/unpack, and the integer COO key builder of the symbolic phase.
// Reference: basis/abstract_basis.py:87-91,106-110 (index_put_ accumulate) and
// basis/basis.py:72-77 (bilinear_form_idx / linear_form_idx).
#include "common.cuh"

namespace tfem {

// out[p] = sum_{s in [seg[p], seg[p+1])} values[perm[s]], left to right.
// One thread per output entry: consecutive threads own consecutive CSR entries, so `seg`, `perm`
// and `out` stream coalesced; `values` is a gather whose locality follows the mesh numbering.
template <typename T>
__global__ void __launch_bounds__(256) segment_reduce_kernel(int n_out, const int32_t* __restrict__ seg,
                                                             const int32_t* __restrict__ perm,
                                                             const T* __restrict__ values,
                                                             T* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_out) return;
  const int begin = __ldg(seg + p);
  const int end = __ldg(seg + p + 1);
  T acc = T(0);
  for (int s = begin; s < end; ++s) acc += __ldg(values + __ldg(perm + s));
  out[p] = acc;
}

template <typename T>
__global__ void __launch_bounds__(256) pack_kernel(int n, const int32_t* __restrict__ idx,
                                                   const T* __restrict__ src, T* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = __ldg(src + __ldg(idx + i));
}

template <typename T>
__global__ void __launch_bounds__(256) unpack_add_kernel(int n, const int32_t* __restrict__ idx,
                                                         const T* __restrict__ buf, T* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int j = __ldg(idx + i);
    dst[j] = dst[j] + __ldg(buf + i);
  }
}

__global__ void __launch_bounds__(256) coo_keys_kernel(int64_t n_el, const int32_t* __restrict__ dof_conn,
                                                       int64_t n_dof, int64_t* __restrict__ keys) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
  int64_t v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) v[k] = __ldg(dof_conn + 3 * e + k);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) keys[9 * e + 3 * i + j] = v[j] * n_dof + v[i];  // row = conn[e,j], col = conn[e,i]
}

template <typename T>
int segment_reduce(int64_t n_out, const int32_t* seg, const int32_t* perm, const T* values, T* out,
                   void* stream) {
  if (n_out < 0) return TFEM_ERR_BAD_ARG;
  if (n_out == 0) return TFEM_OK;
  if (!seg || !perm || !values || !out) return TFEM_ERR_BAD_ARG;
  if (n_out > kMaxIndex) return TFEM_ERR_TOO_LARGE;
  const int threads = 256;
  segment_reduce_kernel<T><<<blocks_for(n_out, threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      (int)n_out, seg, perm, values, out);
  return check_launch();
}

// pack_kernel behind a device-side dependency: wait until `progress` (bumped with release semantics by
// the assembly kernel's warps) has reached `target`; the comparison is wrap-safe.
template <typename T>
__global__ void __launch_bounds__(128) pack_after_kernel(int n, const int32_t* __restrict__ idx, const T* src,
                                                         T* __restrict__ buf, const uint32_t* progress, uint32_t target) {
  if (threadIdx.x == 0) {
    uint32_t seen;
    for (uint32_t spin = 0;; ++spin) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(progress) : "memory");
      if ((int32_t)(seen - target) >= 0) break;
      if (spin > (1u << 26)) __trap();  // a few seconds: the producer never ran
      __nanosleep(200);
    }
  }
  __syncthreads();
  // src is being written by another kernel during this launch: plain loads, no __ldg
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = src[__ldg(idx + i)];
}

template <typename T>
int iface_pack_after(int64_t n, const int32_t* idx, const T* src, T* buf, const uint32_t* progress, uint32_t target,
                     void* stream) {
  if (n < 0) return TFEM_ERR_BAD_ARG;
  if (n == 0) return TFEM_OK;
  if (!idx || !src || !buf || !progress) return TFEM_ERR_BAD_ARG;
  if (n > kMaxIndex) return TFEM_ERR_TOO_LARGE;
  // at most 32 small blocks of 128 threads: they spin beside a persistent grid and must fit in what it leaves free --
  // next to a CTA of the role-specialised kernel (832 threads x 72 registers) an SM still has 5 632 registers, i.e. one
  // such block (128 x 24) fits on EVERY SM, reserved or not
  const unsigned blocks = blocks_for(n, 128) < 32u ? blocks_for(n, 128) : 32u;
  pack_after_kernel<T><<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>((int)n, idx, src, buf, progress, target);
  return check_launch();
}

template <typename T>
int iface_pack(int64_t n, const int32_t* idx, const T* src, T* buf, void* stream) {
  if (n < 0) return TFEM_ERR_BAD_ARG;
  if (n == 0) return TFEM_OK;
  if (!idx || !src || !buf) return TFEM_ERR_BAD_ARG;
  if (n > kMaxIndex) return TFEM_ERR_TOO_LARGE;
  pack_kernel<T><<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>((int)n, idx, src, buf);
  return check_launch();
}

template <typename T>
int iface_unpack_add(int64_t n, const int32_t* idx, const T* buf, T* dst, void* stream) {
  if (n < 0) return TFEM_ERR_BAD_ARG;
  if (n == 0) return TFEM_OK;
  if (!idx || !buf || !dst) return TFEM_ERR_BAD_ARG;
  if (n > kMaxIndex) return TFEM_ERR_TOO_LARGE;
  unpack_add_kernel<T><<<blocks_for(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>((int)n, idx, buf, dst);
  return check_launch();
}

}  // namespace tfem

#define TFEM_SCATTER_API(T, SUF)                                                                    \
  extern "C" int tfem_scatter_bilinear_##SUF(int64_t nnz, const int32_t* seg, const int32_t* perm,  \
                                             const T* local, T* csr_val, void* stream) {            \
    return tfem::segment_reduce<T>(nnz, seg, perm, local, csr_val, stream);                         \
  }                                                                                                 \
  extern "C" int tfem_scatter_linear_##SUF(int64_t n_dof, const int32_t* seg, const int32_t* perm,  \
                                           const T* local, T* vec, void* stream) {                  \
    return tfem::segment_reduce<T>(n_dof, seg, perm, local, vec, stream);                           \
  }                                                                                                 \
  extern "C" int tfem_iface_pack_##SUF(int64_t n, const int32_t* idx, const T* src, T* buf,         \
                                       void* stream) {                                              \
    return tfem::iface_pack<T>(n, idx, src, buf, stream);                                           \
  }                                                                                                 \
  extern "C" int tfem_iface_pack_after_##SUF(int64_t n, const int32_t* idx, const T* src, T* buf,   \
                                             const uint32_t* progress, uint32_t target, void* stream) { \
    return tfem::iface_pack_after<T>(n, idx, src, buf, progress, target, stream);                   \
  }                                                                                                 \
  extern "C" int tfem_iface_unpack_add_##SUF(int64_t n, const int32_t* idx, const T* buf, T* dst,   \
                                             void* stream) {                                        \
    return tfem::iface_unpack_add<T>(n, idx, buf, dst, stream);                                     \
  }

TFEM_SCATTER_API(double, f64)
TFEM_SCATTER_API(float, f32)

extern "C" int tfem_coo_keys(int64_t n_el, const int32_t* dof_conn, int64_t n_dof, int64_t* keys, void* stream) {
  if (n_el < 0 || n_dof <= 0) return TFEM_ERR_BAD_ARG;
  if (n_el == 0) return TFEM_OK;
  if (!dof_conn || !keys) return TFEM_ERR_BAD_ARG;
  tfem::coo_keys_kernel<<<tfem::blocks_for(n_el, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n_el, dof_conn, n_dof, keys);
  return tfem::check_launch();
}

extern "C" int tfem_abi_version(void) { return 1; }

extern "C" const char* tfem_status_string(int status) {
  switch (status) {
    case TFEM_OK: return "ok";
    case TFEM_ERR_BAD_ARG: return "bad argument";
    case TFEM_ERR_UNSUPPORTED: return "unsupported quadrature order";
    case TFEM_ERR_LAUNCH: return "kernel launch failed";
    case TFEM_ERR_TOO_LARGE: return "count exceeds 32-bit index range";
    default: return "unknown status";
  }
}

extern "C" int tfem_set_device(int device) {
  return cudaSetDevice(device) == cudaSuccess ? TFEM_OK : TFEM_ERR_BAD_ARG;
}

extern "C" int tfem_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}
