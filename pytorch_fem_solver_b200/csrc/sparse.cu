// CSR sparse matrix-vector product for the iterative solve on the assembled system
// (replaces the dense torch.linalg.solve of basis/abstract_basis.py:177-195 when
// the system is too large to densify).  HBM bound: 12 B per stored entry (value + column) plus the
// vectors; P1 rows hold ~7 entries, so 8 lanes share a row and a warp streams 4 consecutive rows,
// i.e. contiguous csr_val / col segments.
#include "common.cuh"

namespace tfem {

// y[i] = keep[i] ? sum_k val[k] x[col[k]] : 0   (keep == nullptr: every row)
template <typename T, int LANES>
__global__ void __launch_bounds__(256) csr_spmv_kernel(int n_rows, const int32_t* __restrict__ crow,
                                                       const int32_t* __restrict__ col, const T* __restrict__ val,
                                                       const T* __restrict__ x, const uint8_t* __restrict__ keep,
                                                       T* __restrict__ y) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = (int)(t / LANES);
  const int lane = (int)(t % LANES);
  T acc = T(0);
  const bool live = row < n_rows && (keep == nullptr || keep[row] != 0);
  if (live) {
    const int begin = __ldg(crow + row), end = __ldg(crow + row + 1);
    for (int k = begin + lane; k < end; k += LANES) acc = fma(__ldg(val + k), __ldg(x + __ldg(col + k)), acc);
  }
#pragma unroll
  for (int offset = LANES / 2; offset > 0; offset >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, offset, LANES);
  if (lane == 0 && row < n_rows) y[row] = live ? acc : T(0);
}

template <typename T>
int csr_spmv(int64_t n_rows, const int32_t* crow, const int32_t* col, const T* val, const T* x, const uint8_t* keep,
             T* y, void* stream) {
  if (n_rows < 0) return TFEM_ERR_BAD_ARG;
  if (n_rows == 0) return TFEM_OK;
  if (!crow || !col || !val || !x || !y) return TFEM_ERR_BAD_ARG;
  if (n_rows > kMaxIndex / 8) return TFEM_ERR_TOO_LARGE;
  constexpr int kLanes = 8;
  csr_spmv_kernel<T, kLanes><<<blocks_for(n_rows * kLanes, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (int)n_rows, crow, col, val, x, keep, y);
  return check_launch();
}

}  // namespace tfem

extern "C" int tfem_csr_spmv_f64(int64_t n_rows, const int32_t* crow, const int32_t* col, const double* val,
                                 const double* x, const uint8_t* keep, double* y, void* stream) {
  return tfem::csr_spmv<double>(n_rows, crow, col, val, x, keep, y, stream);
}

extern "C" int tfem_csr_spmv_f32(int64_t n_rows, const int32_t* crow, const int32_t* col, const float* val,
                                 const float* x, const uint8_t* keep, float* y, void* stream) {
  return tfem::csr_spmv<float>(n_rows, crow, col, val, x, keep, y, stream);
}
