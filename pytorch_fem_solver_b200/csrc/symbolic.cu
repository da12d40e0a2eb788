// One-time integer set-up of the sparse system (SURVEY 8(a) row a10, 8(b) tfem_csr_symbolic): the sorted-unique CSR
// pattern of the COO index maps of basis/basis.py:64-85 and the stable COO -> CSR permutations the deterministic
// scatter kernels walk, entirely on the device: CUB radix sort of the 64-bit (row, col) keys (only the bits a key can
// have are sorted), run-length encoding, prefix sums and binary searches.  Replaces the torch program
// sort / unique_consecutive / cumsum / bincount of csr.build_pattern (kept there for CPU tensors and as the oracle of
// the bit-exactness test).
#include <cub/cub.cuh>

#include "common.cuh"

namespace tfem {

__global__ void iota_kernel(int32_t* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)i;
}

// col[i] = key[i] % n_dof for the nnz unique keys; seg[nnz] = total (closes the exclusive scan of the run lengths)
__global__ void split_keys_kernel(const int64_t* __restrict__ keys, const int64_t* __restrict__ nnz_ptr, int64_t n_dof, int64_t total,
                                  int32_t* __restrict__ col, int32_t* __restrict__ seg) {
  const int64_t nnz = *nnz_ptr;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) col[i] = (int32_t)(keys[i] % n_dof);
  if (i == nnz) seg[nnz] = (int32_t)total;
}

// out[r] = number of sorted values < r * scale   (r = 0 .. n_dof): row pointers by binary search
template <typename K>
__global__ void lower_bound_kernel(const K* __restrict__ sorted, const int64_t* __restrict__ n_ptr, int64_t n_fixed, int64_t n_dof, int64_t scale,
                                   int32_t* __restrict__ out) {
  const int64_t n = n_ptr ? *n_ptr : n_fixed;
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_dof) return;
  const int64_t target = r * scale;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)sorted[mid] < target) lo = mid + 1;
    else hi = mid;
  }
  out[r] = (int32_t)lo;
}

inline int bits_for(int64_t v) {  // bits needed to represent values < v
  int b = 1;
  while (b < 63 && (int64_t(1) << b) < v) ++b;
  return b;
}

struct SymbolicLayout {
  size_t keys_in, keys_sorted, vals_in, counts, form_in, form_sorted, lvals_in, cub_temp, cub_bytes, total;
};

inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

inline int symbolic_layout(int64_t n_el, int64_t n_dof, SymbolicLayout* out) {
  const int64_t n9 = 9 * n_el, n3 = 3 * n_el;
  size_t sort_pairs = 0, sort_lin = 0, rle = 0, scan = 0;
  if (cub::DeviceRadixSort::SortPairs(nullptr, sort_pairs, (const int64_t*)nullptr, (int64_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr,
                                      (int)n9) != cudaSuccess)
    return TFEM_ERR_LAUNCH;
  if (cub::DeviceRadixSort::SortPairs(nullptr, sort_lin, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr,
                                      (int)n3) != cudaSuccess)
    return TFEM_ERR_LAUNCH;
  if (cub::DeviceRunLengthEncode::Encode(nullptr, rle, (const int64_t*)nullptr, (int64_t*)nullptr, (int32_t*)nullptr, (int64_t*)nullptr, (int)n9) !=
      cudaSuccess)
    return TFEM_ERR_LAUNCH;
  if (cub::DeviceScan::ExclusiveSum(nullptr, scan, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(n9 + 1)) != cudaSuccess) return TFEM_ERR_LAUNCH;
  size_t cub_bytes = sort_pairs;
  cub_bytes = sort_lin > cub_bytes ? sort_lin : cub_bytes;
  cub_bytes = rle > cub_bytes ? rle : cub_bytes;
  cub_bytes = scan > cub_bytes ? scan : cub_bytes;
  size_t at = 0;
  out->keys_in = at;     at += align256(8 * (size_t)n9);
  out->keys_sorted = at; at += align256(8 * (size_t)n9);
  out->vals_in = at;     at += align256(4 * (size_t)n9);
  out->counts = at;      at += align256(4 * (size_t)(n9 + 1));
  out->form_in = at;     at += align256(4 * (size_t)n3);
  out->form_sorted = at; at += align256(4 * (size_t)n3);
  out->lvals_in = at;    at += align256(4 * (size_t)n3);
  out->cub_temp = at;    at += align256(cub_bytes);
  out->cub_bytes = cub_bytes;
  out->total = at;
  (void)n_dof;
  return TFEM_OK;
}

}  // namespace tfem

extern "C" int tfem_coo_keys(int64_t n_el, const int32_t* dof_conn, int64_t n_dof, int64_t* keys, void* stream);

extern "C" int tfem_csr_symbolic_workspace(int64_t n_el, int64_t n_dof, int64_t* bytes) {
  if (n_el < 0 || n_dof < 0 || !bytes) return TFEM_ERR_BAD_ARG;
  if (9 * n_el > tfem::kMaxIndex) return TFEM_ERR_TOO_LARGE;
  tfem::SymbolicLayout lay{};
  const int status = tfem::symbolic_layout(n_el > 0 ? n_el : 1, n_dof, &lay);
  if (status != TFEM_OK) return status;
  *bytes = (int64_t)lay.total;
  return TFEM_OK;
}

extern "C" int tfem_csr_symbolic(int64_t n_el, const int32_t* dof_conn, int64_t n_dof, void* workspace, int64_t workspace_bytes, int32_t* crow,
                                 int32_t* col, int32_t* seg, int32_t* perm, int32_t* lin_seg, int32_t* lin_perm, int64_t* keys, int64_t* nnz,
                                 void* stream) {
  using namespace tfem;
  if (n_el <= 0 || n_dof <= 0) return TFEM_ERR_BAD_ARG;
  if (!dof_conn || !workspace || !crow || !col || !seg || !perm || !lin_seg || !lin_perm || !keys || !nnz) return TFEM_ERR_BAD_ARG;
  if (9 * n_el > kMaxIndex || n_dof > kMaxIndex) return TFEM_ERR_TOO_LARGE;
  SymbolicLayout lay{};
  int status = symbolic_layout(n_el, n_dof, &lay);
  if (status != TFEM_OK) return status;
  if ((size_t)workspace_bytes < lay.total) return TFEM_ERR_BAD_ARG;
  auto s = static_cast<cudaStream_t>(stream);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  int64_t* keys_in = reinterpret_cast<int64_t*>(ws + lay.keys_in);
  int64_t* keys_sorted = reinterpret_cast<int64_t*>(ws + lay.keys_sorted);
  int32_t* vals_in = reinterpret_cast<int32_t*>(ws + lay.vals_in);
  int32_t* counts = reinterpret_cast<int32_t*>(ws + lay.counts);
  int32_t* form_sorted = reinterpret_cast<int32_t*>(ws + lay.form_sorted);
  int32_t* lvals_in = reinterpret_cast<int32_t*>(ws + lay.lvals_in);
  void* cub_temp = ws + lay.cub_temp;
  size_t cub_bytes = lay.cub_bytes;
  const int64_t n9 = 9 * n_el, n3 = 3 * n_el;

  // ---- bilinear form: keys, stable sort with the entry index as payload, unique keys and their run lengths --------
  status = tfem_coo_keys(n_el, dof_conn, n_dof, keys_in, stream);
  if (status != TFEM_OK) return status;
  iota_kernel<<<blocks_for(n9, 256), 256, 0, s>>>(vals_in, n9);
  const int key_bits = bits_for(n_dof) * 2 < 63 ? bits_for(n_dof * n_dof) : 63;
  if (cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, keys_in, keys_sorted, vals_in, perm, (int)n9, 0, key_bits, s) != cudaSuccess)
    return TFEM_ERR_LAUNCH;
  if (cudaMemsetAsync(counts, 0, 4 * (size_t)(n9 + 1), s) != cudaSuccess) return TFEM_ERR_LAUNCH;
  cub_bytes = lay.cub_bytes;
  if (cub::DeviceRunLengthEncode::Encode(cub_temp, cub_bytes, keys_sorted, keys, counts, nnz, (int)n9, s) != cudaSuccess) return TFEM_ERR_LAUNCH;
  cub_bytes = lay.cub_bytes;
  // seg[i] = first COO entry of unique key i (i < nnz); the run lengths behind nnz are zero, split_keys closes seg[nnz]
  if (cub::DeviceScan::ExclusiveSum(cub_temp, cub_bytes, counts, seg, (int)(n9 + 1), s) != cudaSuccess) return TFEM_ERR_LAUNCH;
  split_keys_kernel<<<blocks_for(n9 + 1, 256), 256, 0, s>>>(keys, nnz, n_dof, n9, col, seg);
  lower_bound_kernel<int64_t><<<blocks_for(n_dof + 1, 256), 256, 0, s>>>(keys, nnz, 0, n_dof, n_dof, crow);

  // ---- linear form: the 3 n_el (element, local vertex) entries sorted by DOF --------------------------------------
  iota_kernel<<<blocks_for(n3, 256), 256, 0, s>>>(lvals_in, n3);
  cub_bytes = lay.cub_bytes;
  if (cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, dof_conn, form_sorted, lvals_in, lin_perm, (int)n3, 0, bits_for(n_dof), s) != cudaSuccess)
    return TFEM_ERR_LAUNCH;
  lower_bound_kernel<int32_t><<<blocks_for(n_dof + 1, 256), 256, 0, s>>>(form_sorted, nullptr, n3, n_dof, 1, lin_seg);
  return check_launch();
}
