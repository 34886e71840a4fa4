// fdk_symbolic.cuh -- one-time symbolic CSR pattern build on the device.
//
// Reproduces, bit for bit, the pattern the reference obtains with NumPy/SciPy:
//   key = row * n_cols + col (int64) over all nne^2 node pairs of all elements, np.unique
//   -> block row I holds the sorted distinct J sharing an element with I
//   (fedoo/core/_sparsematrix.py:257-284), then scipy.sparse.bmat tiles the block nvar x nvar
//   in variable-major order (:310-315): row v*n+I = concat_v' (v'*n + blockrow(I)).
// Sorting is a CUB radix sort restricted to the significant key bits; np.unique -> CUB unique.
#pragma once
#include <cub/cub.cuh>

#include "fdk_common.cuh"

namespace fdk {

__global__ void k_pair_keys(int64_t n_elems, int nne, int64_t n_nodes, const int32_t* __restrict__ conn,
                            uint64_t* __restrict__ keys) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = n_elems * nne * nne;
  if (t >= total) return;
  const int nn2 = nne * nne;
  const int64_t e = t / nn2;
  const int r = (int)(t - e * nn2);
  const int i = r / nne, j = r - i * nne;
  keys[t] = (uint64_t)conn[e * nne + i] * (uint64_t)n_nodes + (uint64_t)conn[e * nne + j];
}

__global__ void k_block_csr(int n_nodes, int64_t blk_nnz, const uint64_t* __restrict__ keys,
                            int64_t* __restrict__ indptr, int32_t* __restrict__ indices) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < blk_nnz) indices[t] = (int32_t)(keys[t] % (uint64_t)n_nodes);
  if (t <= n_nodes) {
    // indptr[I] = first position with key >= I * n_nodes
    const uint64_t target = (uint64_t)t * (uint64_t)n_nodes;
    int64_t lo = 0, hi = blk_nnz;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    indptr[t] = lo;
  }
}

template <typename IdxT>
__global__ void k_expand_indptr(int n_nodes, int nvar, int n_global, int64_t blk_nnz,
                                const int64_t* __restrict__ bptr, IdxT* __restrict__ indptr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_rows = (int64_t)nvar * n_nodes;
  if (t < n_rows) {
    const int64_t v = t / n_nodes, I = t - v * n_nodes;
    indptr[t] = (IdxT)(v * nvar * blk_nnz + (int64_t)nvar * bptr[I]);
  } else if (t <= n_rows + n_global) {
    indptr[t] = (IdxT)((int64_t)nvar * nvar * blk_nnz);
  }
}

// one thread per (block entry, v): writes the nvar copies of its column index
template <typename IdxT>
__global__ void k_expand_indices(int n_nodes, int nvar, int64_t blk_nnz, const int64_t* __restrict__ bptr,
                                 const int32_t* __restrict__ bidx, const int32_t* __restrict__ brow,
                                 IdxT* __restrict__ indices) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= blk_nnz) return;
  const int64_t I = brow[t];
  const int64_t b0 = bptr[I], deg = bptr[I + 1] - b0, pcol = t - b0;
  const int64_t J = bidx[t];
  for (int v = 0; v < nvar; ++v) {
    const int64_t base = (int64_t)v * nvar * blk_nnz + (int64_t)nvar * b0;
    for (int vp = 0; vp < nvar; ++vp) indices[base + (int64_t)vp * deg + pcol] = (IdxT)((int64_t)vp * n_nodes + J);
  }
}

__global__ void k_block_rows(int64_t blk_nnz, int64_t n_nodes, const uint64_t* __restrict__ keys,
                             int32_t* __restrict__ brow) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < blk_nnz) brow[t] = (int32_t)(keys[t] / (uint64_t)n_nodes);
}

inline int sym_block_keys(int n_nodes, int64_t n_elems, int nne, const int32_t* conn, uint64_t* keys_out,
                          int64_t* blk_nnz_h, cudaStream_t stream) {
  const int64_t total = n_elems * nne * nne;
  *blk_nnz_h = 0;
  if (total == 0) return 0;
  uint64_t* keys_in = nullptr;
  int64_t* d_num = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0, tmp2 = 0;
  int end_bit = 1;
  while (end_bit < 64 && ((uint64_t)n_nodes * (uint64_t)n_nodes) >> end_bit) ++end_bit;
  FDK_CUDA(cudaMallocAsync(&keys_in, (size_t)total * 8 * 2, stream));
  uint64_t* keys_alt = keys_in + total;
  FDK_CUDA(cudaMallocAsync(&d_num, 8, stream));
  k_pair_keys<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(n_elems, nne, n_nodes, conn, keys_in);
  FDK_CUDA(cudaGetLastError());
  cub::DoubleBuffer<uint64_t> db(keys_in, keys_alt);
  FDK_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, db, total, 0, end_bit, stream));
  FDK_CUDA(cub::DeviceSelect::Unique(nullptr, tmp2, keys_in, keys_out, d_num, total, stream));
  if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
  FDK_CUDA(cudaMallocAsync(&tmp, tmp_bytes, stream));
  FDK_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, db, total, 0, end_bit, stream));
  FDK_CUDA(cub::DeviceSelect::Unique(tmp, tmp_bytes, db.Current(), keys_out, d_num, total, stream));
  FDK_CUDA(cudaMemcpyAsync(blk_nnz_h, d_num, 8, cudaMemcpyDeviceToHost, stream));
  FDK_CUDA(cudaFreeAsync(tmp, stream));
  FDK_CUDA(cudaFreeAsync(d_num, stream));
  FDK_CUDA(cudaFreeAsync(keys_in, stream));
  FDK_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

inline int sym_block_csr(int n_nodes, int64_t blk_nnz, const uint64_t* keys, int64_t* indptr, int32_t* indices,
                         cudaStream_t stream) {
  const int64_t work = (blk_nnz > n_nodes + 1) ? blk_nnz : (int64_t)n_nodes + 1;
  k_block_csr<<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(n_nodes, blk_nnz, keys, indptr, indices);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void k_rows_from_ptr(int n_nodes, int64_t blk_nnz, const int64_t* __restrict__ bptr,
                                int32_t* __restrict__ brow) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= blk_nnz) return;
  int lo = 0, hi = n_nodes;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (bptr[mid] <= t) lo = mid; else hi = mid;
  }
  brow[t] = lo;
}

inline int sym_expand_csr(int n_nodes, int nvar, int n_global, int64_t blk_nnz, const int64_t* bptr,
                          const int32_t* bidx, int index_bytes, void* indptr, void* indices, cudaStream_t stream) {
  FDK_REQUIRE(index_bytes == 4 || index_bytes == 8, FDK_EINVAL, "index_bytes must be 4 or 8");
  const int64_t nnz = (int64_t)nvar * nvar * blk_nnz;
  const int64_t n_rows = (int64_t)nvar * n_nodes + n_global;
  if (index_bytes == 4)
    FDK_REQUIRE(nnz <= INT32_MAX && n_rows <= INT32_MAX, FDK_EOVERFLOW, "nnz %lld needs 64-bit indices",
                (long long)nnz);
  int32_t* brow = nullptr;  // block-row index of every block entry
  if (blk_nnz > 0) FDK_CUDA(cudaMallocAsync(&brow, (size_t)blk_nnz * 4, stream));
  const unsigned gb = (unsigned)((n_rows + 1 + 255) / 256);
  if (index_bytes == 4)
    k_expand_indptr<int32_t><<<gb, 256, 0, stream>>>(n_nodes, nvar, n_global, blk_nnz, bptr, (int32_t*)indptr);
  else
    k_expand_indptr<int64_t><<<gb, 256, 0, stream>>>(n_nodes, nvar, n_global, blk_nnz, bptr, (int64_t*)indptr);
  FDK_CUDA(cudaGetLastError());
  if (blk_nnz > 0) {
    k_rows_from_ptr<<<(unsigned)((blk_nnz + 255) / 256), 256, 0, stream>>>(n_nodes, blk_nnz, bptr, brow);
    FDK_CUDA(cudaGetLastError());
    const unsigned ge = (unsigned)((blk_nnz + 255) / 256);
    if (index_bytes == 4)
      k_expand_indices<int32_t><<<ge, 256, 0, stream>>>(n_nodes, nvar, blk_nnz, bptr, bidx, brow, (int32_t*)indices);
    else
      k_expand_indices<int64_t><<<ge, 256, 0, stream>>>(n_nodes, nvar, blk_nnz, bptr, bidx, brow, (int64_t*)indices);
    FDK_CUDA(cudaGetLastError());
    FDK_CUDA(cudaFreeAsync(brow, stream));
  }
  return 0;
}

}  // namespace fdk
