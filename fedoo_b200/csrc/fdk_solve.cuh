// fdk_solve.cuh -- what the callers of the assembly path do with K on the device:
// CSR sparse matrix-vector product with Dirichlet masking, diagonal extraction and a Jacobi-preconditioned
// conjugate gradient, so that Problem.solve() never has to bring the 23 GB matrix to the host.
//
// Reference: fedoo/core/problem.py:277-298 (elimination of the imposed dofs: MatCB^T A MatCB x = MatCB^T (B - A Xbc);
// for pure Dirichlet conditions MatCB selects the free dofs, which is what the mask does here without forming
// the product), fedoo/core/base.py:521-537 (scipy.sparse.linalg.cg with M = diag(1 / A.diagonal())).
//
// All kernels are HBM-bound streaming kernels: the SpMV reads 8 + index_bytes bytes per stored entry once
// (values and column indices with ld.global.nc, coalesced: one sub-warp per row) and gathers x through L2;
// the vector updates are fused so that every CG iteration reads / writes each vector once.  Reductions are
// two-stage with a fixed block order (no floating-point atomics): results are bit-reproducible.
#pragma once
#include <cstdlib>

#include "fdk_common.cuh"

namespace fdk {

constexpr int RED_BLOCKS = 1184;  // 8 x 148: partial sums of the dot products
constexpr int RED_THREADS = 256;

// MASK_COLS = false: the caller guarantees x == 0 on the imposed dofs (the CG direction is), only rows are masked.
template <class Idx, int LPR, bool MASK_COLS>  // LPR lanes per row (power of two <= 32)
__global__ void __launch_bounds__(256) k_csr_spmv(int64_t n_rows, const Idx* __restrict__ indptr,
                                                   const Idx* __restrict__ indices, const double* __restrict__ data,
                                                   const double* __restrict__ x, const unsigned char* __restrict__ mask,
                                                   double* __restrict__ y) {
  constexpr int RPW = 32 / LPR;  // rows per warp
  constexpr int UN = 4;          // entries per lane in flight: all index / value loads, then all gathers
  const int lane = threadIdx.x & (LPR - 1);
  const int sub = (threadIdx.x & 31) / LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t rb = warp * RPW; rb < n_rows; rb += n_warps * RPW) {  // warp-uniform trip count (shuffles below)
    const int64_t r = rb + sub;
    double s = 0.0;
    const bool live = r < n_rows && (mask == nullptr || mask[r]);
    if (live) {
      const int64_t e0 = __ldg(indptr + r), e1 = __ldg(indptr + r + 1);
      for (int64_t eb = e0 + lane; eb < e1; eb += UN * LPR) {
        int64_t c[UN];
        double v[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int64_t e = eb + u * LPR;
          const bool in = e < e1;
          c[u] = in ? (int64_t)__ldg(indices + e) : -1;
          v[u] = in ? __ldg(data + e) : 0.0;
        }
        double xv[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          bool take = c[u] >= 0;
          if (MASK_COLS) take = take && (mask == nullptr || mask[c[u]]);
          xv[u] = take ? __ldg(x + c[u]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) s = fma(v[u], xv[u], s);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, LPR);
    if (lane == 0 && r < n_rows) y[r] = s;
  }
}

// The same product for the TILED pattern of the assembly (scipy.sparse.bmat layout: row v n + I = concat over v' of
// v' n + blockrow(I), values at v NV blk_nnz + NV bp(I) + v' deg(I) + pcol): one sub-warp per NODE row reads the block
// row's column list once for its NV x NV scalar rows -- the generic kernel reads each column index NV x NV times, a
// third of its traffic -- and gathers the NV dofs of every column node once instead of NV times.
template <int NV, int LPR, bool MASK_COLS>
__global__ void __launch_bounds__(256) k_bcsr_spmv(int n_nodes, int64_t blk_nnz, const int64_t* __restrict__ blk_indptr,
                                                    const int32_t* __restrict__ blk_indices,
                                                    const double* __restrict__ data, const double* __restrict__ x,
                                                    const unsigned char* __restrict__ mask, double* __restrict__ y) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & (LPR - 1);
  const int sub = (threadIdx.x & 31) / LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t rb = warp * RPW; rb < n_nodes; rb += n_warps * RPW) {
    const int64_t I = rb + sub;
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = 0.0;
    if (I < n_nodes) {
      const int64_t e0 = __ldg(blk_indptr + I), e1 = __ldg(blk_indptr + I + 1);
      const int64_t deg = e1 - e0;
      for (int64_t pc = lane; pc < deg; pc += LPR) {
        const int64_t J = __ldg(blk_indices + e0 + pc);
        double xj[NV];
#pragma unroll
        for (int w = 0; w < NV; ++w) {
          const int64_t c = (int64_t)w * n_nodes + J;
          const bool take = !MASK_COLS || mask == nullptr || mask[c];
          xj[w] = take ? __ldg(x + c) : 0.0;
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const double* row = data + ((int64_t)v * NV * blk_nnz + (int64_t)NV * e0 + pc);
#pragma unroll
          for (int w = 0; w < NV; ++w) acc[v] = fma(__ldg(row + (int64_t)w * deg), xj[w], acc[v]);
        }
      }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double s = acc[v];
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, LPR);
      if (lane == 0 && I < n_nodes) {
        const int64_t r = (int64_t)v * n_nodes + I;
        y[r] = (mask == nullptr || mask[r]) ? s : 0.0;
      }
    }
  }
}

struct BlockPattern {  // optional tiled-pattern description of a CSR matrix assembled by this library
  int n_nodes = 0, nvar = 0;
  int64_t blk_nnz = 0;
  const int64_t* blk_indptr = nullptr;
  const int32_t* blk_indices = nullptr;
};

template <bool MASK_COLS>
int launch_bspmv(const BlockPattern& b, const double* data, const double* x, const unsigned char* mask, double* y,
                 cudaStream_t stream) {
  if (b.n_nodes == 0) return 0;
  const int threads = 256;
  const double avg = (double)b.blk_nnz / (double)b.n_nodes;
  int lpr = avg >= 160 ? 32 : avg >= 80 ? 16 : avg >= 40 ? 8 : 4;  // measured at 27 blocks per row: 4.1 ms with 8 lanes,
  if (const char* e = getenv("FDK_BSPMV_LANES")) lpr = atoi(e);      // 4.7 with 16, 6.3 with 32 (more rows in flight)
  const int64_t need = ((int64_t)b.n_nodes * lpr + threads - 1) / threads;
  const unsigned grid = (unsigned)(need < 148 * 32 ? need : 148 * 32);
#define FDK_BSPMV(NV_, LPR_)                                                                                       \
  k_bcsr_spmv<NV_, LPR_, MASK_COLS><<<grid, threads, 0, stream>>>(b.n_nodes, b.blk_nnz, b.blk_indptr, b.blk_indices, \
                                                                   data, x, mask, y)
  if (b.nvar == 3) {
    if (lpr == 32) FDK_BSPMV(3, 32); else if (lpr == 16) FDK_BSPMV(3, 16); else if (lpr == 8) FDK_BSPMV(3, 8); else FDK_BSPMV(3, 4);
  } else if (b.nvar == 2) {
    if (lpr == 32) FDK_BSPMV(2, 32); else if (lpr == 16) FDK_BSPMV(2, 16); else if (lpr == 8) FDK_BSPMV(2, 8); else FDK_BSPMV(2, 4);
  } else if (b.nvar == 1) {
    if (lpr == 32) FDK_BSPMV(1, 32); else if (lpr == 16) FDK_BSPMV(1, 16); else if (lpr == 8) FDK_BSPMV(1, 8); else FDK_BSPMV(1, 4);
  } else {
    set_error("tiled SpMV: nvar must be 1, 2 or 3 (got %d)", b.nvar);
    return FDK_EINVAL;
  }
#undef FDK_BSPMV
  FDK_CUDA(cudaGetLastError());
  return 0;
}

template <class Idx>
__global__ void k_csr_diagonal(int64_t n_rows, const Idx* __restrict__ indptr, const Idx* __restrict__ indices,
                               const double* __restrict__ data, double* __restrict__ diag) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  // sorted columns (scipy canonical form): binary search of the diagonal entry
  int64_t lo = indptr[r], hi = indptr[r + 1];
  double d = 0.0;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t c = indices[mid];
    if (c == r) {
      d = data[mid];
      break;
    }
    if (c < r) lo = mid + 1;
    else hi = mid;
  }
  diag[r] = d;
}

// block-level sum in a fixed order; the result is valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[RED_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < RED_THREADS / 32; ++w) s += sh[w];
  }
  __syncthreads();
  return s;
}

// scal[k] = sum of the RED_BLOCKS partials part[k][*], k < n_sums (one block)
__global__ void __launch_bounds__(RED_THREADS) k_reduce_final(const double* __restrict__ part, int n_sums,
                                                              double* __restrict__ scal) {
  for (int k = 0; k < n_sums; ++k) {
    double v = 0.0;
    for (int i = threadIdx.x; i < RED_BLOCKS; i += RED_THREADS) v += part[k * RED_BLOCKS + i];
    const double s = block_sum(v);
    if (threadIdx.x == 0) scal[k] = s;
  }
}

// scalars on the device: 0 rz, 1 pq, 2 rz_new, 3 rr, 4 bb
enum { S_RZ = 0, S_PQ = 1, S_RZN = 2, S_RR = 3, S_BB = 4, S_COUNT = 8 };

// r = b (masked), z = dinv r, p = z; partials of rz, rr
__global__ void __launch_bounds__(RED_THREADS) k_pcg_init(int64_t n, const double* __restrict__ b,
                                                          const double* __restrict__ diag,
                                                          const unsigned char* __restrict__ mask, double* __restrict__ x,
                                                          double* __restrict__ r, double* __restrict__ z,
                                                          double* __restrict__ p, double* __restrict__ part) {
  double rz = 0.0, rr = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool live = mask == nullptr || mask[i];
    const double ri = live ? b[i] : 0.0;
    const double di = diag[i];
    const double zi = (live && di != 0.0) ? ri / di : 0.0;
    x[i] = 0.0;
    r[i] = ri;
    z[i] = zi;
    p[i] = zi;
    rz = fma(ri, zi, rz);
    rr = fma(ri, ri, rr);
  }
  const double a = block_sum(rz), c = block_sum(rr);
  if (threadIdx.x == 0) {
    part[0 * RED_BLOCKS + blockIdx.x] = a;
    part[1 * RED_BLOCKS + blockIdx.x] = c;
  }
}

__global__ void __launch_bounds__(RED_THREADS) k_dot(int64_t n, const double* __restrict__ a,
                                                     const double* __restrict__ b, double* __restrict__ part) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s = fma(a[i], b[i], s);
  const double v = block_sum(s);
  if (threadIdx.x == 0) part[blockIdx.x] = v;
}

// alpha = rz / pq; x += alpha p; r -= alpha q; z = dinv r; partials of rz_new, rr
__global__ void __launch_bounds__(RED_THREADS) k_pcg_update(int64_t n, const double* __restrict__ scal,
                                                            const double* __restrict__ diag,
                                                            const unsigned char* __restrict__ mask,
                                                            const double* __restrict__ p, const double* __restrict__ q,
                                                            double* __restrict__ x, double* __restrict__ r,
                                                            double* __restrict__ z, double* __restrict__ part) {
  const double pq = scal[S_PQ];
  const double alpha = pq != 0.0 ? scal[S_RZ] / pq : 0.0;
  double rz = 0.0, rr = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool live = mask == nullptr || mask[i];
    const double ri = live ? fma(-alpha, q[i], r[i]) : 0.0;
    const double di = diag[i];
    const double zi = (live && di != 0.0) ? ri / di : 0.0;
    x[i] = fma(alpha, p[i], x[i]);
    r[i] = ri;
    z[i] = zi;
    rz = fma(ri, zi, rz);
    rr = fma(ri, ri, rr);
  }
  const double a = block_sum(rz), c = block_sum(rr);
  if (threadIdx.x == 0) {
    part[0 * RED_BLOCKS + blockIdx.x] = a;
    part[1 * RED_BLOCKS + blockIdx.x] = c;
  }
}

// beta = rz_new / rz; p = z + beta p; then rz <- rz_new (thread 0 of block 0, after everyone has read it:
// the swap is done by the NEXT launch reading S_RZN, see pcg_jacobi)
__global__ void __launch_bounds__(RED_THREADS) k_pcg_direction(int64_t n, const double* __restrict__ scal,
                                                               const double* __restrict__ z, double* __restrict__ p) {
  const double rz = scal[S_RZ];
  const double beta = rz != 0.0 ? scal[S_RZN] / rz : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = fma(beta, p[i], z[i]);
}

__global__ void k_copy_scalar(double* scal, int dst, int src) { scal[dst] = scal[src]; }

template <class Idx, bool MASK_COLS>
int launch_spmv(int64_t n_rows, const Idx* indptr, const Idx* indices, const double* data, const double* x,
                const unsigned char* mask, double* y, int lanes_per_row, cudaStream_t stream) {
  if (n_rows == 0) return 0;
  const int threads = 256;
  auto grid_for = [&](int lpr) {
    const int64_t need = (n_rows * lpr + threads - 1) / threads;
    const int64_t cap = 148 * 32;  // grid-stride beyond a few waves
    return (unsigned)(need < cap ? need : cap);
  };
  switch (lanes_per_row) {
    case 32: k_csr_spmv<Idx, 32, MASK_COLS><<<grid_for(32), threads, 0, stream>>>(n_rows, indptr, indices, data, x, mask, y); break;
    case 16: k_csr_spmv<Idx, 16, MASK_COLS><<<grid_for(16), threads, 0, stream>>>(n_rows, indptr, indices, data, x, mask, y); break;
    case 8: k_csr_spmv<Idx, 8, MASK_COLS><<<grid_for(8), threads, 0, stream>>>(n_rows, indptr, indices, data, x, mask, y); break;
    default: k_csr_spmv<Idx, 4, MASK_COLS><<<grid_for(4), threads, 0, stream>>>(n_rows, indptr, indices, data, x, mask, y); break;
  }
  FDK_CUDA(cudaGetLastError());
  return 0;
}

inline int pick_lanes(int64_t n_rows, int64_t nnz) {
  const double avg = n_rows > 0 ? (double)nnz / (double)n_rows : 0.0;
  // a lane keeps 4 entries in flight: 81-entry elasticity rows fit one pass of 32 lanes, 27-entry rows one of 8
  return avg >= 64 ? 32 : avg >= 32 ? 16 : avg >= 16 ? 8 : 4;
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-point constraints of the periodic boundary conditions (fedoo/constraint/periodic_bc.py:910-1800 builds them as
// MPC objects, fedoo/core/problem.py:277-298 eliminates them through MatCB^T A MatCB):
//   x[slave_s] = x[master_s] + sum_k coef[s][k] * x[n_nodal + k]          (k < n_glob trailing global dofs)
// The reduced operator T^T A T is applied matrix-free on full-length vectors whose slave entries are kept at zero:
// expand (x_s from masters and globals), SpMV, reduce (rows of the slaves folded into their master and into the
// global dofs), zero the slaves.  Reductions run in a fixed order (slaves grouped by master, one block per global dof).
struct MpcMap {
  int64_t n_nodal;
  int n_glob;
  int64_t n_slave;
  const int* slave;     // (n_slave) dof index < n_nodal
  const int* master;    // (n_slave) dof index < n_nodal
  const double* coef;   // (n_slave, n_glob) row-major
  int64_t n_master;     // distinct master dofs
  const int* mst_dof;   // (n_master)
  const int* mst_ptr;   // (n_master + 1) into mst_slv
  const int* mst_slv;   // slave ordinals grouped by master
  double* scratch;      // (MPC_MAX_GLOB * MPC_MAX_GLOB * MPC_FOLD_BLOCKS) partial sums of the fold
};

__global__ void k_mpc_expand(MpcMap m, double* __restrict__ x) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < m.n_slave; s += (int64_t)gridDim.x * blockDim.x) {
    double v = x[m.master[s]];
    for (int k = 0; k < m.n_glob; ++k) v = fma(m.coef[s * m.n_glob + k], x[m.n_nodal + k], v);
    x[m.slave[s]] = v;
  }
}

__global__ void k_mpc_fold_master(MpcMap m, double* __restrict__ q) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m.n_master; i += (int64_t)gridDim.x * blockDim.x) {
    double acc = q[m.mst_dof[i]];
    for (int t = m.mst_ptr[i]; t < m.mst_ptr[i + 1]; ++t) acc += q[m.slave[m.mst_slv[t]]];
    q[m.mst_dof[i]] = acc;
  }
}

// q[n_nodal + k] = sum_s coef[s][k]^(1|2) q[slave_s], two stages in a fixed order: MPC_FOLD_BLOCKS partial sums per
// global dof in the caller's scratch, then one block adds them up
constexpr int MPC_FOLD_BLOCKS = 148;
constexpr int MPC_MAX_GLOB = 9;

template <bool SQUARE>
__global__ void __launch_bounds__(RED_THREADS) k_mpc_fold_glob_part(MpcMap m, const double* __restrict__ q,
                                                                    double* __restrict__ scratch) {
  double acc[MPC_MAX_GLOB];
#pragma unroll
  for (int k = 0; k < MPC_MAX_GLOB; ++k) acc[k] = 0.0;
  for (int64_t s = (int64_t)blockIdx.x * RED_THREADS + threadIdx.x; s < m.n_slave; s += (int64_t)gridDim.x * RED_THREADS) {
    const double v = q[m.slave[s]];
    const double* c = m.coef + s * m.n_glob;
#pragma unroll
    for (int k = 0; k < MPC_MAX_GLOB; ++k)
      if (k < m.n_glob) acc[k] = fma(SQUARE ? c[k] * c[k] : c[k], v, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < MPC_MAX_GLOB; ++k) {
    if (k < m.n_glob) {
      const double v = block_sum(acc[k]);
      if (threadIdx.x == 0) scratch[k * MPC_FOLD_BLOCKS + blockIdx.x] = v;
    }
  }
}

__global__ void __launch_bounds__(RED_THREADS) k_mpc_fold_glob_final(MpcMap m, const double* __restrict__ scratch,
                                                                     double* __restrict__ q) {
  for (int k = 0; k < m.n_glob; ++k) {
    const double v = block_sum(threadIdx.x < MPC_FOLD_BLOCKS ? scratch[k * MPC_FOLD_BLOCKS + threadIdx.x] : 0.0);
    if (threadIdx.x == 0) q[m.n_nodal + k] = v;
  }
}

__global__ void k_mpc_zero_slaves(MpcMap m, double* __restrict__ a, double* __restrict__ b) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < m.n_slave; s += (int64_t)gridDim.x * blockDim.x) {
    a[m.slave[s]] = 0.0;
    if (b != nullptr) b[m.slave[s]] = 0.0;
  }
}

inline unsigned mpc_grid(int64_t n) {
  const int64_t need = (n + 255) / 256;
  return (unsigned)(need < 1 ? 1 : need > 148 * 8 ? 148 * 8 : need);
}

// q <- T^T q (q holds A x on the nodal rows; the global rows are overwritten), then the slave entries of q (and of
// ``also``: the search direction that k_mpc_expand filled) are cleared
inline int mpc_fold(const MpcMap& m, double* q, double* also, bool square, cudaStream_t stream) {
  if (m.n_slave > 0) k_mpc_fold_master<<<mpc_grid(m.n_master), 256, 0, stream>>>(m, q);
  if (m.n_glob > 0) {
    if (square) k_mpc_fold_glob_part<true><<<MPC_FOLD_BLOCKS, RED_THREADS, 0, stream>>>(m, q, m.scratch);
    else k_mpc_fold_glob_part<false><<<MPC_FOLD_BLOCKS, RED_THREADS, 0, stream>>>(m, q, m.scratch);
    k_mpc_fold_glob_final<<<1, RED_THREADS, 0, stream>>>(m, m.scratch, q);
  }
  if (m.n_slave > 0) k_mpc_zero_slaves<<<mpc_grid(m.n_slave), 256, 0, stream>>>(m, q, also);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

inline int mpc_expand(const MpcMap& m, double* x, cudaStream_t stream) {
  if (m.n_slave > 0) k_mpc_expand<<<mpc_grid(m.n_slave), 256, 0, stream>>>(m, x);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

// Jacobi-PCG on the free dofs.  work: 5 n doubles (r, z, p, q, diag) + (2 RED_BLOCKS + S_COUNT) doubles.
// With ``mpc`` the vectors carry n + n_glob entries, ``mask`` marks the independent free dofs (0 on imposed dofs AND on
// slaves), b must already be folded (T^T b) and x returns the independent dofs (expand it afterwards).
template <class Idx>
int pcg_jacobi(int64_t n_rows, int64_t nnz, const Idx* indptr, const Idx* indices, const double* data, const double* b,
               double* x, const unsigned char* mask, double rtol, int max_iter, int check_every, double* work,
               int* iters_h, double* relres_h, cudaStream_t stream, const BlockPattern* blk = nullptr,
               const MpcMap* mpc = nullptr) {
  const int64_t n = n_rows + (mpc != nullptr ? mpc->n_glob : 0);
  const unsigned char* row_mask = mpc != nullptr ? nullptr : mask;  // slave rows are needed, the vector kernels mask
  double* r = work;
  double* z = r + n;
  double* p = z + n;
  double* q = p + n;
  double* diag = q + n;
  double* part = diag + n;
  double* scal = part + 2 * RED_BLOCKS;
  const int lanes = pick_lanes(n_rows, nnz);
  k_csr_diagonal<Idx><<<(unsigned)((n_rows + 255) / 256), 256, 0, stream>>>(n_rows, indptr, indices, data, diag);
  if (mpc != nullptr) {  // diag(T^T A T) without the cross terms A_ms (a preconditioner only needs an SPD diagonal)
    if (int rc = mpc_fold(*mpc, diag, nullptr, true, stream)) return rc;
  }
  k_pcg_init<<<RED_BLOCKS, RED_THREADS, 0, stream>>>(n, b, diag, mask, x, r, z, p, part);
  k_reduce_final<<<1, RED_THREADS, 0, stream>>>(part, 2, scal + S_RZN);  // S_RZN = rz, S_RR = rr
  k_copy_scalar<<<1, 1, 0, stream>>>(scal, S_RZ, S_RZN);
  k_copy_scalar<<<1, 1, 0, stream>>>(scal, S_BB, S_RR);
  FDK_CUDA(cudaGetLastError());
  double h[S_COUNT];
  FDK_CUDA(cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, stream));
  FDK_CUDA(cudaStreamSynchronize(stream));
  const double bb = h[S_BB];
  int it = 0;
  double rr = bb;
  if (bb > 0.0) {
    const double target = rtol * rtol * bb;
    while (it < max_iter) {
      // p vanishes on the imposed dofs by construction: only the rows need the mask
      if (mpc != nullptr) {
        if (int rc = mpc_expand(*mpc, p, stream)) return rc;
      }
      if (blk != nullptr) {
        if (int rc = launch_bspmv<false>(*blk, data, p, row_mask, q, stream)) return rc;
      } else if (int rc = launch_spmv<Idx, false>(n_rows, indptr, indices, data, p, row_mask, q, lanes, stream)) {
        return rc;
      }
      if (mpc != nullptr) {
        if (int rc = mpc_fold(*mpc, q, p, false, stream)) return rc;
      }
      k_dot<<<RED_BLOCKS, RED_THREADS, 0, stream>>>(n, p, q, part);
      k_reduce_final<<<1, RED_THREADS, 0, stream>>>(part, 1, scal + S_PQ);
      k_pcg_update<<<RED_BLOCKS, RED_THREADS, 0, stream>>>(n, scal, diag, mask, p, q, x, r, z, part);
      k_reduce_final<<<1, RED_THREADS, 0, stream>>>(part, 2, scal + S_RZN);
      k_pcg_direction<<<RED_BLOCKS, RED_THREADS, 0, stream>>>(n, scal, z, p);
      k_copy_scalar<<<1, 1, 0, stream>>>(scal, S_RZ, S_RZN);
      ++it;
      if (it % check_every == 0 || it == max_iter) {
        FDK_CUDA(cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, stream));
        FDK_CUDA(cudaStreamSynchronize(stream));
        rr = h[S_RR];
        if (!(rr > target)) break;  // also leaves on NaN
      }
    }
    FDK_CUDA(cudaGetLastError());
  }
  if (iters_h) *iters_h = it;
  if (relres_h) *relres_h = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// R right-hand sides in lockstep (the six load cases of fd.homogen.get_tangent_stiffness share one K,
// fedoo/homogen/tangent_stiffness.py:97-153): vectors are interleaved [n][R], every iteration reads K ONCE for the R
// products (the SpMV is the HBM-bound part of the loop) and runs R independent CG recurrences with per-column scalars.
// ---------------------------------------------------------------------------------------------------------------
template <int NV, int LPR, int R, int MINB>
__global__ void __launch_bounds__(256, MINB) k_bcsr_spmm(int n_nodes, int64_t blk_nnz, const int64_t* __restrict__ blk_indptr,
                                                    const int32_t* __restrict__ blk_indices,
                                                    const double* __restrict__ data, const double* __restrict__ x,
                                                    double* __restrict__ y) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & (LPR - 1);
  const int sub = (threadIdx.x & 31) / LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t rb = warp * RPW; rb < n_nodes; rb += n_warps * RPW) {
    const int64_t I = rb + sub;
    double acc[NV][R];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < R; ++k) acc[v][k] = 0.0;
    if (I < n_nodes) {
      const int64_t e0 = __ldg(blk_indptr + I), e1 = __ldg(blk_indptr + I + 1);
      const int64_t deg = e1 - e0;
      for (int64_t pc = lane; pc < deg; pc += LPR) {
        const int64_t J = __ldg(blk_indices + e0 + pc);
#pragma unroll
        for (int w = 0; w < NV; ++w) {
          const double* xr = x + ((int64_t)w * n_nodes + J) * R;
          double xj[R];
          if constexpr (R % 2 == 0) {  // rows of R doubles are 16-byte aligned
#pragma unroll
            for (int k = 0; k < R; k += 2) {
              const double2 t = __ldg(reinterpret_cast<const double2*>(xr + k));
              xj[k] = t.x;
              xj[k + 1] = t.y;
            }
          } else {
#pragma unroll
            for (int k = 0; k < R; ++k) xj[k] = __ldg(xr + k);
          }
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const double a = __ldg(data + ((int64_t)v * NV * blk_nnz + (int64_t)NV * e0 + pc) + (int64_t)w * deg);
#pragma unroll
            for (int k = 0; k < R; ++k) acc[v][k] = fma(a, xj[k], acc[v][k]);
          }
        }
      }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        double s = acc[v][k];
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, LPR);
        if (lane == 0 && I < n_nodes) y[((int64_t)v * n_nodes + I) * R + k] = s;
      }
    }
  }
}

template <int R>
int launch_bspmm(const BlockPattern& b, const double* data, const double* x, double* y, cudaStream_t stream) {
  if (b.n_nodes == 0) return 0;
  const int threads = 256;
  const double avg = (double)b.blk_nnz / (double)b.n_nodes;
  int lpr = avg >= 80 ? 16 : avg >= 40 ? 8 : 4;
  if (const char* e = getenv("FDK_SPMM_LANES")) lpr = atoi(e);
  int minb = 2;
  if (const char* e = getenv("FDK_SPMM_MINB")) minb = atoi(e);
  const int64_t need = ((int64_t)b.n_nodes * lpr + threads - 1) / threads;
  const unsigned grid = (unsigned)(need < 148 * 32 ? need : 148 * 32);
#define FDK_BSPMM_(NV_, LPR_, MB_) \
  k_bcsr_spmm<NV_, LPR_, R, MB_><<<grid, threads, 0, stream>>>(b.n_nodes, b.blk_nnz, b.blk_indptr, b.blk_indices, data, x, y)
#define FDK_BSPMM(NV_, LPR_)                 \
  do {                                       \
    if (minb >= 3) FDK_BSPMM_(NV_, LPR_, 3); \
    else FDK_BSPMM_(NV_, LPR_, 2);           \
  } while (0)
  if (b.nvar == 3) {
    if (lpr == 16) FDK_BSPMM(3, 16); else if (lpr == 8) FDK_BSPMM(3, 8); else FDK_BSPMM(3, 4);
  } else if (b.nvar == 2) {
    if (lpr == 16) FDK_BSPMM(2, 16); else if (lpr == 8) FDK_BSPMM(2, 8); else FDK_BSPMM(2, 4);
  } else if (b.nvar == 1) {
    if (lpr == 16) FDK_BSPMM(1, 16); else if (lpr == 8) FDK_BSPMM(1, 8); else FDK_BSPMM(1, 4);
  } else {
    set_error("tiled SpMM: nvar must be 1, 2 or 3 (got %d)", b.nvar);
    return FDK_EINVAL;
  }
#undef FDK_BSPMM_
#undef FDK_BSPMM
  FDK_CUDA(cudaGetLastError());
  return 0;
}

// constraint map on interleaved vectors
template <int R>
__global__ void k_mpc_expand_multi(MpcMap m, double* __restrict__ x) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < m.n_slave; s += (int64_t)gridDim.x * blockDim.x) {
    double v[R];
    const double* xm = x + (int64_t)m.master[s] * R;
#pragma unroll
    for (int k = 0; k < R; ++k) v[k] = xm[k];
    for (int g = 0; g < m.n_glob; ++g) {
      const double c = m.coef[s * m.n_glob + g];
      const double* xg = x + (m.n_nodal + g) * R;
#pragma unroll
      for (int k = 0; k < R; ++k) v[k] = fma(c, xg[k], v[k]);
    }
    double* xs = x + (int64_t)m.slave[s] * R;
#pragma unroll
    for (int k = 0; k < R; ++k) xs[k] = v[k];
  }
}

template <int R>
__global__ void k_mpc_fold_master_multi(MpcMap m, double* __restrict__ q) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m.n_master; i += (int64_t)gridDim.x * blockDim.x) {
    double acc[R];
    double* qm = q + (int64_t)m.mst_dof[i] * R;
#pragma unroll
    for (int k = 0; k < R; ++k) acc[k] = qm[k];
    for (int t = m.mst_ptr[i]; t < m.mst_ptr[i + 1]; ++t) {
      const double* qs = q + (int64_t)m.slave[m.mst_slv[t]] * R;
#pragma unroll
      for (int k = 0; k < R; ++k) acc[k] += qs[k];
    }
#pragma unroll
    for (int k = 0; k < R; ++k) qm[k] = acc[k];
  }
}

// one global dof g per blockIdx.y: scratch[(g R + k) MPC_FOLD_BLOCKS + blockIdx.x]
template <int R>
__global__ void __launch_bounds__(RED_THREADS) k_mpc_fold_glob_part_multi(MpcMap m, const double* __restrict__ q,
                                                                          double* __restrict__ scratch) {
  const int g = blockIdx.y;
  double acc[R];
#pragma unroll
  for (int k = 0; k < R; ++k) acc[k] = 0.0;
  for (int64_t s = (int64_t)blockIdx.x * RED_THREADS + threadIdx.x; s < m.n_slave; s += (int64_t)gridDim.x * RED_THREADS) {
    const double c = m.coef[s * m.n_glob + g];
    const double* qs = q + (int64_t)m.slave[s] * R;
#pragma unroll
    for (int k = 0; k < R; ++k) acc[k] = fma(c, qs[k], acc[k]);
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const double v = block_sum(acc[k]);
    if (threadIdx.x == 0) scratch[((int64_t)g * R + k) * MPC_FOLD_BLOCKS + blockIdx.x] = v;
  }
}

template <int R>
__global__ void __launch_bounds__(RED_THREADS) k_mpc_fold_glob_final_multi(MpcMap m, const double* __restrict__ scratch,
                                                                           double* __restrict__ q) {
  const int j = blockIdx.x;  // g R + k
  const double v = block_sum(threadIdx.x < MPC_FOLD_BLOCKS ? scratch[(int64_t)j * MPC_FOLD_BLOCKS + threadIdx.x] : 0.0);
  if (threadIdx.x == 0) q[(m.n_nodal + j / R) * R + j % R] = v;
}

template <int R>
__global__ void k_mpc_zero_slaves_multi(MpcMap m, double* __restrict__ a, double* __restrict__ b) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m.n_slave * R; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = (int64_t)m.slave[t / R] * R + t % R;
    a[i] = 0.0;
    if (b != nullptr) b[i] = 0.0;
  }
}

template <int R>
int mpc_expand_multi(const MpcMap& m, double* x, cudaStream_t stream) {
  if (m.n_slave > 0) k_mpc_expand_multi<R><<<mpc_grid(m.n_slave), 256, 0, stream>>>(m, x);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

template <int R>
int mpc_fold_multi(const MpcMap& m, double* q, double* also, cudaStream_t stream) {
  if (m.n_slave > 0) k_mpc_fold_master_multi<R><<<mpc_grid(m.n_master), 256, 0, stream>>>(m, q);
  if (m.n_glob > 0) {
    k_mpc_fold_glob_part_multi<R><<<dim3(MPC_FOLD_BLOCKS, m.n_glob), RED_THREADS, 0, stream>>>(m, q, m.scratch);
    k_mpc_fold_glob_final_multi<R><<<m.n_glob * R, RED_THREADS, 0, stream>>>(m, m.scratch, q);
  }
  if (m.n_slave > 0) k_mpc_zero_slaves_multi<R><<<mpc_grid(m.n_slave * R), 256, 0, stream>>>(m, q, also);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

// per-column scalars: scal[k * S_COUNT + S_xx]; partial sums part[(j R + k) RED_BLOCKS + block], j = which sum.
// The vector kernels walk the interleaved arrays flat (fully coalesced) with MULTI_BLOCKS x RED_THREADS threads: that
// count is a multiple of 3 and 6, so a thread keeps its column (flat index % R) along its grid-stride loop.
constexpr int MULTI_BLOCKS = 1182;
static_assert((MULTI_BLOCKS * RED_THREADS) % 6 == 0 && MULTI_BLOCKS <= RED_BLOCKS, "column-stable grid");

// sums of v[s] over the threads of the block that share a column -> part[(s R + column) RED_BLOCKS + block]
template <int R, int NS>
__device__ __forceinline__ void block_column_sums(const double (&v)[NS], double* __restrict__ part) {
  __shared__ double sh[NS][RED_THREADS];
#pragma unroll
  for (int s = 0; s < NS; ++s) sh[s][threadIdx.x] = v[s];
  __syncthreads();
  if (threadIdx.x < R) {
    const int shift = (int)(((int64_t)blockIdx.x * RED_THREADS) % R);
    const int first = (threadIdx.x - shift + R) % R;  // first thread of the block working on column threadIdx.x
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      double acc = 0.0;
      for (int t = first; t < RED_THREADS; t += R) acc += sh[s][t];
      part[((int64_t)s * R + threadIdx.x) * RED_BLOCKS + blockIdx.x] = acc;
    }
  }
  __syncthreads();
}

// one block per (sum j, column k)
template <int R>
__global__ void __launch_bounds__(RED_THREADS) k_reduce_final_multi(const double* __restrict__ part, int first,
                                                                    double* __restrict__ scal) {
  const int j = blockIdx.x / R, k = blockIdx.x % R;
  double v = 0.0;
  for (int i = threadIdx.x; i < MULTI_BLOCKS; i += RED_THREADS) v += part[((int64_t)j * R + k) * RED_BLOCKS + i];
  const double s = block_sum(v);
  if (threadIdx.x == 0) scal[k * S_COUNT + first + j] = s;
}

template <int R>
__global__ void __launch_bounds__(RED_THREADS) k_pcg_init_multi(int64_t n, const double* __restrict__ b,
                                                                const double* __restrict__ diag,
                                                                const unsigned char* __restrict__ mask,
                                                                double* __restrict__ x, double* __restrict__ r,
                                                                double* __restrict__ z, double* __restrict__ p,
                                                                double* __restrict__ part) {
  double acc[2] = {0.0, 0.0};
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n * R; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = j / R;
    const bool live = mask == nullptr || mask[i];
    const double di = diag[i];
    const double ri = live ? b[j] : 0.0;
    const double zi = (live && di != 0.0) ? ri / di : 0.0;
    x[j] = 0.0;
    r[j] = ri;
    z[j] = zi;
    p[j] = zi;
    acc[0] = fma(ri, zi, acc[0]);
    acc[1] = fma(ri, ri, acc[1]);
  }
  block_column_sums<R, 2>(acc, part);
}

template <int R>
__global__ void __launch_bounds__(RED_THREADS) k_dot_multi(int64_t n, const double* __restrict__ a,
                                                           const double* __restrict__ b, double* __restrict__ part) {
  double acc[1] = {0.0};
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n * R; j += (int64_t)gridDim.x * blockDim.x)
    acc[0] = fma(a[j], b[j], acc[0]);
  block_column_sums<R, 1>(acc, part);
}

template <int R>
__global__ void __launch_bounds__(RED_THREADS) k_pcg_update_multi(int64_t n, const double* __restrict__ scal,
                                                                  const double* __restrict__ diag,
                                                                  const unsigned char* __restrict__ mask,
                                                                  const double* __restrict__ p, const double* __restrict__ q,
                                                                  double* __restrict__ x, double* __restrict__ r,
                                                                  double* __restrict__ z, double* __restrict__ part) {
  const int k = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) % R);
  const double pq = scal[k * S_COUNT + S_PQ];
  const double alpha = pq != 0.0 ? scal[k * S_COUNT + S_RZ] / pq : 0.0;
  double acc[2] = {0.0, 0.0};
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n * R; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = j / R;
    const bool live = mask == nullptr || mask[i];
    const double di = diag[i];
    const double ri = live ? fma(-alpha, q[j], r[j]) : 0.0;
    const double zi = (live && di != 0.0) ? ri / di : 0.0;
    x[j] = fma(alpha, p[j], x[j]);
    r[j] = ri;
    z[j] = zi;
    acc[0] = fma(ri, zi, acc[0]);
    acc[1] = fma(ri, ri, acc[1]);
  }
  block_column_sums<R, 2>(acc, part);
}

template <int R>
__global__ void __launch_bounds__(RED_THREADS) k_pcg_direction_multi(int64_t n, const double* __restrict__ scal,
                                                                     const double* __restrict__ z, double* __restrict__ p) {
  const int k = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) % R);
  const double rz = scal[k * S_COUNT + S_RZ];
  const double beta = rz != 0.0 ? scal[k * S_COUNT + S_RZN] / rz : 0.0;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n * R; j += (int64_t)gridDim.x * blockDim.x)
    p[j] = fma(beta, p[j], z[j]);
}

template <int R>
__global__ void k_copy_scalar_multi(double* scal, int dst, int src) {
  if (threadIdx.x < R) scal[threadIdx.x * S_COUNT + dst] = scal[threadIdx.x * S_COUNT + src];
}

inline int64_t pcg_multi_work_doubles(int64_t n, int n_rhs) {
  return 4 * n * n_rhs + n + 2 * (int64_t)n_rhs * RED_BLOCKS + (int64_t)n_rhs * S_COUNT;
}

// R systems T^T A T y_k = b_k at once on the tiled pattern.  b, x: (n_rows + n_glob, R) row-major; b already folded;
// x returns the EXPANDED solutions (T y_k).  relres_h: R entries.
template <class Idx, int R>
int pcg_jacobi_multi(int64_t n_rows, const Idx* indptr, const Idx* indices, const double* data, const double* b, double* x,
                     const unsigned char* mask, double rtol, int max_iter, int check_every, double* work, int* iters_h,
                     double* relres_h, cudaStream_t stream, const BlockPattern& blk, const MpcMap* mpc) {
  const int64_t n = n_rows + (mpc != nullptr ? mpc->n_glob : 0);
  double* r = work;
  double* z = r + n * R;
  double* p = z + n * R;
  double* q = p + n * R;
  double* diag = q + n * R;
  double* part = diag + n;
  double* scal = part + 2 * (int64_t)R * RED_BLOCKS;
  k_csr_diagonal<Idx><<<(unsigned)((n_rows + 255) / 256), 256, 0, stream>>>(n_rows, indptr, indices, data, diag);
  if (mpc != nullptr) {
    if (int rc = mpc_fold(*mpc, diag, nullptr, true, stream)) return rc;
  }
  FDK_CUDA(cudaMemsetAsync(q, 0, sizeof(double) * n * R, stream));  // the global rows of q are only written by the fold
  k_pcg_init_multi<R><<<MULTI_BLOCKS, RED_THREADS, 0, stream>>>(n, b, diag, mask, x, r, z, p, part);
  k_reduce_final_multi<R><<<2 * R, RED_THREADS, 0, stream>>>(part, S_RZN, scal);  // S_RZN = rz, S_RR = rr
  k_copy_scalar_multi<R><<<1, 32, 0, stream>>>(scal, S_RZ, S_RZN);
  k_copy_scalar_multi<R><<<1, 32, 0, stream>>>(scal, S_BB, S_RR);
  FDK_CUDA(cudaGetLastError());
  double h[R * S_COUNT];
  FDK_CUDA(cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, stream));
  FDK_CUDA(cudaStreamSynchronize(stream));
  double bb[R], rr[R];
  bool any = false;
  for (int k = 0; k < R; ++k) {
    bb[k] = rr[k] = h[k * S_COUNT + S_BB];
    any = any || bb[k] > 0.0;
  }
  int it = 0;
  if (any) {
    while (it < max_iter) {
      if (mpc != nullptr) {
        if (int rc = mpc_expand_multi<R>(*mpc, p, stream)) return rc;
      }
      if (int rc = launch_bspmm<R>(blk, data, p, q, stream)) return rc;
      if (mpc != nullptr) {
        if (int rc = mpc_fold_multi<R>(*mpc, q, p, stream)) return rc;
      }
      k_dot_multi<R><<<MULTI_BLOCKS, RED_THREADS, 0, stream>>>(n, p, q, part);
      k_reduce_final_multi<R><<<R, RED_THREADS, 0, stream>>>(part, S_PQ, scal);
      k_pcg_update_multi<R><<<MULTI_BLOCKS, RED_THREADS, 0, stream>>>(n, scal, diag, mask, p, q, x, r, z, part);
      k_reduce_final_multi<R><<<2 * R, RED_THREADS, 0, stream>>>(part, S_RZN, scal);
      k_pcg_direction_multi<R><<<MULTI_BLOCKS, RED_THREADS, 0, stream>>>(n, scal, z, p);
      k_copy_scalar_multi<R><<<1, 32, 0, stream>>>(scal, S_RZ, S_RZN);
      ++it;
      if (it % check_every == 0 || it == max_iter) {
        FDK_CUDA(cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, stream));
        FDK_CUDA(cudaStreamSynchronize(stream));
        bool done = true;
        for (int k = 0; k < R; ++k) {
          rr[k] = h[k * S_COUNT + S_RR];
          done = done && !(rr[k] > rtol * rtol * bb[k]);  // a NaN also counts as finished
        }
        if (done) break;
      }
    }
    FDK_CUDA(cudaGetLastError());
  }
  if (mpc != nullptr) {
    if (int rc = mpc_expand_multi<R>(*mpc, x, stream)) return rc;
  }
  if (iters_h) *iters_h = it;
  if (relres_h)
    for (int k = 0; k < R; ++k) relres_h[k] = bb[k] > 0.0 ? sqrt(rr[k] / bb[k]) : 0.0;
  return 0;
}


}  // namespace fdk
