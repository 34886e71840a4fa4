// fdk_rows.cuh -- rows of HIGH-VALENCE nodes, one CTA per node (no capacity limit).
//
// The cluster kernels keep the geometry of every touched element of a cluster in shared memory, which bounds the
// number of elements around ONE node (tet10: 36, hex8: 45/80).  Unstructured tetrahedral meshes exceed that at a few
// vertices -- the reference's own util/meshes/octet_truss_quad.msh (BASELINE config 5) does.  The plan leaves such
// nodes out of the clusters (fedoo_b200/plan.py: ``heavy_nodes``) and this kernel assembles their rows: the CTA walks
// the incidences (element, local node) of its node ONE ELEMENT AT A TIME -- geometry of the element by NGP threads,
// then one thread per scalar entry of the element's block row -- and accumulates the row in shared memory at the
// positions found by binary search in the row's column list.  Fixed order (element-ascending), no atomics.  A slow path
// by construction: a handful of rows per mesh.
//
// Same arithmetic as the block phase of k_assemble (fdk_assemble.cuh): isotropic closed form K_IJ = lam S + mu S^T +
// mu tr(S) 1 with S = sum_g w grad N_I (x) grad N_J, or B_I^T C_g B_J for a uniform / per-Gauss-point tangent;
// residual D_I = -K_row . U (fused) or -sum_g w B_I^T sigma_g.  Reference: fedoo/core/assembly.py:143-470,
// fedoo/core/_sparsematrix.py:55-174,286-315.
#pragma once
#include "fdk_assemble.cuh"

namespace fdk {

struct RowArgs {
  int n_rows;
  const int32_t* rows;  // the nodes whose rows this launch assembles
  int n_nodes;
  int64_t n_elems;
  const int32_t* conn;
  const double* coords;
  const int64_t* node_ptr;  // node -> incidences (element * nne + local node), element-ascending
  const int32_t* node_inc;
  const int64_t* blk_indptr;
  const int32_t* blk_indices;
  int64_t blk_nnz;
  int max_deg;
  double lam, mu;
  double C[36];
  const double* tangent_gp;
  const double* U;
  const double* stress_gp;
  int compute;
  int fuse_ku;
  double* K;
  double* D;
};

// column a of B_k (Voigt rows xx, yy, zz, xy, xz, yz; engineering shears) for the gradient g of node k
template <int DIM>
__device__ __forceinline__ void b_column(const double* g, int a, double (&b)[6]) {
#pragma unroll
  for (int s = 0; s < 6; ++s) b[s] = 0.0;
  if constexpr (DIM == 3) {
    if (a == 0) { b[0] = g[0]; b[3] = g[1]; b[4] = g[2]; }
    else if (a == 1) { b[1] = g[1]; b[3] = g[0]; b[5] = g[2]; }
    else { b[2] = g[2]; b[4] = g[0]; b[5] = g[1]; }
  } else {
    if (a == 0) { b[0] = g[0]; b[3] = g[1]; }
    else { b[1] = g[1]; b[3] = g[0]; }
  }
}

template <class El, int PHYS>
__global__ void __launch_bounds__(128) k_assemble_rows(const __grid_constant__ RowArgs a) {
  static_assert(PHYS == PHYS_ISO || PHYS == PHYS_GENERAL, "rows kernel: elasticity");
  constexpr int NNE = El::NNE, NGP = El::NGP, DIM = El::DIM, NV = DIM, BLK = NV * NV;
  constexpr int THREADS = 128;
  static_assert(NNE * BLK <= THREADS && NGP + NNE <= THREADS, "one thread per entry of an element's block row");
  extern __shared__ double s_row[];  // [deg][BLK]
  __shared__ double sG[NGP][NNE][DIM];
  __shared__ double sWd[NGP];
  __shared__ int sCol[NNE];
  __shared__ double sF[NV];
  const int tid = threadIdx.x;
  const int I = a.rows[blockIdx.x];
  const ElemTable& tab = c_tab[El::ID];
  const int64_t bp = a.blk_indptr[I];
  const int deg = (int)(a.blk_indptr[I + 1] - bp);
  const int32_t* cols = a.blk_indices + bp;
  const bool do_mat = (a.compute & FDK_MATRIX) || a.fuse_ku;
  const bool do_bts = (a.compute & FDK_VECTOR) && !a.fuse_ku;
  for (int t = tid; t < deg * BLK; t += THREADS) s_row[t] = 0.0;
  if (tid < NV) sF[tid] = 0.0;
  __syncthreads();
  for (int64_t t = a.node_ptr[I]; t < a.node_ptr[I + 1]; ++t) {
    const int inc = a.node_inc[t];
    const int64_t e = inc / NNE;
    const int i = inc - (int)e * NNE;
    if (tid < NGP) {  // geometry of Gauss point tid
      double X[NNE][DIM];
#pragma unroll
      for (int k = 0; k < NNE; ++k) {
        const int nd = a.conn[e * NNE + k];
#pragma unroll
        for (int d = 0; d < DIM; ++d) X[k][d] = a.coords[(int64_t)nd * DIM + d];
      }
      double G[NNE][DIM];
      const double w = gp_geometry<NNE, DIM>(tab.dN + tid * DIM * NNE, tab.w[tid], X, G);
#pragma unroll
      for (int k = 0; k < NNE; ++k)
#pragma unroll
        for (int d = 0; d < DIM; ++d) sG[tid][k][d] = G[k][d];
      sWd[tid] = w;
    } else if (tid < NGP + NNE) {  // where the element's nodes sit in the row (sorted column list)
      const int k = tid - NGP;
      const int J = a.conn[e * NNE + k];
      int lo = 0, hi = deg - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cols[mid] < J) lo = mid + 1;
        else hi = mid;
      }
      sCol[k] = lo;
    }
    __syncthreads();
    if (do_mat && tid < NNE * BLK) {
      const int j = tid / BLK, b = tid - j * BLK, cc = b / NV, aa = b - cc * NV;
      double v = 0.0;
      if constexpr (PHYS == PHYS_ISO) {
        double s_ca = 0.0, s_ac = 0.0, tr = 0.0;
        for (int g = 0; g < NGP; ++g) {
          const double w = sWd[g];
          const double* gi = sG[g][i];
          const double* gj = sG[g][j];
          s_ca = fma(w * gi[cc], gj[aa], s_ca);
          s_ac = fma(w * gi[aa], gj[cc], s_ac);
          double dot = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) dot = fma(gi[d], gj[d], dot);
          tr = fma(w, dot, tr);
        }
        v = fma(a.lam, s_ca, a.mu * s_ac);
        if (cc == aa) v = fma(a.mu, tr, v);
      } else {
        for (int g = 0; g < NGP; ++g) {
          const double* Cg;
          int si, sj;
          if (a.tangent_gp != nullptr) {
            Cg = a.tangent_gp + 36 * ((int64_t)g * a.n_elems + e);
            si = 1;
            sj = 6;
          } else {
            Cg = a.C;
            si = 6;
            sj = 1;
          }
          double bi[6], bj[6];
          b_column<DIM>(sG[g][i], cc, bi);
          b_column<DIM>(sG[g][j], aa, bj);
          double q = 0.0;
#pragma unroll
          for (int s1 = 0; s1 < 6; ++s1) {
            double r = 0.0;
#pragma unroll
            for (int s2 = 0; s2 < 6; ++s2) r = fma(Cg[s1 * si + s2 * sj], bj[s2], r);
            q = fma(bi[s1], r, q);
          }
          v = fma(sWd[g], q, v);
        }
      }
      s_row[sCol[j] * BLK + b] += v;
    }
    if (do_bts && tid < NV) {
      double f = 0.0;
      for (int g = 0; g < NGP; ++g) {
        const double* sg = a.stress_gp + 6 * ((int64_t)g * a.n_elems + e);
        double bi[6];
        b_column<DIM>(sG[g][i], tid, bi);
        double q = 0.0;
#pragma unroll
        for (int s = 0; s < 6; ++s) q = fma(bi[s], sg[s], q);
        f = fma(sWd[g], q, f);
      }
      sF[tid] += f;
    }
    __syncthreads();
  }
  if (a.compute & FDK_MATRIX) {
    for (int t = tid; t < deg * BLK; t += THREADS) {
      const int s = t / BLK, b = t - s * BLK, cc = b / NV, aa = b - cc * NV;
      a.K[(int64_t)cc * NV * a.blk_nnz + (int64_t)NV * bp + (int64_t)aa * deg + s] = s_row[t];
    }
  }
  if ((a.compute & FDK_VECTOR) && tid < NV) {
    double d;
    if (a.fuse_ku) {
      d = 0.0;
      for (int s = 0; s < deg; ++s) {
        const int J = cols[s];
#pragma unroll
        for (int aa = 0; aa < NV; ++aa) d = fma(s_row[s * BLK + tid * NV + aa], a.U[(int64_t)aa * a.n_nodes + J], d);
      }
    } else {
      d = sF[tid];
    }
    a.D[(int64_t)tid * a.n_nodes + I] = -d;
  }
}

template <class El, int PHYS>
int launch_assemble_rows(const RowArgs& a, cudaStream_t stream) {
  if (a.n_rows == 0) return 0;
  if (int rc = ensure_device_tables()) return rc;
  const size_t smem = (size_t)a.max_deg * El::DIM * El::DIM * sizeof(double);
  FDK_REQUIRE(smem <= 40 * 1024, FDK_ECAP, "row degree %d exceeds the rows kernel's shared memory", a.max_deg);
  k_assemble_rows<El, PHYS><<<a.n_rows, 128, smem, stream>>>(a);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fdk
