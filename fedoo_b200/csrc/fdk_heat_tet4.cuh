// fdk_heat_tet4.cuh -- heat equation on tet4 meshes: row-owner kernel (BASELINE config [2]).
//
// tet4 has a CONSTANT gradient: the element matrix is V grad N_I . (k grad N_J) + delta_IJ (rho c / dt) V / 4 -- about
// 100 flops -- while the cluster kernel of fdk_assemble.cuh pays its descriptors, staging and barriers per incidence as
// if the element were expensive (measured: 13.7 ms for 20 M elements, 1 % of the HBM roofline; the compulsory traffic
// is 51 B per element).  Here ONE THREAD OWNS ONE NODE ROW: it walks the incidences (element, local node) of its node,
// recomputes the element's gradients from the four vertices (each element is visited by its four nodes: rho = 4 of a
// 100-flop computation, cheaper than any exchange), and adds the four entries of its row into the row's accumulator,
// which lives in shared memory as acc[slot][thread] (slot-major: the threads of a warp hit 32 consecutive banks).
// The position of every entry inside the row comes from a one-time table (4 x u8 per incidence, fedoo_b200/assembly.py).
// No atomics, a fixed summation order per row (element-ascending): bit-reproducible.  The residual
// D_I = -sum_g w [grad N_I . (k grad T) + (rho c / dt) N_I (T_g - T_start,g)] is accumulated on the way (the four nodal
// temperatures are the only extra loads), so K + D is one launch.
//
// Reference: fedoo/weakform/heat_equation.py:78-119 (conduction), :168-187 (capacity, lumped: assembly option
// mat_lumping, core/_sparsematrix.py:91-98); tet4 tables fedoo/lib_elements/tetrahedron.py:21-29,72-73,106-130
// (4 Gauss points of weight 1/24: for a constant integrand their sum is |det J| / 6 to rounding).
#pragma once
#include "fdk_assemble.cuh"

namespace fdk {

struct HeatTet4Args {
  int n_nodes;
  int64_t n_elems;
  const int32_t* conn;        // (n_elems, 4)
  const double* coords;       // (n_nodes, 3)
  const int64_t* node_ptr;    // [n_nodes + 1] incidences of every node ...
  const int2* inc_rec;        // ... {element * 4 + local node, 4 x u8 = position of column conn[e][j] in the node's row},
                              // element-ascending
  const int64_t* blk_indptr;  // [n_nodes + 1] block-CSR row pointers (nvar = 1: the CSR itself)
  const double* T;            // nodal temperature (NULL: no vector)
  const double* T_start;      // NULL = 0
  double cond[9];
  double rcdt;
  int max_deg;
  int n_rows;                 // rows this launch computes: all n_nodes, or ...
  const int32_t* rows;        // ... the listed ones (a rank's owned nodes; NULL = 0 .. n_nodes - 1)
  double* K;  // NULL: vector only
  double* D;
};

// Gradients of the four shape functions of a tet4 and |det J|, specialised to the reference element of
// fedoo/lib_elements/tetrahedron.py:106-130 (N = [eta, zeta, 1 - xi - eta - zeta, xi]: node 2 is the origin, xi runs to
// node 3, eta to node 0, zeta to node 1).  The table's dN/dxi entries are 0 and +-1, so J = dN . X is three edge
// vectors and G = J^-1 dN is the three columns of J^-1 and minus their sum: the same numbers, bit for bit, as the generic
// gp_geometry<4, 3> (same cofactor formulas, same order of the additions) for a third of the FP64 work.
__device__ __forceinline__ double tet4_gradients(const double (&X)[4][3], double (&G)[4][3]) {
  double J[3][3];
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    J[0][x] = X[3][x] - X[2][x];
    J[1][x] = X[0][x] - X[2][x];
    J[2][x] = X[1][x] - X[2][x];
  }
  double iJ[3][3];
  const double det = invert<3>(J, iJ);
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    G[3][x] = iJ[x][0];
    G[0][x] = iJ[x][1];
    G[1][x] = iJ[x][2];
    G[2][x] = ((-iJ[x][0]) - iJ[x][1]) - iJ[x][2];
  }
  return fabs(det);
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_heat_tet4_rows(const __grid_constant__ HeatTet4Args a) {
  extern __shared__ double s_acc[];  // [max_deg][THREADS]
  const int tid = threadIdx.x;
  const int r = blockIdx.x * THREADS + tid;
  if (r >= a.n_rows) return;
  const int I = a.rows != nullptr ? a.rows[r] : r;
  const ElemTable& tab = c_tab[FDK_TET4];
  const bool want_K = a.K != nullptr, want_D = a.D != nullptr && a.T != nullptr;
  const int64_t bp = a.blk_indptr[I];
  const int deg = (int)(a.blk_indptr[I + 1] - bp);
  if (want_K) {
    for (int s = 0; s < deg; ++s) s_acc[s * THREADS + tid] = 0.0;
  }
  // quadrature sums of the (symmetric) 4-point rule, from the table: sum_g w, sum_g w N_i, sum_g w N_i N_k (i = k, i != k)
  double wsum = 0.0, wn = 0.0, m_diag = 0.0, m_off = 0.0;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const double w = tab.w[g], n0 = tab.N[g * 4 + 0], n1 = tab.N[g * 4 + 1];
    wsum += w;
    wn = fma(w, n0, wn);
    m_diag = fma(w * n0, n0, m_diag);
    m_off = fma(w * n0, n1, m_off);
  }
  double dsum = 0.0;
  const int64_t t0 = a.node_ptr[I], t1 = a.node_ptr[I + 1];
  // software pipeline over the incidences: the record and the connectivity of incidence t + 1 are in flight while
  // incidence t computes (the chain record -> connectivity -> coordinates is three dependent global loads, and at 24
  // warps per SM there is little else to hide it behind)
  int2 rec_n = make_int2(0, 0);
  int4 nd_n = make_int4(0, 0, 0, 0);
  if (t0 < t1) {
    rec_n = a.inc_rec[t0];
    nd_n = *reinterpret_cast<const int4*>(a.conn + (int64_t)(rec_n.x >> 2) * 4);
  }
  for (int64_t t = t0; t < t1; ++t) {
    const int2 rec = rec_n;
    const int inc = rec.x;
    const int i = inc & 3;
    const int4 nd4 = nd_n;
    const int nd[4] = {nd4.x, nd4.y, nd4.z, nd4.w};
    if (t + 1 < t1) {  // (a second record in flight was measured too: no further gain)
      rec_n = a.inc_rec[t + 1];
      nd_n = *reinterpret_cast<const int4*>(a.conn + (int64_t)(rec_n.x >> 2) * 4);
    }
    double X[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int d = 0; d < 3; ++d) X[k][d] = a.coords[(int64_t)nd[k] * 3 + d];
    // the nodal temperatures are requested together with the coordinates, before the geometry is computed: they sit
    // behind a run-time flag, so the compiler leaves them where they are written (ncu had them as the top stall)
    double Tk[4] = {0.0, 0.0, 0.0, 0.0}, Ts[4] = {0.0, 0.0, 0.0, 0.0};
    if (want_D) {
#pragma unroll
      for (int k = 0; k < 4; ++k) Tk[k] = a.T[nd[k]];
      if (a.rcdt != 0.0 && a.T_start != nullptr) {
#pragma unroll
        for (int k = 0; k < 4; ++k) Ts[k] = a.T_start[nd[k]];
      }
    }
    double G[4][3];
    const double wdet = tet4_gradients(X, G);  // |det J|; the gradient is the same at every Gauss point
    // gradient of the row node (selects, not a run-time index: G stays in registers)
    double gi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) gi[d] = i == 0 ? G[0][d] : (i == 1 ? G[1][d] : (i == 2 ? G[2][d] : G[3][d]));
    const double V = wsum * wdet;
    if (want_K) {
      // k grad N_I (nothing here assumes a symmetric conductivity): row vector G_I . cond
      double kg[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) kg[c] = gi[0] * a.cond[0 * 3 + c] + gi[1] * a.cond[1 * 3 + c] + gi[2] * a.cond[2 * 3 + c];
      const uint32_t pos = (uint32_t)rec.y;
      const double cap = a.rcdt * wdet * wn;  // lumped capacity: row sum of the consistent matrix (sum_k N_k = 1)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double v = V * (kg[0] * G[j][0] + kg[1] * G[j][1] + kg[2] * G[j][2]);
        if (j == i) v += cap;
        s_acc[((pos >> (8 * j)) & 0xFF) * THREADS + tid] += v;
      }
    }
    if (want_D) {
      double gT[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int d = 0; d < 3; ++d) gT[d] = fma(Tk[k], G[k][d], gT[d]);
      }
      // grad N_I . (cond grad T)
      double q = 0.0;
#pragma unroll
      for (int r = 0; r < 3; ++r) q += gi[r] * (a.cond[r * 3 + 0] * gT[0] + a.cond[r * 3 + 1] * gT[1] + a.cond[r * 3 + 2] * gT[2]);
      double f = V * q;
      if (a.rcdt != 0.0) {
        // sum_g w N_i(g) sum_k N_k(g) dT_k = m_off sum_k dT_k + (m_diag - m_off) dT_i
        double dT[4], ssum = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          dT[k] = Tk[k] - Ts[k];
          ssum += dT[k];
        }
        const double dTi = i == 0 ? dT[0] : (i == 1 ? dT[1] : (i == 2 ? dT[2] : dT[3]));
        f = fma(a.rcdt * wdet, fma(m_diag - m_off, dTi, m_off * ssum), f);
      }
      dsum += f;
    }
  }
  if (want_K) {
    double* row = a.K + bp;
    for (int s = 0; s < deg; ++s) __stcs(row + s, s_acc[s * THREADS + tid]);
  }
  if (want_D) a.D[I] = -dsum;
}

template <int THREADS, int MINB>
int launch_heat_tet4_t(const HeatTet4Args& a, size_t smem, cudaStream_t stream) {
  static thread_local size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    FDK_CUDA(cudaFuncSetAttribute(k_heat_tet4_rows<THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  k_heat_tet4_rows<THREADS, MINB><<<(unsigned)((a.n_rows + THREADS - 1) / THREADS), THREADS, smem, stream>>>(a);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

inline int launch_heat_tet4(const HeatTet4Args& a, cudaStream_t stream) {
  if (a.n_nodes == 0 || a.n_rows == 0) return 0;
  if (int rc = ensure_device_tables()) return rc;
  constexpr int THREADS = 128;
  const size_t smem = (size_t)(a.K != nullptr ? a.max_deg : 0) * THREADS * sizeof(double);
  FDK_REQUIRE(a.max_deg <= 255 && smem <= 200 * 1024, FDK_ECAP, "row degree %d exceeds the row-owner kernel's capacity", a.max_deg);
  // resident CTAs per SM the register allocation aims at (128 / 96 / 80 / 64 registers); FDK_HEAT_MINB overrides
  // (diagnostic).  Measured at 20 M elements, K + D: 5 -> 1.38 ms (no spills), 6 -> 1.64 ms, 8 slower still
  static const int minb = [] {
    const char* e = getenv("FDK_HEAT_MINB");
    return e ? atoi(e) : 5;
  }();
  if (minb >= 8) return launch_heat_tet4_t<THREADS, 8>(a, smem, stream);
  if (minb >= 6) return launch_heat_tet4_t<THREADS, 6>(a, smem, stream);
  if (minb >= 5) return launch_heat_tet4_t<THREADS, 5>(a, smem, stream);
  return launch_heat_tet4_t<THREADS, 4>(a, smem, stream);
}

}  // namespace fdk
