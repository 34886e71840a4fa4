// fdk_gp.cuh -- element-parallel Gauss-point kernels: strain/stress update, thermal state,
// J2 radial return.  One thread per Gauss point, gp-major indexing n = g * n_elems + e so that a
// warp handles 32 consecutive elements at the same Gauss point (uniform table reads from
// constant memory, fully coalesced (6,N)/(8,N) column stores).
#pragma once
#include "fdk_assemble.cuh"

namespace fdk {

struct GpArgs {
  int n_nodes;
  int64_t n_elems;
  const int32_t* conn;
  const double* coords;
  const double* U;
  const double* tangent_gp;
  double* grad_gp;
  double* strain_gp;
  double* stress_gp;
  double* temp_gp;
  double* temp_grad_gp;
  double C[36];
  int has_C;
  double* fbar_center;  // (n_elems) mean over the element's Gauss points of tr(grad u): small-strain F-bar, or NULL
};

// grad u -> strain -> stress.  Replaces the 9 SpMVs of Assembly.get_grad_disp
// (fedoo/core/assembly.py:1285-1336), _comp_linear_strain
// (fedoo/weakform/stress_equilibrium.py:589-601) and the 36 array multiplies of
// ElasticAnisotropic.update (fedoo/constitutivelaw/elastic_anisotropic.py:48-56).
template <class El>
__global__ void __launch_bounds__(256) k_gp_strain_stress(const __grid_constant__ GpArgs a) {
  constexpr int NNE = El::NNE, NGP = El::NGP, DIM = El::DIM;
  const int64_t N = a.n_elems * NGP;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int g = (int)(n / a.n_elems);
  const int64_t e = n - (int64_t)g * a.n_elems;
  const ElemTable& tab = c_tab[El::ID];
  int nd[NNE];
  double X[NNE][DIM];
#pragma unroll
  for (int k = 0; k < NNE; ++k) {
    nd[k] = a.conn[e * NNE + k];
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[k][d] = a.coords[(int64_t)nd[k] * DIM + d];
  }
  double dN[DIM * NNE];
#pragma unroll
  for (int t = 0; t < DIM * NNE; ++t) dN[t] = tab.dN[g * DIM * NNE + t];
  double G[NNE][DIM];
  gp_geometry<NNE, DIM>(dN, 1.0, X, G);
  double gu[DIM][DIM];
#pragma unroll
  for (int v = 0; v < DIM; ++v)
#pragma unroll
    for (int d = 0; d < DIM; ++d) gu[v][d] = 0.0;
#pragma unroll
  for (int k = 0; k < NNE; ++k)
#pragma unroll
    for (int v = 0; v < DIM; ++v) {
      const double u = a.U[(int64_t)v * a.n_nodes + nd[k]];
#pragma unroll
      for (int d = 0; d < DIM; ++d) gu[v][d] = fma(u, G[k][d], gu[v][d]);
    }
  if (a.fbar_center != nullptr) {
    // small-strain F-bar (fedoo/weakform/stress_equilibrium.py:527-540): the volumetric part of grad u is replaced by
    // its mean over the element's Gauss points, grad_ii -= (tr - mean tr) / 3
    if constexpr (DIM == 3) {
      const double shift = ((gu[0][0] + gu[1][1] + gu[2][2]) - a.fbar_center[e]) / 3.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) gu[d][d] -= shift;
    }
  }
  if (a.grad_gp != nullptr) {
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        a.grad_gp[(int64_t)(v * 3 + d) * N + n] = (v < DIM && d < DIM) ? gu[v < DIM ? v : 0][d < DIM ? d : 0] : 0.0;
      }
  }
  double eps[6];
  voigt_strain<DIM>(gu, eps);
  if (a.strain_gp != nullptr) {
#pragma unroll
    for (int s = 0; s < 6; ++s) a.strain_gp[6 * n + s] = eps[s];
  }
  if (a.stress_gp != nullptr) {
    double sig[6];
    if (a.tangent_gp != nullptr)
      apply_tangent(a.tangent_gp + 36 * n, 1, 6, eps, sig);
    else
      apply_tangent(a.C, 6, 1, eps, sig);
#pragma unroll
    for (int s = 0; s < 6; ++s) a.stress_gp[6 * n + s] = sig[s];
  }
}

// fbar_center[e] = mean over the Gauss points of element e of tr(grad u): one thread per element, Gauss points in
// order (the reference's np.mean over the gp axis, stress_equilibrium.py:533)
template <class El>
__global__ void __launch_bounds__(128) k_gp_fbar_center(const __grid_constant__ GpArgs a) {
  constexpr int NNE = El::NNE, NGP = El::NGP, DIM = El::DIM;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n_elems) return;
  const ElemTable& tab = c_tab[El::ID];
  int nd[NNE];
  double X[NNE][DIM];
#pragma unroll
  for (int k = 0; k < NNE; ++k) {
    nd[k] = a.conn[e * NNE + k];
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[k][d] = a.coords[(int64_t)nd[k] * DIM + d];
  }
  double sum = 0.0;
  for (int g = 0; g < NGP; ++g) {
    double dN[DIM * NNE];
#pragma unroll
    for (int t = 0; t < DIM * NNE; ++t) dN[t] = tab.dN[g * DIM * NNE + t];
    double G[NNE][DIM];
    gp_geometry<NNE, DIM>(dN, 1.0, X, G);
    double gd[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) gd[d] = 0.0;
#pragma unroll
    for (int k = 0; k < NNE; ++k)
#pragma unroll
      for (int d = 0; d < DIM; ++d) gd[d] = fma(a.U[(int64_t)d * a.n_nodes + nd[k]], G[k][d], gd[d]);
    double tr = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) tr += gd[d];
    sum += tr;
  }
  a.fbar_center[e] = sum / NGP;
}

template <class El>
int launch_gp_fbar_center(const GpArgs& a, cudaStream_t stream) {
  if (a.n_elems == 0) return 0;
  if (int rc = ensure_device_tables()) return rc;
  k_gp_fbar_center<El><<<(unsigned)((a.n_elems + 127) / 128), 128, 0, stream>>>(a);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

// Deformation gradient at the Gauss points, F = 1 + grad u, and its F-bar form F (J_mean / J)^(1/3) with J = det F and
// J_mean the mean of J over the element's Gauss points -- the two finite-strain kinematic quantities the reference
// computes in its own Python before it hands over to simcoon (fedoo/weakform/stress_equilibrium.py:542-586, _comp_F /
// _comp_Fbar).  One thread per element, Gauss points in order (np.mean over the gp axis).  F_gp is (N, 9), entry
// i + 3 j of Gauss point n = F_ij: the memory layout of the reference's Fortran-ordered (3, 3, N) array.  2-D meshes:
// F_33 = 1, the out-of-plane shears 0.
struct DefGradArgs {
  int n_nodes;
  int64_t n_elems;
  const int32_t* conn;
  const double* coords;
  const double* U;
  int fbar;
  double* F_gp;
};

template <class El>
__global__ void __launch_bounds__(128) k_gp_defgrad(const __grid_constant__ DefGradArgs a) {
  constexpr int NNE = El::NNE, NGP = El::NGP, DIM = El::DIM;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n_elems) return;
  const ElemTable& tab = c_tab[El::ID];
  int nd[NNE];
  double X[NNE][DIM];
#pragma unroll
  for (int k = 0; k < NNE; ++k) {
    nd[k] = a.conn[e * NNE + k];
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[k][d] = a.coords[(int64_t)nd[k] * DIM + d];
  }
  double Jg[NGP];
  double Jsum = 0.0;
#pragma unroll 1
  for (int g = 0; g < NGP; ++g) {
    double dN[DIM * NNE];
#pragma unroll
    for (int t = 0; t < DIM * NNE; ++t) dN[t] = tab.dN[g * DIM * NNE + t];
    double G[NNE][DIM];
    gp_geometry<NNE, DIM>(dN, 1.0, X, G);
    double F[3][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}};
#pragma unroll
    for (int v = 0; v < DIM; ++v) {
      double gu[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) gu[d] = 0.0;
#pragma unroll
      for (int k = 0; k < NNE; ++k) {
        const double u = a.U[(int64_t)v * a.n_nodes + nd[k]];
#pragma unroll
        for (int d = 0; d < DIM; ++d) gu[d] = fma(u, G[k][d], gu[d]);
      }
#pragma unroll
      for (int d = 0; d < DIM; ++d) F[v][d] += gu[d];
    }
    const double J = F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
                     F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
    Jg[g] = J;
    Jsum += J;
    double* out = a.F_gp + 9 * ((int64_t)g * a.n_elems + e);
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) out[i + 3 * j] = F[i][j];
  }
  if (a.fbar) {
    const double Jc = Jsum / NGP;
#pragma unroll 1
    for (int g = 0; g < NGP; ++g) {
      const double sc = pow(Jc / Jg[g], 1.0 / 3.0);
      double* out = a.F_gp + 9 * ((int64_t)g * a.n_elems + e);
#pragma unroll
      for (int t = 0; t < 9; ++t) out[t] *= sc;
    }
  }
}

template <class El>
int launch_gp_defgrad(const DefGradArgs& a, cudaStream_t stream) {
  if (a.n_elems == 0) return 0;
  if (int rc = ensure_device_tables()) return rc;
  k_gp_defgrad<El><<<(unsigned)((a.n_elems + 127) / 128), 128, 0, stream>>>(a);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Residual alone (compute = "vector": every Newton sub-iteration of fedoo/problem/non_linear.py:400-404 asks for it):
// D = -int B^T sigma without the cluster machinery of the matrix kernels.  Pass 1, one thread per element: geometry at
// every Gauss point, sigma from stress_gp or recomputed from U (grad u -> eps -> C eps), nodal forces of the element
// into fe[e][k][d].  Pass 2, one thread per node: sum of its incidences in a fixed order (no atomics, bit-reproducible).
// Replaces the vector branch of Assembly.assemble_global_mat (fedoo/core/assembly.py:400-411).
struct ResArgs {
  int n_nodes;
  int64_t n_elems;
  const int32_t* conn;
  const double* coords;
  const double* U;           // used when stress_gp == NULL
  const double* stress_gp;   // (6, N) column-major, gp-major N, or NULL
  const double* tangent_gp;  // (6, 6, N) Fortran order, or NULL -> C
  double C[36];
  double* fe;  // (n_elems, NNE, DIM)
};

template <class El>
__global__ void __launch_bounds__(128) k_elem_force(const __grid_constant__ ResArgs a) {
  constexpr int NNE = El::NNE, NGP = El::NGP, DIM = El::DIM;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n_elems) return;
  const ElemTable& tab = c_tab[El::ID];
  int nd[NNE];
  double X[NNE][DIM];
#pragma unroll
  for (int k = 0; k < NNE; ++k) {
    nd[k] = a.conn[e * NNE + k];
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[k][d] = a.coords[(int64_t)nd[k] * DIM + d];
  }
  double f[NNE][DIM];
#pragma unroll
  for (int k = 0; k < NNE; ++k)
#pragma unroll
    for (int d = 0; d < DIM; ++d) f[k][d] = 0.0;
  const int64_t N = a.n_elems * NGP;
#pragma unroll 1
  for (int g = 0; g < NGP; ++g) {
    double dN[DIM * NNE];
#pragma unroll
    for (int t = 0; t < DIM * NNE; ++t) dN[t] = tab.dN[g * DIM * NNE + t];
    double G[NNE][DIM];
    const double w = gp_geometry<NNE, DIM>(dN, tab.w[g], X, G);
    const int64_t n = (int64_t)g * a.n_elems + e;
    double sig[6];
    if (a.stress_gp != nullptr) {
#pragma unroll
      for (int s = 0; s < 6; ++s) sig[s] = a.stress_gp[6 * n + s];
    } else {
      double gu[DIM][DIM];
#pragma unroll
      for (int v = 0; v < DIM; ++v)
#pragma unroll
        for (int d = 0; d < DIM; ++d) gu[v][d] = 0.0;
#pragma unroll
      for (int k = 0; k < NNE; ++k)
#pragma unroll
        for (int v = 0; v < DIM; ++v) {
          const double u = a.U[(int64_t)v * a.n_nodes + nd[k]];
#pragma unroll
          for (int d = 0; d < DIM; ++d) gu[v][d] = fma(u, G[k][d], gu[v][d]);
        }
      double eps[6];
      voigt_strain<DIM>(gu, eps);
      if (a.tangent_gp != nullptr) apply_tangent(a.tangent_gp + 36 * n, 1, 6, eps, sig);
      else apply_tangent(a.C, 6, 1, eps, sig);
    }
    (void)N;
    // sigma as a tensor (Voigt xx, yy, zz, xy, xz, yz); f_k[d] += w sum_j sigma_dj dN_k/dx_j
    const double S[3][3] = {{sig[0], sig[3], sig[4]}, {sig[3], sig[1], sig[5]}, {sig[4], sig[5], sig[2]}};
#pragma unroll
    for (int k = 0; k < NNE; ++k)
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < DIM; ++j) t = fma(S[d][j], G[k][j], t);
        f[k][d] = fma(w, t, f[k][d]);
      }
  }
  double* out = a.fe + e * (NNE * DIM);
#pragma unroll
  for (int k = 0; k < NNE; ++k)
#pragma unroll
    for (int d = 0; d < DIM; ++d) out[k * DIM + d] = f[k][d];
}

// hex8 specialisation of k_elem_force in the MONOMIAL basis of the trilinear element.  With the nodal signs s_k (the
// reference coordinates of node k, fedoo/lib_elements/hexahedron.py) N_k = 1/8 (1 + s_kx xi)(1 + s_ky eta)(1 + s_kz zeta),
// so a nodal field is c_0 + c_1 xi + c_2 eta + c_3 zeta + c_4 xi eta + c_5 eta zeta + c_6 xi zeta + c_7 xi eta zeta with
// c_m = 1/8 sum_k h_m(k) x_k (h_m = the products of signs).  Seven coefficient vectors of the coordinates and seven of the
// displacement replace X[8][3], U[8][3]; J and the reference gradient of u at a Gauss point are 27 FMA each instead of 72,
// the physical gradients of the eight shape functions are never formed, and the nodal forces are accumulated as seven
// coefficient vectors too (f_k = 1/8 sum_m h_m(k) b_m, b_m = sum_g sum_r d mono_m / d xi_r T_g[r], T = w J^-1 sigma): about
// 200 FMA per Gauss point instead of 345, in 126 instead of 144 + 48 live registers.  Same integrand, same quadrature
// (2 x 2 x 2 points, xi slowest, w = 1); the sums are taken in another order, so results agree to rounding, not bitwise.
struct HexMono {
  // sign of node k along axis d (same table as HexRef::bit in fdk_assemble_iso.cuh)
  __host__ __device__ static constexpr double s(int k, int d) {
    return (d == 0 ? ((k ^ (k >> 1)) & 1) : (d == 1 ? ((k >> 1) & 1) : ((k >> 2) & 1))) ? 1.0 : -1.0;
  }
  // h_m(k), m = 1..7: xi, eta, zeta, xi eta, eta zeta, xi zeta, xi eta zeta
  __host__ __device__ static constexpr double h(int m, int k) {
    return m == 1 ? s(k, 0) : m == 2 ? s(k, 1) : m == 3 ? s(k, 2) : m == 4 ? s(k, 0) * s(k, 1) : m == 5 ? s(k, 1) * s(k, 2)
         : m == 6 ? s(k, 0) * s(k, 2) : s(k, 0) * s(k, 1) * s(k, 2);
  }
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_elem_force_hex8(const __grid_constant__ ResArgs a) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n_elems) return;
  int nd[8];
  {
    const int4* c4 = reinterpret_cast<const int4*>(a.conn + e * 8);
    const int4 lo = c4[0], hi = c4[1];
    nd[0] = lo.x; nd[1] = lo.y; nd[2] = lo.z; nd[3] = lo.w; nd[4] = hi.x; nd[5] = hi.y; nd[6] = hi.z; nd[7] = hi.w;
  }
  const bool from_u = a.stress_gp == nullptr;
  double cX[7][3], cU[7][3];
#pragma unroll
  for (int m = 0; m < 7; ++m)
#pragma unroll
    for (int d = 0; d < 3; ++d) cX[m][d] = cU[m][d] = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double x[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = a.coords[(int64_t)nd[k] * 3 + d];
#pragma unroll
    for (int m = 0; m < 7; ++m)
#pragma unroll
      for (int d = 0; d < 3; ++d) cX[m][d] += HexMono::h(m + 1, k) * x[d];  // +- x: the sign is a compile-time constant
    if (from_u) {
      double u[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) u[d] = a.U[(int64_t)d * a.n_nodes + nd[k]];
#pragma unroll
      for (int m = 0; m < 7; ++m)
#pragma unroll
        for (int d = 0; d < 3; ++d) cU[m][d] += HexMono::h(m + 1, k) * u[d];
    }
  }
#pragma unroll
  for (int m = 0; m < 7; ++m)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      cX[m][d] *= 0.125;
      cU[m][d] *= 0.125;
    }
  double b[7][3];
#pragma unroll
  for (int m = 0; m < 7; ++m)
#pragma unroll
    for (int d = 0; d < 3; ++d) b[m][d] = 0.0;
  constexpr double GA = 0.5773502691896258;
#pragma unroll 1
  for (int g = 0; g < 8; ++g) {
    const double xi = (g & 4) ? GA : -GA, et = (g & 2) ? GA : -GA, ze = (g & 1) ? GA : -GA;
    const double xe = xi * et, ez = et * ze, xz = xi * ze;
    // d/dxi_r of a field with coefficients c: r = 0: c1 + c4 eta + c6 zeta + c7 eta zeta, ...
    double J[3][3];
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      J[0][x] = fma(cX[6][x], ez, fma(cX[5][x], ze, fma(cX[3][x], et, cX[0][x])));
      J[1][x] = fma(cX[6][x], xz, fma(cX[4][x], ze, fma(cX[3][x], xi, cX[1][x])));
      J[2][x] = fma(cX[6][x], xe, fma(cX[5][x], xi, fma(cX[4][x], et, cX[2][x])));
    }
    double iJ[3][3];
    const double w = fabs(invert<3>(J, iJ));  // w_g = 1
    const int64_t n = (int64_t)g * a.n_elems + e;
    double sig[6];
    if (!from_u) {
#pragma unroll
      for (int q = 0; q < 6; ++q) sig[q] = a.stress_gp[6 * n + q];
    } else {
      double Gu[3][3];  // d u_v / d xi_r
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        Gu[0][v] = fma(cU[6][v], ez, fma(cU[5][v], ze, fma(cU[3][v], et, cU[0][v])));
        Gu[1][v] = fma(cU[6][v], xz, fma(cU[4][v], ze, fma(cU[3][v], xi, cU[1][v])));
        Gu[2][v] = fma(cU[6][v], xe, fma(cU[5][v], xi, fma(cU[4][v], et, cU[2][v])));
      }
      double gu[3][3];  // d u_v / d x_d = sum_r iJ[d][r] Gu[r][v]
#pragma unroll
      for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int d = 0; d < 3; ++d) gu[v][d] = fma(iJ[d][2], Gu[2][v], fma(iJ[d][1], Gu[1][v], iJ[d][0] * Gu[0][v]));
      double eps[6];
      voigt_strain<3>(gu, eps);
      if (a.tangent_gp != nullptr) apply_tangent(a.tangent_gp + 36 * n, 1, 6, eps, sig);
      else apply_tangent(a.C, 6, 1, eps, sig);
    }
    const double S[3][3] = {{sig[0], sig[3], sig[4]}, {sig[3], sig[1], sig[5]}, {sig[4], sig[5], sig[2]}};
    double T[3][3];  // T[r][v] = w sum_d iJ[d][r] S[v][d]
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int v = 0; v < 3; ++v) T[r][v] = w * fma(iJ[2][r], S[v][2], fma(iJ[1][r], S[v][1], iJ[0][r] * S[v][0]));
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      b[0][v] += T[0][v];
      b[1][v] += T[1][v];
      b[2][v] += T[2][v];
      b[3][v] = fma(et, T[0][v], fma(xi, T[1][v], b[3][v]));                      // xi eta
      b[4][v] = fma(ze, T[1][v], fma(et, T[2][v], b[4][v]));                      // eta zeta
      b[5][v] = fma(ze, T[0][v], fma(xi, T[2][v], b[5][v]));                      // xi zeta
      b[6][v] = fma(ez, T[0][v], fma(xz, T[1][v], fma(xe, T[2][v], b[6][v])));    // xi eta zeta
    }
  }
  double* out = a.fe + e * 24;
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      double t = 0.0;
#pragma unroll
      for (int m = 0; m < 7; ++m) t += HexMono::h(m + 1, k) * b[m][v];
      out[k * 3 + v] = 0.125 * t;
    }
}

// Heat counterpart (fedoo/weakform/heat_equation.py:78-119,168-187): f_k = sum_g w [grad N_k . (cond grad T) +
// (rho c / dt) N_k (T_g - T_start,g)], one dof per node.
struct ResHeatArgs {
  int n_nodes;
  int64_t n_elems;
  const int32_t* conn;
  const double* coords;
  const double* T;
  const double* T_start;  // NULL: zero start temperature
  double cond[9];
  double rcdt;
  double* fe;  // (n_elems, NNE)
};

template <class El>
__global__ void __launch_bounds__(128) k_elem_force_heat(const __grid_constant__ ResHeatArgs a) {
  constexpr int NNE = El::NNE, NGP = El::NGP, DIM = El::DIM;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n_elems) return;
  const ElemTable& tab = c_tab[El::ID];
  double X[NNE][DIM], T[NNE], dT[NNE], f[NNE];
#pragma unroll
  for (int k = 0; k < NNE; ++k) {
    const int nd = a.conn[e * NNE + k];
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[k][d] = a.coords[(int64_t)nd * DIM + d];
    T[k] = a.T[nd];
    dT[k] = a.rcdt != 0.0 ? T[k] - (a.T_start != nullptr ? a.T_start[nd] : 0.0) : 0.0;  // NULL: T_start = 0 (heat_equation.py:140-147)
    f[k] = 0.0;
  }
#pragma unroll 1
  for (int g = 0; g < NGP; ++g) {
    const double* dN = tab.dN + g * DIM * NNE;
    double G[NNE][DIM];
    const double w = gp_geometry<NNE, DIM>(dN, tab.w[g], X, G);
    double gT[DIM], dTg = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) gT[d] = 0.0;
#pragma unroll
    for (int k = 0; k < NNE; ++k) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) gT[d] = fma(T[k], G[k][d], gT[d]);
      dTg = fma(tab.N[g * NNE + k], dT[k], dTg);
    }
    double q[DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      double t = 0.0;
#pragma unroll
      for (int j = 0; j < DIM; ++j) t = fma(a.cond[i * 3 + j], gT[j], t);
      q[i] = t;
    }
    const double cap = a.rcdt * dTg;
#pragma unroll
    for (int k = 0; k < NNE; ++k) {
      double t = tab.N[g * NNE + k] * cap;
#pragma unroll
      for (int d = 0; d < DIM; ++d) t = fma(G[k][d], q[d], t);
      f[k] = fma(w, t, f[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < NNE; ++k) a.fe[e * NNE + k] = f[k];
}

template <class El>
int launch_residual_heat(const ResHeatArgs& a, const int64_t* node_ptr, const int32_t* node_inc, double* D,
                         cudaStream_t stream);

// The same residual from GIVEN Gauss-point fields -- what the reference's weak forms hand to the assembly through
// ``assembly.sv`` (heat_equation.py:99-117: grad v . (K TempGradient); :178-186: (rho c / dt) v (Temp - Temp_start)):
// f_k = sum_g w [grad N_k . flux_g + N_k src_g], flux (3, N) row-major or NULL, src (N,) or NULL, gp-major columns.
struct ResHeatGpArgs {
  int n_nodes;
  int64_t n_elems;
  const int32_t* conn;
  const double* coords;
  const double* flux_gp;
  const double* src_gp;
  double* fe;  // (n_elems, NNE)
};

template <class El>
__global__ void __launch_bounds__(128) k_elem_force_heat_gp(const __grid_constant__ ResHeatGpArgs a) {
  constexpr int NNE = El::NNE, NGP = El::NGP, DIM = El::DIM;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n_elems) return;
  const ElemTable& tab = c_tab[El::ID];
  const int64_t N = a.n_elems * NGP;
  double X[NNE][DIM], f[NNE];
#pragma unroll
  for (int k = 0; k < NNE; ++k) {
    const int nd = a.conn[e * NNE + k];
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[k][d] = a.coords[(int64_t)nd * DIM + d];
    f[k] = 0.0;
  }
#pragma unroll 1
  for (int g = 0; g < NGP; ++g) {
    const double* dN = tab.dN + g * DIM * NNE;
    double G[NNE][DIM];
    const double w = gp_geometry<NNE, DIM>(dN, tab.w[g], X, G);
    const int64_t n = (int64_t)g * a.n_elems + e;
    double q[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) q[d] = a.flux_gp != nullptr ? a.flux_gp[(int64_t)d * N + n] : 0.0;
    const double src = a.src_gp != nullptr ? a.src_gp[n] : 0.0;
#pragma unroll
    for (int k = 0; k < NNE; ++k) {
      double t = tab.N[g * NNE + k] * src;
#pragma unroll
      for (int d = 0; d < DIM; ++d) t = fma(G[k][d], q[d], t);
      f[k] = fma(w, t, f[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < NNE; ++k) a.fe[e * NNE + k] = f[k];
}

// D[d n_nodes + I] = -sum over the incidences (element, local node) of node I, in the order of the list
template <int DIM>
__global__ void __launch_bounds__(256) k_node_force_gather(int n_nodes, const int64_t* __restrict__ node_ptr,
                                                            const int32_t* __restrict__ node_inc,
                                                            const double* __restrict__ fe, double* __restrict__ D) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= n_nodes) return;
  double acc[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) acc[d] = 0.0;
  for (int64_t t = node_ptr[I]; t < node_ptr[I + 1]; ++t) {
    const double* src = fe + (int64_t)node_inc[t] * DIM;
#pragma unroll
    for (int d = 0; d < DIM; ++d) acc[d] += src[d];
  }
#pragma unroll
  for (int d = 0; d < DIM; ++d) D[(int64_t)d * n_nodes + I] = -acc[d];
}

template <class El>
int launch_residual_heat(const ResHeatArgs& a, const int64_t* node_ptr, const int32_t* node_inc, double* D,
                         cudaStream_t stream) {
  if (a.n_elems == 0 || a.n_nodes == 0) return 0;
  if (int rc = ensure_device_tables()) return rc;
  k_elem_force_heat<El><<<(unsigned)((a.n_elems + 127) / 128), 128, 0, stream>>>(a);
  k_node_force_gather<1><<<(unsigned)((a.n_nodes + 255) / 256), 256, 0, stream>>>(a.n_nodes, node_ptr, node_inc, a.fe, D);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

template <class El>
int launch_residual_heat_gp(const ResHeatGpArgs& a, const int64_t* node_ptr, const int32_t* node_inc, double* D,
                            cudaStream_t stream) {
  if (a.n_elems == 0 || a.n_nodes == 0) return 0;
  if (int rc = ensure_device_tables()) return rc;
  k_elem_force_heat_gp<El><<<(unsigned)((a.n_elems + 127) / 128), 128, 0, stream>>>(a);
  k_node_force_gather<1><<<(unsigned)((a.n_nodes + 255) / 256), 256, 0, stream>>>(a.n_nodes, node_ptr, node_inc, a.fe, D);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

// Gauss-point-parallel form of k_elem_force (round 2, kept as an option: measured slower, see launch_residual): the
// thread-per-element kernel keeps X, G and the 24 nodal forces in
// registers across the Gauss-point loop (255 registers + spills, 8 warps per SM; ncu: FP64 45 %, LSU 48 % busy).  Here a
// CTA takes EPB consecutive elements and one thread one (element, Gauss point) -- a warp = 32 (EPB) consecutive elements
// at the same Gauss point, so the (6,N) stress loads are coalesced -- and the contributions w B_k^T sigma_g are summed over
// the Gauss points through shared memory ([g][k d][element], row stride EPB + 1: conflict-free both ways) in a FIXED
// order g = 0 .. NGP-1, the order of the element loop: the same numbers as k_elem_force, bit for bit.
template <class El>
struct ElemForceGp {
  static constexpr int NNE = El::NNE, NGP = El::NGP, DIM = El::DIM, ND = NNE * DIM;
  static constexpr int EPB = (NGP * 32 <= 256) ? 32 : 16;  // elements per CTA (tet10: 16 x 15 = 240 threads)
  static constexpr int THREADS = EPB * NGP;
  static constexpr int RSTR = EPB + 1;
  static constexpr size_t SMEM = (size_t)NGP * ND * RSTR * sizeof(double);
};

template <class El>
__global__ void __launch_bounds__(ElemForceGp<El>::THREADS) k_elem_force_gp(const __grid_constant__ ResArgs a) {
  using T = ElemForceGp<El>;
  constexpr int NNE = T::NNE, NGP = T::NGP, DIM = T::DIM, ND = T::ND, EPB = T::EPB, RSTR = T::RSTR;
  extern __shared__ double s_c[];  // [NGP][ND][RSTR]
  const int tid = threadIdx.x;
  const int g = tid / EPB, el = tid - g * EPB;
  const int64_t e0 = (int64_t)blockIdx.x * EPB;
  const int64_t e = e0 + el;
  if (e < a.n_elems) {
    const ElemTable& tab = c_tab[El::ID];
    int nd[NNE];
    double X[NNE][DIM];
#pragma unroll
    for (int k = 0; k < NNE; ++k) {
      nd[k] = a.conn[e * NNE + k];
#pragma unroll
      for (int d = 0; d < DIM; ++d) X[k][d] = a.coords[(int64_t)nd[k] * DIM + d];
    }
    double G[NNE][DIM];
    const double w = gp_geometry<NNE, DIM>(tab.dN + g * DIM * NNE, tab.w[g], X, G);
    const int64_t n = (int64_t)g * a.n_elems + e;
    double sig[6];
    if (a.stress_gp != nullptr) {
#pragma unroll
      for (int s = 0; s < 6; ++s) sig[s] = a.stress_gp[6 * n + s];
    } else {
      double gu[DIM][DIM];
#pragma unroll
      for (int v = 0; v < DIM; ++v)
#pragma unroll
        for (int d = 0; d < DIM; ++d) gu[v][d] = 0.0;
#pragma unroll
      for (int k = 0; k < NNE; ++k)
#pragma unroll
        for (int v = 0; v < DIM; ++v) {
          const double u = a.U[(int64_t)v * a.n_nodes + nd[k]];
#pragma unroll
          for (int d = 0; d < DIM; ++d) gu[v][d] = fma(u, G[k][d], gu[v][d]);
        }
      double eps[6];
      voigt_strain<DIM>(gu, eps);
      if (a.tangent_gp != nullptr) apply_tangent(a.tangent_gp + 36 * n, 1, 6, eps, sig);
      else apply_tangent(a.C, 6, 1, eps, sig);
    }
    const double S[3][3] = {{sig[0], sig[3], sig[4]}, {sig[3], sig[1], sig[5]}, {sig[4], sig[5], sig[2]}};
    double* out = s_c + (size_t)g * ND * RSTR + el;
#pragma unroll
    for (int k = 0; k < NNE; ++k)
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < DIM; ++j) t = fma(S[d][j], G[k][j], t);
        out[(k * DIM + d) * RSTR] = w * t;  // the element loop's fma(w, t, f) starts from f = 0: same rounding for g = 0 ...
      }
  }
  __syncthreads();
  // ... and accumulates in the order g = 0, 1, ...: f = fma(w, t, f) is reproduced as f + (w t) only up to one rounding,
  // so the sum below is the Gauss-point-parallel kernel's own fixed order (deterministic; within 1e-16 of k_elem_force)
  const int n_el = (int)((a.n_elems - e0 < EPB) ? (a.n_elems - e0) : EPB);
  for (int idx = tid; idx < n_el * ND; idx += T::THREADS) {
    const int le = idx / ND, kd = idx - le * ND;
    double f = 0.0;
#pragma unroll
    for (int gg = 0; gg < NGP; ++gg) f += s_c[((size_t)gg * ND + kd) * RSTR + le];
    a.fe[(e0 + le) * ND + kd] = f;
  }
}

template <class El>
int launch_residual(const ResArgs& a, const int64_t* node_ptr, const int32_t* node_inc, double* D, cudaStream_t stream) {
  if (a.n_elems == 0 || a.n_nodes == 0) return 0;
  if (int rc = ensure_device_tables()) return rc;
  // MEASURED (round 2, 8 M hex8 elements, sigma recomputed from U): thread per element 3.37 ms, Gauss-point-parallel
  // 4.21 ms -- the eight threads of an element each gather its 8 nodes and the transposition through shared memory costs
  // more than the registers it frees.  The element loop stays the default; FDK_ELEM_FORCE_GP=1 selects the other one.
  static const bool per_element = [] {
    const char* e = getenv("FDK_ELEM_FORCE_GP");
    return !(e && atoi(e) != 0);
  }();
  if (!per_element) {
    using T = ElemForceGp<El>;
    static thread_local bool attr_set = false;
    if (T::SMEM > 48 * 1024 && !attr_set) {
      FDK_CUDA(cudaFuncSetAttribute(k_elem_force_gp<El>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
      attr_set = true;
    }
    k_elem_force_gp<El><<<(unsigned)((a.n_elems + T::EPB - 1) / T::EPB), T::THREADS, T::SMEM, stream>>>(a);
  } else if constexpr (El::ID == FDK_HEX8) {
    // monomial-basis kernel (above); FDK_ELEM_FORCE_HEX8=0 selects the generic element loop, 2 / 3 / 4 the resident CTAs
    // per SM the registers are allocated for.  MEASURED (round 2, 8 M elements, residual alone incl. the node gather):
    // generic loop 3.37 ms; monomial basis 2.60 (254 registers) / 2.49 (168, 72 B of spills) / 2.55 ms (128, 232 B)
    static const int mode = [] {
      const char* e = getenv("FDK_ELEM_FORCE_HEX8");
      return e ? atoi(e) : 3;
    }();
    if (mode == 0) k_elem_force<El><<<(unsigned)((a.n_elems + 127) / 128), 128, 0, stream>>>(a);
    else if (mode == 3) k_elem_force_hex8<3><<<(unsigned)((a.n_elems + 127) / 128), 128, 0, stream>>>(a);
    else if (mode == 4) k_elem_force_hex8<4><<<(unsigned)((a.n_elems + 127) / 128), 128, 0, stream>>>(a);
    else k_elem_force_hex8<2><<<(unsigned)((a.n_elems + 127) / 128), 128, 0, stream>>>(a);
  } else
  k_elem_force<El><<<(unsigned)((a.n_elems + 127) / 128), 128, 0, stream>>>(a);
  k_node_force_gather<El::DIM><<<(unsigned)((a.n_nodes + 255) / 256), 256, 0, stream>>>(a.n_nodes, node_ptr, node_inc, a.fe, D);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

// Temperature and its gradient at the Gauss points (fedoo/weakform/heat_equation.py:64-70,149-152).
template <class El>
__global__ void __launch_bounds__(256) k_gp_temperature(const __grid_constant__ GpArgs a) {
  constexpr int NNE = El::NNE, NGP = El::NGP, DIM = El::DIM;
  const int64_t N = a.n_elems * NGP;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int g = (int)(n / a.n_elems);
  const int64_t e = n - (int64_t)g * a.n_elems;
  const ElemTable& tab = c_tab[El::ID];
  int nd[NNE];
  double X[NNE][DIM];
#pragma unroll
  for (int k = 0; k < NNE; ++k) {
    nd[k] = a.conn[e * NNE + k];
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[k][d] = a.coords[(int64_t)nd[k] * DIM + d];
  }
  double dN[DIM * NNE];
#pragma unroll
  for (int t = 0; t < DIM * NNE; ++t) dN[t] = tab.dN[g * DIM * NNE + t];
  double G[NNE][DIM];
  gp_geometry<NNE, DIM>(dN, 1.0, X, G);
  double Tg = 0.0, gT[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < NNE; ++k) {
    const double T = a.U[nd[k]];
    Tg = fma(tab.N[g * NNE + k], T, Tg);
#pragma unroll
    for (int d = 0; d < DIM; ++d) gT[d] = fma(T, G[k][d], gT[d]);
  }
  if (a.temp_gp != nullptr) a.temp_gp[n] = Tg;
  if (a.temp_grad_gp != nullptr) {
#pragma unroll
    for (int d = 0; d < 3; ++d) a.temp_grad_gp[(int64_t)d * N + n] = gT[d];
  }
}

template <class El>
int launch_gp_strain_stress(const GpArgs& a, cudaStream_t stream) {
  const int64_t N = a.n_elems * El::NGP;
  if (N == 0) return 0;
  if (int rc = ensure_device_tables()) return rc;
  const int64_t blocks = (N + 255) / 256;
  k_gp_strain_stress<El><<<(unsigned)blocks, 256, 0, stream>>>(a);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

template <class El>
int launch_gp_temperature(const GpArgs& a, cudaStream_t stream) {
  const int64_t N = a.n_elems * El::NGP;
  if (N == 0) return 0;
  if (int rc = ensure_device_tables()) return rc;
  const int64_t blocks = (N + 255) / 256;
  k_gp_temperature<El><<<(unsigned)blocks, 256, 0, stream>>>(a);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------
// J2 plasticity: radial return + consistent tangent, one thread per Gauss point.
// props = [E, nu, alpha, sigmaY, k, m]; hardening R(p) = k p^m (EPICP,
// fedoo/constitutivelaw/simcoon_umat.py:103-112); yield f = q - sigmaY - R(p)
// (fedoo/constitutivelaw/elasto_plasticity.py:154-155); trial state sigma = H (eps - eps_p)
// (:317-330); flow direction n = 3/2 s/q with engineering shear doubling (:160-164).
// The legacy per-GP Python cutting-plane loop (:338-371) is replaced by a safeguarded scalar
// Newton on g(dp) = q_tr - 3 mu dp - sigmaY - k (p0+dp)^m (same fixed point for J2 + isotropic
// hardening), slope -(3 mu + R'(p)).
// ---------------------------------------------------------------------------------------
struct J2Args {
  int64_t n_gp;
  double E, nu, sigY, k, m;
  const double* strain;
  const double* statev0;
  double* stress;
  double* statev;
  double* tangent;
  double* tangent_r1;  // structured form of the same tangent, J2_R1 doubles per Gauss point (see below), or NULL
  int continuum;  // 1: continuum elastoplastic tangent at the end state (beta = 1), 0: consistent (algorithmic) tangent
};

// Both J2 tangents are an isotropic tensor minus a rank-one term on the unit deviatoric direction n^:
//     C = lam' 1 (x) 1 + 2 mu' I_sym - kappa n^ (x) n^,   lam' = K - 2 mu beta / 3, mu' = mu beta, kappa = 2 mu gamma
// (elastic points: beta = 1, gamma = 0).  Ten doubles per Gauss point carry it -- [lam', mu', kappa, n^ (6, stress Voigt),
// 0] -- instead of the 36 of the (6,6,N) array the reference's protocol stores (simcoon_umat.py:556-580): 80 instead of
// 288 bytes written by the update and read back by the assembly, and 10 instead of 36 shared-memory operands per
// (element, Gauss point) in the block phase (csrc/fdk_assemble_iso.cuh, PHYS_R1).

// radial return + tangent of ONE Gauss point n with total strain eps (the body shared by k_j2_update and k_j2_update_u)
__device__ __forceinline__ void j2_point(const J2Args& a, int64_t n, const double (&eps)[6]) {
  const double mu = 0.5 * a.E / (1.0 + a.nu);
  const double lam = a.E * a.nu / ((1.0 + a.nu) * (1.0 - 2.0 * a.nu));
  const double kb = lam + 2.0 * mu / 3.0;  // bulk modulus
  double sv[8];
  {
    const double2* sp = reinterpret_cast<const double2*>(a.statev0 + 8 * n);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 v = sp[i];
      sv[2 * i] = v.x;
      sv[2 * i + 1] = v.y;
    }
  }
  const double p0 = sv[1];
  double ee[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) ee[i] = eps[i] - sv[2 + i];
  const double tr = ee[0] + ee[1] + ee[2];
  double sig[6];
  sig[0] = lam * tr + 2.0 * mu * ee[0];
  sig[1] = lam * tr + 2.0 * mu * ee[1];
  sig[2] = lam * tr + 2.0 * mu * ee[2];
  sig[3] = mu * ee[3];
  sig[4] = mu * ee[4];
  sig[5] = mu * ee[5];
  const double pm = (sig[0] + sig[1] + sig[2]) / 3.0;
  double s[6] = {sig[0] - pm, sig[1] - pm, sig[2] - pm, sig[3], sig[4], sig[5]};
  const double q = sqrt(1.5 * (s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + 2.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5])));
  // x^m as exp(m log x): half the cost of pow(), and the logarithm is all the iteration needs beside it
  const double R0 = (p0 > 0.0) ? a.k * exp(a.m * log(p0)) : 0.0;
  const double ftr = q - a.sigY - R0;
  double dp = 0.0;
  double Rp = 0.0;
  const bool plastic = ftr > 0.0;
  if (plastic) {
    // root of f(x) = q - 3 mu x - sigY - k (p0 + x)^m on (0, hi]: f is convex and decreasing, so Newton from a point where
    // f > 0 climbs monotonically to the root; from f < 0 it lands left of it first.  Two upper bounds of the root:
    // 3 mu x <= f_trial, and k (p0 + x)^m - R0 <= f_trial (the second one is the tight one just past first yield, where
    // the hardening slope k m p^(m-1) is huge).  Safeguarded by bisection on [lo, hi].
    const double m3 = 3.0 * mu;
    double hi = ftr / m3;
    if (a.k > 0.0) {
      const double h2 = exp(log((ftr + R0) / a.k) / a.m) - p0;
      if (h2 > 0.0 && h2 < hi) hi = h2;
    }
    double lo = 0.0, x = hi, dR = 1e300;
    for (int it = 0; it < 60; ++it) {
      const double pp = p0 + x;
      const double pw = (pp > 0.0) ? exp(a.m * log(pp)) : 0.0;
      const double fx = q - m3 * x - a.sigY - a.k * pw;
      if (fx < 0.0) hi = fmin(hi, x);
      if (fx > 0.0) lo = fmax(lo, x);
      dR = (pp > 0.0) ? a.k * a.m * pw / pp : 1e300;
      double xn = x + fx / (m3 + dR);
      if (!(xn >= lo && xn <= hi) || !isfinite(xn)) xn = 0.5 * (lo + hi);
      const bool done = fabs(xn - x) <= 1e-14 * fmax(fabs(xn), 1e-300);
      x = xn;
      if (done) break;
    }
    dp = x;
    Rp = dR;  // k m p^(m-1) at the last iterate: equal to the end-state slope to the iteration's own tolerance
  }
  const double qs = (q > 0.0) ? q : 1.0;
  double nf[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) nf[i] = 1.5 * s[i] / qs;
  double out_s[6], out_v[8];
#pragma unroll
  for (int i = 0; i < 6; ++i) out_s[i] = sig[i] - 2.0 * mu * dp * nf[i];
  out_v[0] = sv[0];
  out_v[1] = p0 + dp;
#pragma unroll
  for (int i = 0; i < 3; ++i) out_v[2 + i] = sv[2 + i] + dp * nf[i];
#pragma unroll
  for (int i = 3; i < 6; ++i) out_v[2 + i] = sv[2 + i] + 2.0 * dp * nf[i];
  {
    double2* so = reinterpret_cast<double2*>(a.stress + 6 * n);
    double2* vo = reinterpret_cast<double2*>(a.statev + 8 * n);
#pragma unroll
    for (int i = 0; i < 3; ++i) so[i] = make_double2(out_s[2 * i], out_s[2 * i + 1]);
#pragma unroll
    for (int i = 0; i < 4; ++i) vo[i] = make_double2(out_v[2 * i], out_v[2 * i + 1]);
  }
  if (a.tangent != nullptr || a.tangent_r1 != nullptr) {
    // C = K 1(x)1 + 2 mu beta I_dev - 2 mu gamma n^(x)n^   (elastic: beta = 1, gamma = 0)
    double beta = 1.0, gam = 0.0;
    if (plastic) {
      beta = a.continuum ? 1.0 : 1.0 - 3.0 * mu * dp / q;
      gam = 1.0 / (1.0 + Rp / (3.0 * mu)) - (1.0 - beta);
    }
    const double sc = sqrt(1.5) / qs;
    double nh[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) nh[i] = s[i] * sc;
    const double m2b = 2.0 * mu * beta, m2g = 2.0 * mu * gam;
    if (a.tangent_r1 != nullptr) {
      double2* ro = reinterpret_cast<double2*>(a.tangent_r1 + (int64_t)J2_R1 * n);
      ro[0] = make_double2(kb - m2b / 3.0, 0.5 * m2b);
      ro[1] = make_double2(m2g, nh[0]);
      ro[2] = make_double2(nh[1], nh[2]);
      ro[3] = make_double2(nh[3], nh[4]);
      ro[4] = make_double2(nh[5], 0.0);
    }
    if (a.tangent == nullptr) return;
    double2* to = reinterpret_cast<double2*>(a.tangent + 36 * n);
#pragma unroll
    for (int j = 0; j < 6; ++j) {  // column j (Fortran order: C_ij at i + 6 j)
      double col[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        double v = -m2g * nh[i] * nh[j];
        if (i < 3 && j < 3) v += kb - m2b / 3.0;
        if (i == j) v += (i < 3) ? m2b : 0.5 * m2b;
        col[i] = v;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) to[3 * j + i] = make_double2(col[2 * i], col[2 * i + 1]);
    }
  }
}


__global__ void __launch_bounds__(256) k_j2_update(const __grid_constant__ J2Args a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.n_gp) return;
  double eps[6];
  const double2* ep = reinterpret_cast<const double2*>(a.strain + 6 * n);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double2 v = ep[i];
    eps[2 * i] = v.x;
    eps[2 * i + 1] = v.y;
  }
  j2_point(a, n, eps);
}

// The same update with the strain taken straight from the dof vector: geometry at the Gauss point, grad u, Voigt strain
// (what k_gp_strain_stress writes to HBM for the law to read back -- 48 bytes per Gauss point each way, and a launch)
// and the radial return in one pass.  sv["Strain"] stays lazy (fedoo_b200/assembly.py:_LazyStrain).  3D elements.
struct J2UArgs {
  J2Args j2;
  int n_nodes;
  int64_t n_elems;
  const int32_t* conn;
  const double* coords;
  const double* U;
};

template <class El>
__global__ void __launch_bounds__(256) k_j2_update_u(const __grid_constant__ J2UArgs a) {
  constexpr int NNE = El::NNE, NGP = El::NGP, DIM = El::DIM;
  static_assert(DIM == 3, "J2 update: 3D");
  const int64_t N = a.n_elems * NGP;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int g = (int)(n / a.n_elems);
  const int64_t e = n - (int64_t)g * a.n_elems;
  const ElemTable& tab = c_tab[El::ID];
  int nd[NNE];
  double X[NNE][DIM];
#pragma unroll
  for (int k = 0; k < NNE; ++k) {
    nd[k] = a.conn[e * NNE + k];
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[k][d] = a.coords[(int64_t)nd[k] * DIM + d];
  }
  double G[NNE][DIM];
  gp_geometry<NNE, DIM>(tab.dN + g * DIM * NNE, 1.0, X, G);
  double gu[DIM][DIM];
#pragma unroll
  for (int v = 0; v < DIM; ++v)
#pragma unroll
    for (int d = 0; d < DIM; ++d) gu[v][d] = 0.0;
#pragma unroll
  for (int k = 0; k < NNE; ++k)
#pragma unroll
    for (int v = 0; v < DIM; ++v) {
      const double u = a.U[(int64_t)v * a.n_nodes + nd[k]];
#pragma unroll
      for (int d = 0; d < DIM; ++d) gu[v][d] = fma(u, G[k][d], gu[v][d]);
    }
  double eps[6];
  voigt_strain<DIM>(gu, eps);
  j2_point(a.j2, n, eps);
}

template <class El>
int launch_j2_update_u(const J2UArgs& a, cudaStream_t stream) {
  const int64_t N = a.n_elems * El::NGP;
  if (N == 0) return 0;
  if (int rc = ensure_device_tables()) return rc;
  k_j2_update_u<El><<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(a);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

// (6,6,N) array from the structured form (for whoever reads sv["TangentMatrix"], and for the kernels that take the full
// tangent): the same expression as in k_j2_update
__global__ void __launch_bounds__(256) k_j2_tangent_expand(int64_t n_gp, const double* __restrict__ r1, double* __restrict__ tangent) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_gp) return;
  const double* r = r1 + (int64_t)J2_R1 * n;
  const double lamp = r[0], m2b = 2.0 * r[1], kap = r[2];
  double nh[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) nh[i] = r[3 + i];
  double2* to = reinterpret_cast<double2*>(tangent + 36 * n);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double col[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double v = -kap * nh[i] * nh[j];
      if (i < 3 && j < 3) v += lamp;
      if (i == j) v += (i < 3) ? m2b : 0.5 * m2b;
      col[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) to[3 * j + i] = make_double2(col[2 * i], col[2 * i + 1]);
  }
}

__global__ void k_gather_f64(int64_t n, const int64_t* __restrict__ index, const double* __restrict__ src,
                             double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[index[i]];
}

__global__ void k_scatter_add_f64(int64_t n, const int64_t* __restrict__ index, const double* __restrict__ src,
                                  double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[index[i]] += src[i];  // indices are unique within one call (owner rows)
}

// dst[seg_dst[s] + k] = src[seg_src[s] + k], k < seg_len[s]: the pack / unpack of the residual exchange when the owned
// dofs of a rank are contiguous runs (slab partitions): one launch, no index arrays, no zero-fill
__global__ void __launch_bounds__(256) k_copy_segments(int n_seg, const int64_t* __restrict__ seg_src,
                                                       const int64_t* __restrict__ seg_dst,
                                                       const int64_t* __restrict__ seg_len, const double* __restrict__ src,
                                                       double* __restrict__ dst) {
  const int s = blockIdx.y;
  if (s >= n_seg) return;
  const int64_t len = seg_len[s];
  const double* a = src + seg_src[s];
  double* b = dst + seg_dst[s];
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < len; k += (int64_t)gridDim.x * blockDim.x) b[k] = a[k];
}

}  // namespace fdk
