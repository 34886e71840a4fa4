// fdk_assemble.cuh -- owner-computes cluster kernels for K (CSR values) and D (global vector).
//
// One CTA per node cluster.  The cluster OWNS a compact set of nodes, hence the nvar CSR rows
// of each of them, and computes every contribution to those rows itself:
//
//   phase 0  stage tables, coordinates, dof values, local connectivity and the slot offsets of
//            the cluster in shared memory; every thread prefetches the descriptor of "its"
//            incidence (owned node I, element e containing I) into registers
//   phase 1  one task per (touched element, Gauss point): Jacobian, inverse, |det J| w and
//            dN/dx (and w*sigma for the residual) -> shared memory, computed ONCE per cluster
//            and reused by all incidences (the reference caches these as sparse operators,
//            fedoo/core/assembly.py:776-928; here they never leave the SM)
//   phase 2  one thread per incidence, in ELEMENT-major order so that the lanes of one element
//            read the same dN/dx rows (shared-memory broadcast, 128-bit loads): accumulates in
//            registers, over the Gauss points, the nne blocks S_IJ = sum_g w G_I (x) G_J
//            (isotropic closed form), B_I^T C_g B_J (general tangent) or the scalar conduction
//            term (heat) -- the batched A^T diag(c) B of fedoo/core/_sparsematrix.py:83-89 --
//            and the nodal force B_I^T sigma (fedoo/core/assembly.py:400-411); then scatters the
//            blocks into a staging array SORTED BY CSR SLOT (positions precomputed by the plan)
//   phase 3  slot-centric gather: one thread per CSR block slot (I, J) sums the contiguous run
//            of staged blocks of the elements shared by I and J in a fixed order (the
//            cluster-local analogue of the reference's Matrix_convertCOOtoCSR SpMV,
//            fedoo/core/_sparsematrix.py:256-302), applies the constitutive closed form and
//            writes the nvar x nvar scalars to their final positions of the variable-major tiled
//            CSR (scipy.sparse.bmat layout).  Slots with many contributions (the diagonal) are
//            pre-reduced by a lane-balanced pass so that the main pass diverges little.
//
// No floating-point atomics anywhere: each K value and each D entry is written exactly once,
// so results are bit-reproducible run to run.  Rows are never exchanged between CTAs (or GPUs).
#pragma once
#include "fdk_common.cuh"
#include "fdk_tables.cuh"

namespace fdk {

enum Physics { PHYS_ISO = 0, PHYS_GENERAL = 1, PHYS_HEAT = 2, PHYS_R1 = 3 };  // R1: structured J2 tangent (balanced kernel)

constexpr int HEAVY_T = 4;  // slots with more contributions are pre-reduced (must match plan.py)

// Balanced hex8 kernel (fdk_assemble_iso.cuh), 4 threads per incidence: which two column blocks thread `part` takes.
// ADJ = 1: j = 2 part + jj (6 contiguous doubles of dN/dx: three 128-bit loads per Gauss point);
// ADJ = 0: j = part + 4 jj (six 64-bit loads).  The staging colouring (fdk_color.cuh) follows the same rule.
#ifndef FDK_ISO_COLS_ADJ
#define FDK_ISO_COLS_ADJ 1
#endif
constexpr bool ISO_COLS_ADJ = FDK_ISO_COLS_ADJ != 0;
__host__ __device__ constexpr int iso_col(int part, int jj) { return ISO_COLS_ADJ ? 2 * part + jj : part + 4 * jj; }
__host__ __device__ constexpr int iso_jj(int j) { return ISO_COLS_ADJ ? (j & 1) : (j >> 2); }

struct AsmArgs {
  fdk_plan p;
  const double* coords;
  const double* U;           // elastic: dof vector; heat: T
  const double* U2;          // heat: T_start
  const double* stress_gp;   // optional given stress (6,N)
  const double* tangent_gp;  // optional per-GP tangent (6,6,N)
  const double* tangent_r1;  // PHYS_R1: structured J2 tangent, J2_R1 doubles per Gauss point
  double* K;
  double* D;
  double lam, mu;
  double C[36];     // uniform tangent, row-major
  double cond[9];   // conductivity, row-major
  double rcdt;      // rho c / dt
  int compute;
  int big_doubles;  // size of the aliased geometry / staging region
  int fuse_ku;      // linear law, K and D both requested: D = -(assembled rows) . U in the gather phase
  int no_mma;       // FDK_NO_MMA=1: keep the CUDA-core producer even where the tensor-core one applies
  // fused residual exchange (balanced kernel, multi-GPU): besides D (rank-local numbering) every owned entry is
  // stored at var * n_dst_nodes + node_gid[node] of n_dst destination vectors -- ONE NVLink multicast address
  // (NVSwitch replicates the store into every GPU's copy of the global vector) or the peers' own addresses
  double* D_dst[8];
  int n_dst;
  const int64_t* node_gid;  // rank-local node -> global node
  int64_t n_dst_nodes;
};

template <class El, int PHYS>
struct Layout {
  static constexpr int NNE = El::NNE, NGP = El::NGP, DIM = El::DIM;
  static constexpr int NV = (PHYS == PHYS_HEAT) ? 1 : DIM;    // variables per node
  static constexpr int NU = (PHYS == PHYS_HEAT) ? 2 : DIM;    // staged nodal values per touched node
  static constexpr int BLK = NV * NV;                         // scalars per (I,J) block
  static constexpr int ISTR = (NNE * BLK) | 1;                // staging stride of one incidence (its nne blocks):
                                                              // odd, so that the lanes of a store hit distinct banks
  static constexpr int NSIG = (PHYS == PHYS_HEAT) ? DIM + 1 : (DIM == 3 ? 6 : 3);
  static constexpr int GROW = DIM * NNE;                      // dN/dx of one (element, gp): [k][d]
  static_assert(GROW % 2 == 0, "128-bit rows");
  // shared-memory strides chosen so that (stride / 16 B) is odd: the 128-bit rows of the 8 Gauss
  // points of an element, and of consecutive elements, start in different bank groups
  static constexpr int GSTR = ((GROW + 2) / 2) % 2 ? GROW + 2 : GROW + 4;
  static constexpr int ESTR = ((NGP * GSTR) / 2) % 2 ? NGP * GSTR : NGP * GSTR + 2;
  static constexpr int WSTR = NGP | 1;                        // w_g |det J| per element
  static constexpr int SSTR = (NGP * NSIG) | 1;               // w sigma per element
  static constexpr int TSTR = GROW | 1;                       // padded dN table row
  static constexpr int TAB_DOUBLES = (NGP * TSTR + NGP * NNE + NGP + 1) & ~1;

  // doubles of the geometry view of the big region (w*sigma only on the B^T sigma residual path)
  static long geo_doubles(const fdk_plan& p, bool bts) { return (long)p.cap_te * (ESTR + WSTR + (bts ? SSTR : 0)); }
  // doubles of the staging view
  static long stage_doubles(const fdk_plan& p, bool mma = false) {
    const long nf = p.cap_inc > p.cap_slots ? p.cap_inc : p.cap_slots;  // nodal forces / per-slot K.u products
    return (long)p.cap_inc * ISTR + (mma ? 0 : nf * NV);                 // (tensor-core path: products reuse sJ)
  }

  static constexpr int JSTR = NGP * 10;  // tensor-core path: inverse Jacobian (9) + w per Gauss point

  static size_t smem_bytes(const fdk_plan& p, bool bts, bool mma, int* big_doubles) {
    long big = mma ? 0 : geo_doubles(p, bts);
    const long s = stage_doubles(p, mma);
    if (s > big) big = s;
    big = (big + 1) & ~1L;
    *big_doubles = (int)big;
    long doubles = TAB_DOUBLES + 2 * (((long)p.cap_tn * (DIM + NU) + 1) & ~1L) + big + p.cap_owned;  // + sBptr
    if (mma) {  // sJ lives beside the staging region (not aliased); the K.u products reuse it in phase 3
      const long nj = (long)p.cap_te * JSTR, nr = (long)p.cap_slots * NV;
      doubles += (nj > nr ? nj : nr) + 1;
    }
    size_t bytes = (size_t)doubles * 8;
    bytes += (size_t)((p.cap_ent + 3) & ~3) * 2;      // sEnt
    if (mma) bytes += (size_t)p.cap_te * 4;           // sTe
    bytes += (size_t)(2 * (p.cap_owned + 1)) * 4;     // sSlotBase, sFinc
    bytes += (size_t)(p.cap_slots + 1) * 4;           // sRec
    bytes += (size_t)p.cap_heavy * 4;                 // sHeavy
    bytes += 2 * (size_t)((p.cap_te * NNE + 7) & ~3); // sLconn (double-buffered; + one word: unaligned starts)
    return bytes;
  }
};

template <int DIM>
__device__ __forceinline__ double invert(const double (&J)[DIM][DIM], double (&iJ)[DIM][DIM]);

template <>
__device__ __forceinline__ double invert<3>(const double (&J)[3][3], double (&iJ)[3][3]) {
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  const double r = 1.0 / det;
  iJ[0][0] = c00 * r;
  iJ[1][0] = c01 * r;
  iJ[2][0] = c02 * r;
  iJ[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
  iJ[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
  iJ[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
  iJ[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
  iJ[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
  iJ[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
  return det;
}

template <>
__device__ __forceinline__ double invert<2>(const double (&J)[2][2], double (&iJ)[2][2]) {
  const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  const double r = 1.0 / det;
  iJ[0][0] = J[1][1] * r;
  iJ[0][1] = -J[0][1] * r;
  iJ[1][0] = -J[1][0] * r;
  iJ[1][1] = J[0][0] * r;
  return det;
}

// Geometry of one (element, Gauss point): G[k][d] = dN_k/dx_d and w = w_g |det J|.
// J[r][x] = sum_k dN_k/dxi_r X_k[x]  (fedoo/lib_elements/element_base.py:43-50),
// G = J^-1 dN/dxi (fedoo/core/assembly.py:879-881).
template <int NNE, int DIM>
__device__ __forceinline__ double gp_geometry(const double* __restrict__ dN /*[DIM][NNE]*/, double wg,
                                              const double (&X)[NNE][DIM], double (&G)[NNE][DIM]) {
  double J[DIM][DIM];
#pragma unroll
  for (int r = 0; r < DIM; ++r)
#pragma unroll
    for (int x = 0; x < DIM; ++x) J[r][x] = 0.0;
#pragma unroll
  for (int k = 0; k < NNE; ++k)
#pragma unroll
    for (int r = 0; r < DIM; ++r) {
      const double dn = dN[r * NNE + k];
#pragma unroll
      for (int x = 0; x < DIM; ++x) J[r][x] = fma(dn, X[k][x], J[r][x]);
    }
  double iJ[DIM][DIM];
  const double det = invert<DIM>(J, iJ);
#pragma unroll
  for (int k = 0; k < NNE; ++k)
#pragma unroll
    for (int x = 0; x < DIM; ++x) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < DIM; ++r) s = fma(iJ[x][r], dN[r * NNE + k], s);
      G[k][x] = s;
    }
  return wg * fabs(det);
}

// Voigt strain from the displacement gradient g[a][b] = du_a/dx_b (engineering shears).
template <int DIM>
__device__ __forceinline__ void voigt_strain(const double (&g)[DIM][DIM], double (&e)[6]) {
  e[0] = g[0][0];
  e[1] = g[1][1];
  e[3] = g[0][1] + g[1][0];
  if constexpr (DIM == 3) {
    e[2] = g[2][2];
    e[4] = g[0][2] + g[2][0];
    e[5] = g[1][2] + g[2][1];
  } else {
    e[2] = 0.0;
    e[4] = 0.0;
    e[5] = 0.0;
  }
}

// sigma_i = sum_j C_ij eps_j with C addressed as C[i*si + j*sj].
__device__ __forceinline__ void apply_tangent(const double* __restrict__ C, int si, int sj, const double (&e)[6],
                                              double (&s)[6]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) acc = fma(C[i * si + j * sj], e[j], acc);
    s[i] = acc;
  }
}

// Ampere-style asynchronous global -> shared copy (LDGSTS); BYTES = 4, 8 or 16, both sides aligned.
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Optional phase clocks (-DFDK_PHASE_CLOCKS): SM cycles spent by each CTA between its barriers, summed over
// all CTAs; read back by fdk_debug_phase_clocks.  Compiled out of the production library.
#ifdef FDK_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[16];
#define FDK_CLK_DECL long long _clk_t = clock64(); unsigned long long _clk_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define FDK_CLK(i) { const long long _n = clock64(); _clk_acc[i] += (unsigned long long)(_n - _clk_t); _clk_t = _n; }
#define FDK_CLK_FLUSH if (threadIdx.x == 0) { for (int _i = 0; _i < 10; ++_i) atomicAdd(&g_phase_clk[_i], _clk_acc[_i]); }
#else
#define FDK_CLK_DECL
#define FDK_CLK(i)
#define FDK_CLK_FLUSH
#endif

// Per-cluster header (fdk_plan::cl_hdr, 16 int32 per cluster): one 64-byte load instead of a chain of
// dependent loads through the individual range arrays.
struct ClusterHdr {
  int q0, n_owned, te0, n_te, tn0, n_tn, inc0, n_inc, h0, n_heavy, n_slots, ent0;
  int64_t slot0;
};
__device__ __forceinline__ ClusterHdr load_hdr(const int32_t* __restrict__ hdr, int c) {
  const int4* h4 = reinterpret_cast<const int4*>(hdr + (int64_t)c * 16);
  const int4 a = h4[0], b = h4[1], d = h4[2], e = h4[3];
  ClusterHdr h;
  h.q0 = a.x; h.n_owned = a.y; h.te0 = a.z; h.n_te = a.w;
  h.tn0 = b.x; h.n_tn = b.y; h.inc0 = b.z; h.n_inc = b.w;
  h.h0 = d.x; h.n_heavy = d.y;
  h.slot0 = (int64_t)(((uint64_t)(uint32_t)d.w << 32) | (uint32_t)d.z);
  h.n_slots = e.x;
  h.ent0 = e.y;
  return h;
}

// One m8n8k4 FP64 tensor-core MMA (SASS DMMA.8x8x4): D(8x8) += A(8x4) B(4x8).  Fragments: a = A[lane>>2][lane&3],
// b = B[lane&3][lane>>2], d0/d1 = D[lane>>2][2*(lane&3) + 0/1].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// MMA = true (hex8, isotropic law, matrix requested, no B^T sigma pass): phases 1-2 are replaced by
//   phase 1m  one task per (touched element, gp): inverse Jacobian and w only -> shared memory
//   phase 2m  one WARP per touched element: every lane rebuilds "its" six dN/dx values
//             (node lane>>2, Gauss points lane&3 and (lane&3)+4) in registers and the element matrix
//             S^e = sum_g w G^T G (24 x 24, component-major) is formed by 12 DMMA.8x8x4 on the six
//             upper (c <= a) 8x8 tiles -- operands never touch shared memory; the fragments are
//             scattered to the same slot-sorted staging entries as in the CUDA-core path.
template <class El, int PHYS, int THREADS, int MINB, bool MMA>
__global__ void __launch_bounds__(THREADS, MINB) k_assemble(const __grid_constant__ AsmArgs a) {
  static_assert(!MMA || (El::NNE == 8 && El::NGP == 8 && El::DIM == 3 && PHYS == PHYS_ISO), "MMA path: hex8 isotropic");
  using L = Layout<El, PHYS>;
  constexpr int NNE = L::NNE, NGP = L::NGP, DIM = L::DIM, NV = L::NV, NU = L::NU, BLK = L::BLK, ISTR = L::ISTR;
  constexpr int NSIG = L::NSIG, GROW = L::GROW, GSTR = L::GSTR, ESTR = L::ESTR, WSTR = L::WSTR, SSTR = L::SSTR;
  constexpr int TSTR = L::TSTR;
  constexpr int HT = THREADS / 2;  // incidences per cluster <= HT: two threads per incidence in phase 2
  constexpr int NH = NNE / 2;      // column blocks per thread
  static_assert(NNE % 2 == 0 && HT % 32 == 0, "half rows, warp-uniform halves");
  const fdk_plan& p = a.p;
  const int tid = threadIdx.x;
  const bool do_mat = (a.compute & FDK_MATRIX) != 0;
  const bool do_vec = (a.compute & FDK_VECTOR) != 0;
  const bool fuse_ku = a.fuse_ku != 0;           // residual from the assembled rows (phase 3)
  const bool do_bts = do_vec && !fuse_ku;        // residual as B^T sigma (phases 1-2)

  // ---- persistent CTA: clusters blockIdx.x, blockIdx.x + gridDim.x, ... ----
  const int half = tid / HT;       // warp-uniform
  // incidence of this thread (phase 2).  The second half is rotated by two warps: a cluster usually has far fewer
  // incidences than HT, so only the first warps of each half have work -- unrotated, those are warps 0, 1 and HT/32,
  // HT/32 + 1, which sit on the SAME two of the SM's four sub-partitions (warp id mod 4) and leave the FP64 pipes of
  // the other two idle through the whole block phase (tet10: 62 incidences -> warps {0, 1, 4, 5} of 8)
  static_assert(HT % 64 == 0, "rotation by two warps");
  const int it = MMA ? tid - half * HT : (half == 0 ? tid : (tid - HT + 64) % HT);
  const int j0 = half * NH;        // its column blocks: j0 .. j0 + NH - 1

  extern __shared__ __align__(16) double smem[];
  double* sdN = smem;
  double* sN = sdN + NGP * TSTR;
  double* sW = sN + NGP * NNE;
  // phase-1 inputs are double-buffered: the next cluster's coordinates / dofs / connectivity are
  // fetched (cp.async) while the current cluster computes
  const int xu_doubles = (p.cap_tn * (DIM + NU) + 1) & ~1;
  double* sXbuf = smem + L::TAB_DOUBLES;  // [2][xu_doubles]
  double* sBig = sXbuf + 2 * xu_doubles;
  // geometry view
  double* sG = sBig;                               // [n_te][ESTR]
  double* sWd = sG + (long)p.cap_te * ESTR;        // [n_te][WSTR]      w_g |det J|
  double* sSig = sWd + (long)p.cap_te * WSTR;      // [n_te][SSTR]      w * sigma (B^T sigma path only)
  // staging view (aliases the geometry once phase 2 has read it)
  double* sBlk = sBig;                             // [cap_inc][ISTR]: the nne blocks of every incidence
  double* sF = sBlk + (long)p.cap_inc * ISTR;      // [max(cap_inc, cap_slots)][NV]
  double* sR = sF;                                 // per-slot K.u products (fuse_ku: sF is unused)
  long long* sBptr = reinterpret_cast<long long*>(sBig + a.big_doubles);   // [cap_owned]
  int* sSlotBase = reinterpret_cast<int*>(sBptr + p.cap_owned);            // [cap_owned+1]
  int* sFinc = sSlotBase + (p.cap_owned + 1);                              // [cap_owned+1]
  unsigned* sRec = reinterpret_cast<unsigned*>(sFinc + (p.cap_owned + 1)); // [cap_slots+1]
  unsigned* sHeavy = sRec + (p.cap_slots + 1);                             // [cap_heavy]
  unsigned char* sLbuf = reinterpret_cast<unsigned char*>(sHeavy + p.cap_heavy);  // [2][lc_bytes]
  const int lc_bytes = (p.cap_te * NNE + 7) & ~3;
  // tensor-core path extras
  [[maybe_unused]] double* sJ = reinterpret_cast<double*>(sBptr + ((p.cap_owned + 1) & ~1));  // [cap_te][JSTR]
  unsigned short* sEnt = reinterpret_cast<unsigned short*>(sLbuf + 2 * lc_bytes);  // [cap_ent] slot-sorted sources
  [[maybe_unused]] unsigned* sTe = nullptr;  // per touched element: first thread | owned-node mask << 16
  if constexpr (MMA) {
    sR = sJ;  // sJ is dead once phase 2m is over
    const long nj = (long)p.cap_te * L::JSTR, nr = (long)p.cap_slots * NV;
    sSlotBase = reinterpret_cast<int*>(sJ + (nj > nr ? nj : nr));
    sFinc = sSlotBase + (p.cap_owned + 1);
    sRec = reinterpret_cast<unsigned*>(sFinc + (p.cap_owned + 1));
    sHeavy = sRec + (p.cap_slots + 1);
    sLbuf = reinterpret_cast<unsigned char*>(sHeavy + p.cap_heavy);
    sTe = reinterpret_cast<unsigned*>(sLbuf + 2 * lc_bytes);
    sEnt = reinterpret_cast<unsigned short*>(sTe + p.cap_te);
  }

  // ---------------- prologue: tables (once per CTA) and the first cluster's phase-1 inputs ----------------
  {
    const ElemTable& tab = c_tab[El::ID];
    for (int t = tid; t < NGP * GROW; t += THREADS) {
      const int g = t / GROW, r = t - g * GROW;
      sdN[g * TSTR + r] = tab.dN[t];
    }
    for (int t = tid; t < NGP * NNE; t += THREADS) sN[t] = tab.N[t];
    if (tid < NGP) sW[tid] = tab.w[tid];
  }
  const bool need_u = do_vec && a.U != nullptr;
  constexpr int RT = (256 + THREADS - 1) / THREADS;  // cap_tn <= 256
  int node_r[RT];
  // issue the phase-1 inputs of one cluster into buffer b (node ids already in registers)
  auto fetch_inputs = [&](const ClusterHdr& h, int b) {
    double* dX = sXbuf + b * xu_doubles;
    double* dU = dX + p.cap_tn * DIM;
    unsigned char* dL = sLbuf + b * lc_bytes;
    const unsigned char* lc = p.cl_lconn + (int64_t)h.te0 * NNE;
    if constexpr (NNE % 4 == 0) {
      for (int t = tid; t < h.n_te * NNE / 4; t += THREADS) cp_async<4>(dL + 4 * t, lc + 4 * t);
    } else {
      // tet10: a cluster's byte range starts anywhere.  The aligned words that cover it are copied asynchronously (the
      // buffer keeps the misalignment, readers add it back; the plan pads the array by one word) -- plain byte loads
      // here made every cluster wait one global-memory latency between its block phase and its staging stores
      const int sh = (int)(((int64_t)h.te0 * NNE) & 3);
      const int n_w = (sh + h.n_te * NNE + 3) / 4;
      for (int t = tid; t < n_w; t += THREADS) cp_async<4>(dL + 4 * t, lc - sh + 4 * t);
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int t = tid + r * THREADS;
      const int node = node_r[r];
      if (node >= 0) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) cp_async<8>(dX + t * DIM + d, a.coords + (int64_t)node * DIM + d);
        if (need_u) {
          if constexpr (PHYS == PHYS_HEAT) {
            const double T = a.U[node];
            dU[t * 2 + 0] = T;
            dU[t * 2 + 1] = T - (a.U2 ? a.U2[node] : 0.0);
          } else {
#pragma unroll
            for (int v = 0; v < DIM; ++v) cp_async<8>(dU + t * DIM + v, a.U + (int64_t)v * p.n_nodes + node);
          }
        }
      }
    }
    cp_async_commit();
  };
  auto load_node_ids = [&](const ClusterHdr& h) {
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int t = tid + r * THREADS;
      node_r[r] = (t < h.n_tn) ? p.cl_tn_node[h.tn0 + t] : -1;
    }
  };
  ClusterHdr cur = load_hdr(p.cl_hdr, blockIdx.x);
  load_node_ids(cur);
  fetch_inputs(cur, 0);

  int buf = 0;
  FDK_CLK_DECL
  for (int c = blockIdx.x; c < p.n_clusters; c += gridDim.x, buf ^= 1) {
  const int q0 = cur.q0, n_owned = cur.n_owned, te0 = cur.te0, n_te = cur.n_te, n_inc = cur.n_inc;
  const int inc0 = cur.inc0, h0 = cur.h0, n_heavy = cur.n_heavy, n_slots = cur.n_slots;
  const int64_t slot0 = cur.slot0;
  const double* sX = sXbuf + buf * xu_doubles;
  const double* sU = sX + p.cap_tn * DIM;
  const unsigned char* sLconn = sLbuf + buf * lc_bytes + (NNE % 4 == 0 ? 0 : (int)(((int64_t)te0 * NNE) & 3));
  const int c_next = c + gridDim.x;
  const bool has_next = c_next < p.n_clusters;
  ClusterHdr nxt = cur;
  if (has_next) nxt = load_hdr(p.cl_hdr, c_next);  // consumed after phase 1

  // ---------------- phase 0: this cluster's phase-2/3 descriptors (latency hidden behind phase 1) ----------------
  unsigned my_desc = 0, my_fdst = 0;
  if (!MMA && it < n_inc) {
    my_desc = p.inc_desc[inc0 + it];
    if (do_bts) my_fdst = p.inc_fdst[inc0 + it];
  }
  {
    const unsigned* rec = p.slot_rec + slot0 + c;
    for (int t = tid; t <= n_slots; t += THREADS) cp_async<4>(sRec + t, rec + t);
    for (int t = tid; t < n_owned; t += THREADS) cp_async<8>(sBptr + t, p.cl_bptr + q0 + t);
    {
      const unsigned short* esrc = p.ent_src + cur.ent0;  // even offset: 4-byte aligned
      const int n_ent = n_inc * NNE + n_owned;
      for (int t = tid; t < (n_ent + 1) / 2; t += THREADS) cp_async<4>(sEnt + 2 * t, esrc + 2 * t);
    }
    for (int t = tid; t < n_heavy; t += THREADS) cp_async<4>(sHeavy + t, p.heavy_slot + h0 + t);
    if constexpr (MMA) {
      for (int t = tid; t < n_te; t += THREADS) cp_async<4>(sTe + t, p.te_desc + te0 + t);
    }
    for (int t = tid; t < n_owned; t += THREADS) {
      cp_async<4>(sSlotBase + t, p.cl_slot_loc + q0 + t);
      cp_async<4>(sFinc + t, p.cl_finc_loc + q0 + t);
    }
    cp_async_commit();
    if (tid == 0) {  // end markers come from the header
      sSlotBase[n_owned] = n_slots;
      sFinc[n_owned] = n_inc;
    }
  }
  cp_async_wait_group<1>();  // the phase-1 inputs of this cluster (issued one cluster ago) have landed
  __syncthreads();
  FDK_CLK(1)  // phase 0 wait

  if constexpr (!MMA) {
    // ---------------- phase 1: geometry (+ w*sigma) per (touched element, gp) ----------------
    for (int task = tid; task < n_te * NGP; task += THREADS) {
      const int le = task / NGP, g = task - le * NGP;
      const unsigned char* lc = sLconn + le * NNE;
      int ln[NNE];
      double X[NNE][DIM];
  #pragma unroll
      for (int k = 0; k < NNE; ++k) {
        ln[k] = lc[k];
  #pragma unroll
        for (int d = 0; d < DIM; ++d) X[k][d] = sX[ln[k] * DIM + d];
      }
      double G[NNE][DIM];
      const double w = gp_geometry<NNE, DIM>(sdN + g * TSTR, sW[g], X, G);
      {
        double2* out = reinterpret_cast<double2*>(sG + le * ESTR + g * GSTR);
  #pragma unroll
        for (int t = 0; t < GROW / 2; ++t) {
          const int k0 = (2 * t) / DIM, d0 = (2 * t) % DIM, k1 = (2 * t + 1) / DIM, d1 = (2 * t + 1) % DIM;
          out[t] = make_double2(G[k0][d0], G[k1][d1]);
        }
      }
      sWd[le * WSTR + g] = w;

      if (do_bts) {
        double* so = sSig + le * SSTR + g * NSIG;
        if constexpr (PHYS == PHYS_HEAT) {
          double gT[DIM], dTg = 0.0;
  #pragma unroll
          for (int d = 0; d < DIM; ++d) gT[d] = 0.0;
  #pragma unroll
          for (int k = 0; k < NNE; ++k) {
            const double T = sU[ln[k] * 2 + 0];
  #pragma unroll
            for (int d = 0; d < DIM; ++d) gT[d] = fma(T, G[k][d], gT[d]);
            dTg = fma(sN[g * NNE + k], sU[ln[k] * 2 + 1], dTg);
          }
  #pragma unroll
          for (int i = 0; i < DIM; ++i) {
            double q = 0.0;
  #pragma unroll
            for (int j = 0; j < DIM; ++j) q = fma(a.cond[i * 3 + j], gT[j], q);
            so[i] = w * q;
          }
          so[DIM] = w * a.rcdt * dTg;
        } else {
          double sig[6];
          if (a.stress_gp != nullptr) {
            const int64_t e = p.cl_te_elem[te0 + le];
            const double* sp = a.stress_gp + 6 * ((int64_t)g * p.n_elems + e);
  #pragma unroll
            for (int s = 0; s < 6; ++s) sig[s] = sp[s];
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
                const double u = sU[ln[k] * DIM + v];
  #pragma unroll
                for (int d = 0; d < DIM; ++d) gu[v][d] = fma(u, G[k][d], gu[v][d]);
              }
            double eps[6];
            voigt_strain<DIM>(gu, eps);
            if constexpr (PHYS == PHYS_ISO) {
              // sigma = lambda tr(eps) 1 + 2 mu eps  (H of fedoo/constitutivelaw/elastic_isotrop.py:57-66)
              const double tr = eps[0] + eps[1] + eps[2];
              const double lt = a.lam * tr, m2 = 2.0 * a.mu;
              sig[0] = fma(m2, eps[0], lt);
              sig[1] = fma(m2, eps[1], lt);
              sig[2] = fma(m2, eps[2], lt);
              sig[3] = a.mu * eps[3];
              sig[4] = a.mu * eps[4];
              sig[5] = a.mu * eps[5];
            } else {
              if (a.tangent_gp != nullptr) {
                const int64_t e = p.cl_te_elem[te0 + le];
                apply_tangent(a.tangent_gp + 36 * ((int64_t)g * p.n_elems + e), 1, 6, eps, sig);
              } else {
                apply_tangent(a.C, 6, 1, eps, sig);
              }
            }
          }
          if constexpr (DIM == 3) {
  #pragma unroll
            for (int s = 0; s < 6; ++s) so[s] = w * sig[s];
          } else {
            so[0] = w * sig[0];
            so[1] = w * sig[1];
            so[2] = w * sig[3];
          }
        }
      }
    }
    if (has_next) load_node_ids(nxt);  // consumed after phase 2
    cp_async_wait_group<0>();          // this cluster's descriptors
    __syncthreads();
    FDK_CLK(2)  // phase 1

    // ---------------- phase 2: per-incidence half rows in registers ----------------
    double acc[NH][BLK];
    double f[NV];
  #pragma unroll
    for (int j = 0; j < NH; ++j)
  #pragma unroll
      for (int b = 0; b < BLK; ++b) acc[j][b] = 0.0;
  #pragma unroll
    for (int v = 0; v < NV; ++v) f[v] = 0.0;
    const bool do_f = do_bts && half == 0;  // the nodal force is accumulated by the first half thread only

    if (it < n_inc) {
      const int le = my_desc & 0xFFF, i = my_desc >> 12;
      const double* eb = sG + le * ESTR;
      [[maybe_unused]] int64_t e_glob = 0;
      if constexpr (PHYS == PHYS_GENERAL) {
        if (a.tangent_gp != nullptr) e_glob = p.cl_te_elem[te0 + le];
      }
  #pragma unroll 1
      for (int g = 0; g < NGP; ++g) {
        const double* gb = eb + g * GSTR;
        // dN/dx of this thread's column nodes: the lanes of one element read the same addresses
        double Gr[NH * DIM];
        {
          if constexpr ((NH * DIM) % 2 == 0) {
            const double2* g2 = reinterpret_cast<const double2*>(gb + j0 * DIM);
  #pragma unroll
            for (int t = 0; t < NH * DIM / 2; ++t) {
              const double2 v = g2[t];
              Gr[2 * t] = v.x;
              Gr[2 * t + 1] = v.y;
            }
          } else {
  #pragma unroll
            for (int t = 0; t < NH * DIM; ++t) Gr[t] = gb[j0 * DIM + t];
          }
        }
        const double w = sWd[le * WSTR + g];
        double gi[DIM];
  #pragma unroll
        for (int d = 0; d < DIM; ++d) gi[d] = gb[i * DIM + d];

        if (do_f) {
          const double* ws = sSig + le * SSTR + g * NSIG;
          if constexpr (PHYS == PHYS_HEAT) {
            double s = ws[DIM] * sN[g * NNE + i];
  #pragma unroll
            for (int d = 0; d < DIM; ++d) s = fma(ws[d], gi[d], s);
            f[0] += s;
          } else if constexpr (DIM == 3) {
            f[0] += ws[0] * gi[0] + ws[3] * gi[1] + ws[4] * gi[2];
            f[1] += ws[1] * gi[1] + ws[3] * gi[0] + ws[5] * gi[2];
            f[2] += ws[2] * gi[2] + ws[4] * gi[0] + ws[5] * gi[1];
          } else {
            f[0] += ws[0] * gi[0] + ws[2] * gi[1];
            f[1] += ws[1] * gi[1] + ws[2] * gi[0];
          }
        }

        if (do_mat) {
          if constexpr (PHYS == PHYS_ISO) {
            double wgi[DIM];
  #pragma unroll
            for (int d = 0; d < DIM; ++d) wgi[d] = w * gi[d];
  #pragma unroll
            for (int j = 0; j < NH; ++j) {
  #pragma unroll
              for (int cc = 0; cc < DIM; ++cc)
  #pragma unroll
                for (int aa = 0; aa < DIM; ++aa)
                  acc[j][cc * DIM + aa] = fma(wgi[cc], Gr[j * DIM + aa], acc[j][cc * DIM + aa]);
            }
          } else if constexpr (PHYS == PHYS_HEAT) {
            double kgi[DIM];
  #pragma unroll
            for (int d = 0; d < DIM; ++d) {
              double s = 0.0;
  #pragma unroll
              for (int d2 = 0; d2 < DIM; ++d2) s = fma(gi[d2], a.cond[d2 * 3 + d], s);
              kgi[d] = w * s;
            }
            // lumped capacity: row sum of the consistent mass goes to the diagonal
            double nsum = 0.0;
  #pragma unroll
            for (int j = 0; j < NNE; ++j) nsum += sN[g * NNE + j];
            const double m = a.rcdt * w * sN[g * NNE + i] * nsum;
  #pragma unroll
            for (int j = 0; j < NH; ++j) {
              double s = (j0 + j == i) ? m : 0.0;
  #pragma unroll
              for (int d = 0; d < DIM; ++d) s = fma(kgi[d], Gr[j * DIM + d], s);
              acc[j][0] += s;
            }
          } else {  // PHYS_GENERAL: t[c][s] = w sum_s' B_I[s'][c] C[s'][s]; acc[j][c][a] += sum_s t[c][s] B_J[s][a]
            const double* Cg;
            int si, sj;
            if (a.tangent_gp != nullptr) {
              Cg = a.tangent_gp + 36 * ((int64_t)g * p.n_elems + e_glob);
              si = 1;
              sj = 6;
            } else {
              Cg = a.C;
              si = 6;
              sj = 1;
            }
            double t[DIM][6];
  #pragma unroll
            for (int s = 0; s < 6; ++s) {
              if constexpr (DIM == 3) {
                const double c0 = Cg[0 * si + s * sj], c1 = Cg[1 * si + s * sj], c2 = Cg[2 * si + s * sj];
                const double c3 = Cg[3 * si + s * sj], c4 = Cg[4 * si + s * sj], c5 = Cg[5 * si + s * sj];
                t[0][s] = w * (gi[0] * c0 + gi[1] * c3 + gi[2] * c4);
                t[1][s] = w * (gi[1] * c1 + gi[0] * c3 + gi[2] * c5);
                t[2][s] = w * (gi[2] * c2 + gi[0] * c4 + gi[1] * c5);
              } else {
                const double c0 = Cg[0 * si + s * sj], c1 = Cg[1 * si + s * sj], c3 = Cg[3 * si + s * sj];
                t[0][s] = w * (gi[0] * c0 + gi[1] * c3);
                t[1][s] = w * (gi[1] * c1 + gi[0] * c3);
              }
            }
  #pragma unroll
            for (int j = 0; j < NH; ++j) {
              double gj[DIM];
  #pragma unroll
              for (int d = 0; d < DIM; ++d) gj[d] = Gr[j * DIM + d];
  #pragma unroll
              for (int cc = 0; cc < DIM; ++cc) {
                if constexpr (DIM == 3) {
                  acc[j][cc * 3 + 0] += t[cc][0] * gj[0] + t[cc][3] * gj[1] + t[cc][4] * gj[2];
                  acc[j][cc * 3 + 1] += t[cc][1] * gj[1] + t[cc][3] * gj[0] + t[cc][5] * gj[2];
                  acc[j][cc * 3 + 2] += t[cc][2] * gj[2] + t[cc][4] * gj[0] + t[cc][5] * gj[1];
                } else {
                  acc[j][cc * 2 + 0] += t[cc][0] * gj[0] + t[cc][3] * gj[1];
                  acc[j][cc * 2 + 1] += t[cc][1] * gj[1] + t[cc][3] * gj[0];
                }
              }
            }
          }
        }
      }
    }
    __syncthreads();  // everyone is done reading the geometry region; it becomes the staging region
    FDK_CLK(3)  // phase 2
    if (it < n_inc) {
      if (do_mat) {
  #pragma unroll
        for (int j = 0; j < NH; ++j) {
          double* sp = sBlk + it * ISTR + (j0 + j) * BLK;
  #pragma unroll
          for (int b = 0; b < BLK; ++b) sp[b] = acc[j][b];
        }
      }
      if (do_f) {
  #pragma unroll
        for (int v = 0; v < NV; ++v) sF[my_fdst * NV + v] = f[v];
      }
    }
  } else {
    // ---------------- phase 1m: inverse Jacobian and w per (touched element, gp) ----------------
    static_assert(THREADS % NGP == 0, "the Gauss point of a thread's tasks is fixed");
    double dNr[DIM * NNE];  // reference gradients at this thread's Gauss point
#pragma unroll
    for (int t = 0; t < DIM * NNE; ++t) dNr[t] = sdN[(tid % NGP) * TSTR + t];
    for (int task = tid; task < n_te * NGP; task += THREADS) {
      const int le = task / NGP, g = task - le * NGP;
      const unsigned char* lc = sLconn + le * NNE;
      const double* dN = dNr;
      double J[DIM][DIM];
#pragma unroll
      for (int r = 0; r < DIM; ++r)
#pragma unroll
        for (int x = 0; x < DIM; ++x) J[r][x] = 0.0;
#pragma unroll
      for (int k = 0; k < NNE; ++k) {
        const double* xk = sX + (int)lc[k] * DIM;
        const double x0 = xk[0], x1 = xk[1], x2 = xk[2];
#pragma unroll
        for (int r = 0; r < DIM; ++r) {
          const double dn = dN[r * NNE + k];
          J[r][0] = fma(dn, x0, J[r][0]);
          J[r][1] = fma(dn, x1, J[r][1]);
          J[r][2] = fma(dn, x2, J[r][2]);
        }
      }
      double iJ[DIM][DIM];
      const double det = invert<DIM>(J, iJ);
      double2* out = reinterpret_cast<double2*>(sJ + le * L::JSTR + g * 10);
      out[0] = make_double2(iJ[0][0], iJ[0][1]);
      out[1] = make_double2(iJ[0][2], iJ[1][0]);
      out[2] = make_double2(iJ[1][1], iJ[1][2]);
      out[3] = make_double2(iJ[2][0], iJ[2][1]);
      out[4] = make_double2(iJ[2][2], sW[g] * fabs(det));
    }
    if (has_next) load_node_ids(nxt);  // consumed after phase 2m
    cp_async_wait_group<0>();          // this cluster's descriptors
    __syncthreads();
    FDK_CLK(2)  // phase 1m

    // ---------------- phase 2m: one warp per touched element, S^e by DMMA ----------------
    {
      const int lane = tid & 31, warp = tid >> 5;
      const int r = lane >> 2, q = lane & 3;  // node (MMA row / column) and Gauss point pair (MMA k index)
      // reference gradients of node r at the two Gauss points of this lane (constant for the kernel)
      double dn0[DIM], dn1[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        dn0[d] = sdN[q * TSTR + d * NNE + r];
        dn1[d] = sdN[(q + 4) * TSTR + d * NNE + r];
      }
      for (int le = warp; le < n_te; le += THREADS / 32) {
        const double2* j0p = reinterpret_cast<const double2*>(sJ + le * L::JSTR + q * 10);
        const double2* j1p = j0p + 20;  // Gauss point q + 4
        double G0[DIM], G1[DIM], w0, w1;
        {
          const double2 a0 = j0p[0], a1 = j0p[1], a2 = j0p[2], a3 = j0p[3], a4 = j0p[4];
          G0[0] = fma(a0.x, dn0[0], fma(a0.y, dn0[1], a1.x * dn0[2]));
          G0[1] = fma(a1.y, dn0[0], fma(a2.x, dn0[1], a2.y * dn0[2]));
          G0[2] = fma(a3.x, dn0[0], fma(a3.y, dn0[1], a4.x * dn0[2]));
          w0 = a4.y;
          const double2 b0 = j1p[0], b1 = j1p[1], b2 = j1p[2], b3 = j1p[3], b4 = j1p[4];
          G1[0] = fma(b0.x, dn1[0], fma(b0.y, dn1[1], b1.x * dn1[2]));
          G1[1] = fma(b1.y, dn1[0], fma(b2.x, dn1[1], b2.y * dn1[2]));
          G1[2] = fma(b3.x, dn1[0], fma(b3.y, dn1[1], b4.x * dn1[2]));
          w1 = b4.y;
        }
        double A0[DIM], A1[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          A0[d] = w0 * G0[d];
          A1[d] = w1 * G1[d];
        }
        // nine 8x8 tiles: d[c*3+a][jj] = S^e_{I=r, J=2q+jj}[c][a] (the lower tiles are the transposes of the
        // upper ones; computing them keeps every 3x3 block whole in one lane, so the stores below are
        // contiguous and bank-conflict free -- the FP64 pipe has the headroom, shared memory does not)
        double d[9][2];
#pragma unroll
        for (int cc = 0; cc < DIM; ++cc)
#pragma unroll
          for (int aa = 0; aa < DIM; ++aa) {
            d[cc * 3 + aa][0] = 0.0;
            d[cc * 3 + aa][1] = 0.0;
            dmma884(d[cc * 3 + aa][0], d[cc * 3 + aa][1], A0[cc], G0[aa]);
            dmma884(d[cc * 3 + aa][0], d[cc * 3 + aa][1], A1[cc], G1[aa]);
          }
        // rows of owned nodes go to the incidence-major staging (the element's incidences are consecutive)
        const unsigned ted = sTe[le];
        const unsigned mask = ted >> 16;
        if ((mask >> r) & 1u) {
          double* sp = sBlk + ((int)(ted & 0xFFFF) + __popc(mask & ((1u << r) - 1u))) * ISTR + (2 * q) * BLK;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int b = 0; b < BLK; ++b) sp[jj * BLK + b] = d[b][jj];
        }
      }
    }
  }
  if (has_next) fetch_inputs(nxt, buf ^ 1);  // lands during phase 3 and the next cluster's phase 0
  else cp_async_commit();
  __syncthreads();
  FDK_CLK(4)  // staging stores / next fetch issue

  // slot record: first staging entry | local touched-node index of the column node << 16 | owner << 24;
  // the run of a slot ends where the next one starts (minus the gap entry that closes a block row)
  if (do_mat) {
    // ---------------- phase 3a: lane-balanced pre-reduction of the heavy slots ----------------
    if (n_heavy > 0) {  // uniform over the CTA
      for (int t = tid; t < n_heavy * BLK; t += THREADS) {
        const int h = t / BLK, b = t - h * BLK;
        const int s = sHeavy[h];
        const unsigned r0 = sRec[s], r1 = sRec[s + 1];
        const int e0 = r0 & 0xFFFF;
        const int e1 = (int)(r1 & 0xFFFF) - (((r0 ^ r1) >> 24) ? 1 : 0);
        // four independent chains: the lookups and loads of four contributions are in flight together (a vertex node
        // of a tet mesh gathers 20-40 contributions in its diagonal slot; one dependent chain was latency-bound)
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        int e = e0;
        for (; e + 3 < e1; e += 4) {
          const int s0 = sEnt[e], s1 = sEnt[e + 1], s2 = sEnt[e + 2], s3 = sEnt[e + 3];
          v0 += sBlk[(s0 / NNE) * ISTR + (s0 % NNE) * BLK + b];
          v1 += sBlk[(s1 / NNE) * ISTR + (s1 % NNE) * BLK + b];
          v2 += sBlk[(s2 / NNE) * ISTR + (s2 % NNE) * BLK + b];
          v3 += sBlk[(s3 / NNE) * ISTR + (s3 % NNE) * BLK + b];
        }
        for (; e < e1; ++e) {
          const int src = sEnt[e];
          v0 += sBlk[(src / NNE) * ISTR + (src % NNE) * BLK + b];
        }
        const double v = (v0 + v1) + (v2 + v3);
        const int src0 = sEnt[e0];
        sBlk[(src0 / NNE) * ISTR + (src0 % NNE) * BLK + b] = v;
      }
      __syncthreads();
      FDK_CLK(5)  // phase 3a (heavy)
    }
    // ---------------- phase 3b: slot gather, constitutive closed form, final stores ----------------
    for (int s = tid; s < n_slots; s += THREADS) {
      const unsigned r0 = sRec[s], r1 = sRec[s + 1];
      const int e0 = r0 & 0xFFFF;
      const int n = r0 >> 24;
      int cnt = (int)(r1 & 0xFFFF) - e0 - (((r0 ^ r1) >> 24) ? 1 : 0);  // one gap entry after each row
      if (cnt > HEAVY_T) cnt = 1;                                       // pre-reduced in 3a
      const int sb = sSlotBase[n];
      const int deg = sSlotBase[n + 1] - sb;
      const int pcol = s - sb;
      const int64_t bp = sBptr[n];
      // at most HEAVY_T contributions: all source lookups, then all block loads, are issued together
      // (two shared-memory latencies per slot instead of two per contribution)
      const double* bp_[HEAVY_T];
#pragma unroll
      for (int t = 0; t < HEAVY_T; ++t) {
        const int src = sEnt[e0 + (t < cnt ? t : 0)];
        bp_[t] = sBlk + (src / NNE) * ISTR + (src % NNE) * BLK;
      }
      double S[BLK];
#pragma unroll
      for (int b = 0; b < BLK; ++b) S[b] = 0.0;
#pragma unroll
      for (int t = 0; t < HEAVY_T; ++t) {
        if (t < cnt) {
#pragma unroll
          for (int b = 0; b < BLK; ++b) S[b] += bp_[t][b];
        }
      }
      double Kb[BLK];
      if constexpr (PHYS == PHYS_ISO) {
        // K_IJ = lambda S + mu S^T + mu tr(S) 1
        double tr = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) tr += S[d * DIM + d];
#pragma unroll
        for (int cc = 0; cc < DIM; ++cc)
#pragma unroll
          for (int aa = 0; aa < DIM; ++aa) {
            double v = fma(a.lam, S[cc * DIM + aa], a.mu * S[aa * DIM + cc]);
            if (cc == aa) v = fma(a.mu, tr, v);
            Kb[cc * DIM + aa] = v;
          }
      } else {
#pragma unroll
        for (int b = 0; b < BLK; ++b) Kb[b] = S[b];
      }
#pragma unroll
      for (int cc = 0; cc < NV; ++cc) {
        double* row = a.K + ((int64_t)cc * NV * p.blk_nnz + (int64_t)NV * bp);
#pragma unroll
        for (int aa = 0; aa < NV; ++aa) __stcs(row + (int64_t)aa * deg + pcol, Kb[cc * NV + aa]);
      }
      if constexpr (PHYS != PHYS_HEAT) {
        if (fuse_ku) {  // this slot's share of (K U)_I: the assembled block times the dofs of its column node
          const double* uj = sU + (int)((r0 >> 16) & 0xFF) * DIM;
#pragma unroll
          for (int cc = 0; cc < NV; ++cc) {
            double r = 0.0;
#pragma unroll
            for (int aa = 0; aa < NV; ++aa) r = fma(Kb[cc * NV + aa], uj[aa], r);
            sR[s * NV + cc] = r;
          }
        }
      }
    }
  }
  if (fuse_ku) __syncthreads();  // sR complete (uniform)
  FDK_CLK(6)  // phase 3b
  if (do_vec) {
    // one group of 8 lanes per owned node: the lanes stride over the node's per-slot products (or per-incidence nodal
    // forces) and a three-step shuffle closes the sum -- a fixed order, so D stays bit-reproducible.  (One thread per
    // node was a chain of up to ~90 dependent shared-memory loads for a tet10 vertex node.)
    const int sub = tid & 7;
    for (int n0 = (tid >> 3); n0 < ((n_owned + 3) & ~3); n0 += THREADS / 8) {  // whole warps iterate together
      const bool live = n0 < n_owned;
      const int n = live ? n0 : 0;
      double s[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) s[v] = 0.0;
      const int k0 = fuse_ku ? sSlotBase[n] : sFinc[n], k1 = live ? (fuse_ku ? sSlotBase[n + 1] : sFinc[n + 1]) : k0;
      const double* src = fuse_ku ? sR : sF;
      for (int k = k0 + sub; k < k1; k += 8)
#pragma unroll
        for (int v = 0; v < NV; ++v) s[v] += src[k * NV + v];
#pragma unroll
      for (int o = 4; o > 0; o >>= 1)
#pragma unroll
        for (int v = 0; v < NV; ++v) s[v] += __shfl_xor_sync(0xffffffffu, s[v], o);
      if (live && sub == 0) {
        const int node = p.cl_node[q0 + n];
#pragma unroll
        for (int v = 0; v < NV; ++v) a.D[(int64_t)v * p.n_nodes + node] = -s[v];
      }
    }
  }
  __syncthreads();  // staging, descriptors and row buffers are free for the next cluster
  FDK_CLK(7)  // D tail + end barrier
  cur = nxt;
  }  // cluster loop
  FDK_CLK_FLUSH
}

// ---- host launcher -----------------------------------------------------------------------
template <class El, int PHYS, int THREADS, int MINB, bool MMA>
int launch_assemble_t(AsmArgs& a, cudaStream_t stream) {
  using L = Layout<El, PHYS>;
  const fdk_plan& p = a.p;
  FDK_REQUIRE(2 * p.cap_inc <= THREADS, FDK_ECAP, "cluster with %d incidences exceeds half the CTA size %d",
              p.cap_inc, THREADS);
  FDK_REQUIRE(p.cap_te < 4096 && p.cap_tn <= 256 && p.cap_owned < 255 && p.cap_ent < 65536 && p.cap_slots < 65535,
              FDK_ECAP, "cluster capacity overflow (te=%d tn=%d owned=%d ent=%d slots=%d)", p.cap_te, p.cap_tn,
              p.cap_owned, p.cap_ent, p.cap_slots);
  FDK_REQUIRE(p.nvar == L::NV, FDK_EINVAL, "plan nvar %d does not match the operator (%d)", p.nvar, L::NV);
  const bool bts = (a.compute & FDK_VECTOR) && !a.fuse_ku;
  const size_t smem = L::smem_bytes(p, bts, MMA, &a.big_doubles);
  FDK_REQUIRE(smem <= 227 * 1024, FDK_ECAP, "cluster needs %zu bytes of shared memory (> 227 KB)", smem);
  if (p.n_clusters == 0) return 0;
  if (int rc = ensure_device_tables()) return rc;
  auto kern = k_assemble<El, PHYS, THREADS, MINB, MMA>;
  static thread_local size_t smem_set = 0;  // per instantiation
  if (smem > smem_set) {
    FDK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  // persistent CTAs: as many as are resident at once, each walking the cluster list with stride gridDim
  static thread_local int resident = 0;
  if (resident == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    FDK_CUDA(cudaGetDevice(&dev));
    FDK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FDK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    resident = sms * (per_sm > 0 ? per_sm : 1);
  }
  const int grid = p.n_clusters < resident ? p.n_clusters : resident;
  kern<<<grid, THREADS, smem, stream>>>(a);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

template <class El, int PHYS>
int launch_assemble(AsmArgs& a, cudaStream_t stream) {
  // hex8 + isotropic law + matrix requested + no B^T sigma pass: FP64 tensor-core producer
  if constexpr (El::ID == FDK_HEX8 && PHYS == PHYS_ISO) {
    const bool bts = (a.compute & FDK_VECTOR) && !a.fuse_ku;
    if ((a.compute & FDK_MATRIX) && !bts && !a.no_mma) {
      if (a.p.threads == El::THREADS) return launch_assemble_t<El, PHYS, El::THREADS, 1, true>(a, stream);
      if (a.p.threads == El::THREADS / 2) return launch_assemble_t<El, PHYS, El::THREADS / 2, 2, true>(a, stream);
    }
  }
  // the plan states the CTA size it was built for (fedoo_b200/plan.py); small CTAs run 2 per SM
  if (a.p.threads == El::THREADS) return launch_assemble_t<El, PHYS, El::THREADS, 1, false>(a, stream);
  if (a.p.threads == El::THREADS / 2) return launch_assemble_t<El, PHYS, El::THREADS / 2, 2, false>(a, stream);
  set_error("plan built for %d threads per cluster; this element supports %d or %d", a.p.threads, El::THREADS,
            El::THREADS / 2);
  return FDK_EINVAL;
}

template <int PHYS>
int dispatch_assemble(AsmArgs& a, cudaStream_t stream) {
  switch (a.p.elem_type) {
    case FDK_HEX8: return launch_assemble<Hex8, PHYS>(a, stream);
    case FDK_TET4: return launch_assemble<Tet4, PHYS>(a, stream);
    case FDK_TET10: return launch_assemble<Tet10, PHYS>(a, stream);
    case FDK_QUAD4: return launch_assemble<Quad4, PHYS>(a, stream);
  }
  set_error("unknown element type %d", a.p.elem_type);
  return FDK_EINVAL;
}

}  // namespace fdk
