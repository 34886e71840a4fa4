// fdk_assemble_iso.cuh -- balanced cluster kernel for the headline case: isotropic elasticity, matrix
// requested, residual (if any) taken from the assembled rows (D = -K_row . U).
//
// Same owner-computes algorithm and the same plan as k_assemble (fdk_assemble.cuh); what differs is how the
// work is laid on the SM.  Measured on B200 (profiles/r1_pipe_overlap.txt, r1_phase_clocks.txt): FP64 FMAs and
// shared-memory instructions largely SERIALISE (a DFMA warp-instruction costs 0.5 SM-cycles, an LDS ~1, an
// STS.64 2, and a mix costs ~0.85 x the sum), and one third of the stall samples of k_assemble were warps
// waiting at barriers because every phase ran 1.2-1.7 "rounds" over 512 threads.  Hence:
//   * THREADS = TPI x (max incidences): TPI = 4 threads per incidence (two column blocks each, 18
//     accumulators, <= 64 registers) -> 1024 threads, 32 warps per SM; for a 32-node hex8 cluster the
//     geometry phase (600 tasks), the block phase (256 incidences x 4) and the gather (864 slots) are ONE
//     round each, so every warp carries the same load between two barriers;
//   * sqrt(w_g |det J|) is folded into dN/dx in the geometry phase (S_IJ = sum_g (sqrt(w) G_I) (x) (sqrt(w) G_J)
//     exactly as before up to one rounding): no weight load and no weight multiply in the block phase;
//   * the geometry task streams over the element nodes twice (Jacobian, then dN/dx) instead of holding
//     the coordinates and the 24 derivatives in registers;
//   * four barriers per cluster: the prefetched inputs of the next cluster are waited for BEFORE the
//     barrier that closes the gather, so that barrier also publishes them; the per-node reduction of the
//     residual is spread over 8 lanes per (node, component) and overlaps the next cluster's geometry phase;
//   * heavy slots (the diagonal) are pre-reduced with all source lookups and all block loads in flight
//     together (two shared-memory latencies instead of two per contribution).
//
// PHYS = PHYS_GENERAL (3D): the same organisation for a general tangent -- uniform 6x6 or per Gauss point (6,6,N), the
// consistent tangent of the J2 update -- with the residual either fused (uniform tangent, D = -K_row . U) or
// integrated as B^T sigma from a given stress (the Newton iteration of a plastic step).  Per (touched element,
// Gauss point) the geometry phase also parks the 36 tangent entries (cp.async, issued one cluster ahead) and
// sqrt(w) sigma in shared memory, so the block phase reads them as broadcasts instead of 36 global loads per
// thread and Gauss point (k_assemble: 63 % of its time, with 650 bytes of register spills).  Sized for 16-node
// clusters (plans built with small = True): 512 threads, 4 per incidence, <= 128 registers.
//
// Reference: same as k_assemble -- fedoo/core/assembly.py:143-470, 776-928; fedoo/core/_sparsematrix.py:55-174,
// 256-315; fedoo/constitutivelaw/elastic_isotrop.py:35-68.
#pragma once
#include "fdk_assemble.cuh"

#ifndef FDK_P2_UNROLL
#define FDK_P2_UNROLL 2  // Gauss points in flight per thread in the block phase (register budget: 64)
#endif

#ifndef FDK_HEX_REFLECT
// 1: hex8 geometry phase without table loads (see HexRef below).  MEASURED on B200 (round 2, 8 M elements): 17.8 ms with
// the permuted coordinate loads straight from sX (the eight Gauss-point lanes of an element then read eight different
// nodes: +57 M bank-conflict wavefronts per million elements, exactly what the 49 M table wavefronts saved), 18.7 ms
// with the element-local coordinate copy that removes those conflicts (one more dependent STS -> LDS round per task),
// against 17.7 ms for the table version.  Kept as a compile-time option, off.
#define FDK_HEX_REFLECT 0
#endif

#ifndef FDK_HEX_MONO
// 1: hex8 geometry phase in the monomial basis of the trilinear element (see the kernel, phase 1): twelve lanes of a warp
// compute the coefficient vectors of the warp's four elements, every task then forms J in 27 FMA and builds the
// reference gradients from the Gauss point's signs -- 124 instead of 194 shared-memory wavefronts per warp-task and no
// table.  MEASURED on B200 (round 2, 8 M elements): 20.5 ms against 17.65 ms for the table version (J2 plate: 22.5
// against 19.8 ms).  A warp runs ONE geometry task per cluster, so the phase follows the length of its dependency chain,
// not the wavefront count: coordinate loads -> 8-term sums on 12 of 32 lanes -> store -> warp sync -> loads -> J is
// longer than loads -> J.  (The same basis is a clear win where one thread integrates a whole element: the
// residual-only kernel k_elem_force_hex8, 3.37 -> 2.49 ms.)  Kept as a compile-time option, off.
#define FDK_HEX_MONO 0
#endif

namespace fdk {

// hex8: the reference gradients at Gauss point g are those at Gauss point 0 of the element REFLECTED along the axes
// where g sits on the + side (dN_k/dxi_r (g) = s_r dN_pi(k)/dxi_r (g0), s_r = -1 on a reflected axis; the signs cancel
// in G = J^-1 dN).  A geometry task therefore runs the Gauss-point-0 formula -- 24 COMPILE-TIME constants as DFMA
// operands instead of 24 + 24 table loads from shared memory (96 of the ~200 shared-memory wavefronts of a task) -- on a
// node-permuted element, and stores position j = gradient of node pi_g(j).  Node k sits at position pi_g(k) (an
// involution): the block phase reads it there.  In node codes c = b0 + 2 b1 + 4 b2 (sign bits of the reference
// coordinates) the reflection is an XOR with the Gauss point's bits; code <-> local node: {0,1,3,2,4,5,7,6}.
struct HexRef {
  static constexpr unsigned CODES = 0x67542310u;  // nibble k: code of local node k, and local node of code k
  __host__ __device__ static constexpr int bit(int k, int d) {
    return d == 0 ? ((k ^ (k >> 1)) & 1) : (d == 1 ? ((k >> 1) & 1) : ((k >> 2) & 1));
  }
  // 1 + n_d x_d at Gauss point 0 (x = -a on every axis), as the table computes it (csrc/fdk_tables.cuh)
  __host__ __device__ static constexpr double f(int k, int d) {
    return 1.0 + (bit(k, d) ? 1.0 : -1.0) * (-1.0 * 0.5773502691896258);
  }
  __host__ __device__ static constexpr double dn(int r, int k) {  // dN_k / dxi_r at Gauss point 0
    return 0.125 * (bit(k, r) ? 1.0 : -1.0) * f(k, r == 0 ? 1 : 0) * f(k, r == 2 ? 1 : 2);
  }
  // Gauss point g = ix * 4 + iy * 2 + iz (xi slowest): bits to flip in a node code
  __host__ __device__ static constexpr int gx(int g) { return (g >> 2) | (g & 2) | ((g & 1) << 2); }
  __device__ static __forceinline__ int perm(int code, int gxv) { return (int)((CODES >> (4 * (code ^ gxv))) & 7u); }
};

constexpr int P2_UNROLL = FDK_P2_UNROLL;
#ifndef FDK_ISO_PART_UNIFORM
#define FDK_ISO_PART_UNIFORM 0
#endif
// 1: the threads of a warp work on 32 incidences and the same column part (columns part*NH ..., 128-bit column
// loads shared by the lanes of an element); 0: the TPI threads of an incidence sit in adjacent lanes (columns
// part, part + TPI, ...: own-row loads shared by TPI lanes)
constexpr bool PART_UNIFORM = FDK_ISO_PART_UNIFORM != 0;

template <class El, int TPI>
struct IsoLayout {
  using L = Layout<El, PHYS_ISO>;
  static constexpr int CSTR = 36;  // tangent entries per (touched element, Gauss point), Fortran order i + 6 j
  static constexpr int SGS = 6;    // sqrt(w) sigma per (touched element, Gauss point)
  static constexpr int NNE = El::NNE, NGP = El::NGP, DIM = El::DIM, NV = DIM, BLK = NV * NV;
  static constexpr int GROW = L::GROW, GSTR = L::GSTR;
  static constexpr int NH = NNE / TPI;  // column blocks per thread: j = part, part + TPI, ...
  static_assert(NNE % TPI == 0, "threads per incidence must divide the element nodes");
  // Bank-conflict-free strides (8-byte banks, 16 per half-warp; lanes = 8 incidences x TPI parts):
  //  * staging: lane (it, part) stores block j = part (+ TPI) at it*ISTR + j*BLK + b.  With BLK = 9 and TPI = 4 the
  //    part offsets are {0,9,2,11} mod 16, so ISTR = 4 or 12 mod 16 makes the 16 lanes of a half-warp hit 16
  //    different banks (hex8: 76);
  //  * geometry: the own-row loads of the incidences of elements le and le+1 differ by ESTR + 3 (i - i'); with
  //    ESTR = 8 mod 16 they never collide (3 d = 8 mod 16 has no solution with |d| <= 7) (hex8: 216).
  static constexpr int ISTR = (NNE == 8 && DIM == 3 && TPI == 4 && !PART_UNIFORM) ? 76 : L::ISTR;
  // hex8, 4 threads per incidence: block positions come from the plan (csrc/fdk_color.cuh): the 16 blocks a half-warp
  // stores together share a window of 16 positions, permuted so that the gather is conflict-free as well
  static constexpr bool COLORED = NNE == 8 && TPI == 4 && !PART_UNIFORM;
  // hex8: reflected geometry phase (HexRef) -- Gauss-point-0 constants on a node-permuted element.  The lanes of one
  // element then read eight DIFFERENT nodes, so the coordinates are first copied into an element-local array
  // [le][node][3] (stride 24 doubles: the 16 lanes of a half-warp = 2 elements x 8 Gauss points hit 16 distinct 8-byte banks)
  static constexpr bool HEXREF = El::ID == FDK_HEX8 && COLORED && ISO_COLS_ADJ && (FDK_HEX_REFLECT != 0);
  static constexpr int XESTR = NNE * DIM;
  // hex8: monomial-basis geometry (phase 1 of the kernel): 7 coefficient vectors per touched element share the array
  static constexpr bool MONO = El::ID == FDK_HEX8 && !HEXREF && (FDK_HEX_MONO != 0);
  __host__ __device__ static long xe_doubles(const fdk_plan& p) { return (HEXREF || MONO) ? (long)p.cap_te * XESTR : 0; }
  __host__ __device__ static long stage_doubles(const fdk_plan& p) {
    return COLORED ? (long)((p.cap_inc + 3) / 4) * 32 * BLK : (long)p.cap_inc * ISTR;
  }
  // Gauss points per geometry CHUNK: an element's sqrt(w) dN/dx takes NGP * GSTR doubles of shared memory (tet10: 15 x 34 =
  // 4 KB), which caps the touched elements of a cluster -- tet10 clusters then hold one vertex node and its mid-edge
  // neighbours (8 nodes, 60 incidences) and most threads idle.  With GCH < NGP the geometry and block phases run NGP / GCH
  // times over GCH Gauss points each (the accumulators stay in registers across the chunks), so a cluster can touch
  // NGP / GCH times more elements for the same shared memory.
  static constexpr int GCH = (El::ID == FDK_TET10) ? 5 : NGP;
  static constexpr int NCH = NGP / GCH;
  static_assert(NGP % GCH == 0, "chunks cover the Gauss points");
  static constexpr int ESTR_C = ((GCH * GSTR) / 2) % 2 ? GCH * GSTR : GCH * GSTR + 2;
  static constexpr int ESTR = (NNE == 8 && DIM == 3) ? 216 : (NCH == 1 ? L::ESTR : ESTR_C);
  // even (the reference gradients are read as 128-bit node pairs) and an ODD number of 16-byte units, so that the rows
  // of the Gauss points the lanes of a warp work on start in different bank groups (hex8: 26, tet10: 34 -- 32 would put
  // every Gauss point's row on the same banks)
  static constexpr int TSTR = (((GROW + 2) & ~1) / 2) % 2 ? ((GROW + 2) & ~1) : ((GROW + 2) & ~1) + 2;
  static_assert(NNE % 2 == 0, "node pairs");
  static constexpr int XSTR = 4;  // padded coordinates of a touched node: one 128-bit + one 64-bit load
  static constexpr int TAB_DOUBLES = (NGP * TSTR + NGP + 1) & ~1;
  static_assert(ISTR >= NNE * BLK && ESTR >= GCH * GSTR, "strides cover the rows");

  __host__ __device__ static int xu_doubles(const fdk_plan& p) { return (p.cap_tn * (XSTR + NV) + 1) & ~1; }
  // staging of the blocks; the per-slot K.u products live BEHIND both views of the big region so that the
  // residual reduction of cluster c may overlap the geometry phase of cluster c+1
  __host__ __device__ static long sr_offset(const fdk_plan& p) {
    const long g = (long)p.cap_te * ESTR, s = stage_doubles(p);
    return ((g > s ? g : s) + 1) & ~1L;
  }
  // per-slot K.u products or per-incidence nodal forces (never both)
  __host__ __device__ static long sr_doubles(const fdk_plan& p) {
    const long n = p.cap_slots > p.cap_inc ? p.cap_slots : p.cap_inc;
    return (n * NV + 1) & ~1L;
  }
  // sqrt(w) sigma is dead once the block phase is over and the nodal forces are parked after it: they share one
  // region (the price: a closing barrier per cluster on the B^T sigma path, see the kernel)
  __host__ __device__ static long rs_doubles(const fdk_plan& p, bool bts) {
    const long r = sr_doubles(p), g = bts ? (((long)p.cap_te * NGP * SGS + 1) & ~1L) : 0;
    return r > g ? r : g;
  }
  __host__ __device__ static long extra_doubles(const fdk_plan& p, bool tangent_gp, int cstr = CSTR) {
    return tangent_gp ? (long)p.cap_te * NGP * cstr : 0;
  }
  static size_t smem_bytes(const fdk_plan& p, bool tangent_gp = false, bool bts = false, int cstr = CSTR) {
    long doubles = TAB_DOUBLES + 2L * xu_doubles(p) + sr_offset(p) + rs_doubles(p, bts) + extra_doubles(p, tangent_gp, cstr) +
                   xe_doubles(p) + p.cap_owned;
    size_t bytes = (size_t)doubles * 8;
    bytes += (size_t)(2 * (p.cap_owned + 1)) * 4;      // sSlotBase, sFinc
    bytes += (size_t)(p.cap_slots + 1) * 4;            // sRec
    bytes += (size_t)p.cap_heavy * 4;                  // sHeavy
    bytes += 2 * (size_t)((p.cap_te * NNE + 3) & ~3);  // sLconn (double-buffered)
    bytes += (size_t)((p.cap_ent + 3) & ~3) * 2;       // sEnt
    return bytes;
  }
};

// DIST: the fused multi-GPU exchange of the residual (a separate instantiation: the single-GPU kernel pays nothing)
template <class El, int THREADS, int TPI, int PHYS = PHYS_ISO, bool DIST = false>
__global__ void __launch_bounds__(THREADS, (PHYS == PHYS_ISO && 1024 / THREADS > 0 ? 1024 / THREADS : 1))
    k_assemble_iso(const __grid_constant__ AsmArgs a) {
  using IL = IsoLayout<El, TPI>;
  constexpr bool R1 = PHYS == PHYS_R1;  // structured J2 tangent: [lam', mu', kappa, n^(6), 0] per (element, Gauss point)
  constexpr bool GEN = PHYS == PHYS_GENERAL || R1;
  constexpr bool COLORED = IL::COLORED;
  static_assert(PHYS == PHYS_ISO || (GEN && El::DIM == 3), "balanced kernel: isotropic, or general tangent in 3D");
  constexpr int CSTR = R1 ? J2_R1 : IL::CSTR, SGS = IL::SGS;
  constexpr int NNE = IL::NNE, NGP = IL::NGP, DIM = IL::DIM, NV = IL::NV, BLK = IL::BLK, ISTR = IL::ISTR;
  constexpr int GROW = IL::GROW, GSTR = IL::GSTR, ESTR = IL::ESTR, TSTR = IL::TSTR, NH = IL::NH, XSTR = IL::XSTR;
  constexpr int INC = THREADS / TPI;  // incidences per cluster <= INC
  constexpr int GCH = IL::GCH, NCH = IL::NCH;  // Gauss points per chunk, chunks
  static_assert(NCH == 1 || PHYS == PHYS_ISO, "chunked Gauss points: isotropic path only");
  const fdk_plan& p = a.p;
  const int tid = threadIdx.x;
  const bool fuse_ku = a.fuse_ku != 0;
  const bool do_bts = GEN && (a.compute & FDK_VECTOR) && !fuse_ku;  // residual as B^T sigma from a given stress
  const bool per_gp = GEN && (R1 || a.tangent_gp != nullptr);
  [[maybe_unused]] const double* tan_src = R1 ? a.tangent_r1 : a.tangent_gp;
  // the TPI threads of an incidence sit in adjacent lanes: a warp covers 32 / TPI incidences of 2-3 elements
  // (element-major order), so its own-row loads touch 32 / TPI addresses and its column loads ~10
  const int it = PART_UNIFORM ? tid % INC : tid / TPI;    // incidence of this thread (phase 2)
  const int part = PART_UNIFORM ? tid / INC : tid % TPI;  // its column blocks
  // column block jj of this thread: part*NH + jj (contiguous) or part + jj*TPI (interleaved)
  // the thread's two columns are adjacent (2 part, 2 part + 1): 6 contiguous doubles of dN/dx, three 128-bit loads
  constexpr bool ADJ = ISO_COLS_ADJ && NH == 2 && !PART_UNIFORM && DIM == 3;
  auto col = [&](int jj) { return PART_UNIFORM ? part * NH + jj : ((COLORED || ADJ) ? iso_col(part, jj) : part + jj * TPI); };
  constexpr bool HEXREF = IL::HEXREF;  // reflected geometry layout (HexRef)

  extern __shared__ __align__(16) double smem[];
  double* sdN = smem;              // [NGP][TSTR] reference gradients, [g][d][k]
  double* sW = sdN + NGP * TSTR;   // [NGP]
  const int xu_doubles = IL::xu_doubles(p);
  double* sXbuf = smem + IL::TAB_DOUBLES;  // [2][xu_doubles]: padded coordinates, then dofs
  double* sBig = sXbuf + 2 * xu_doubles;
  double* sG = sBig;                       // geometry view  [n_te][ESTR]: sqrt(w) dN/dx, [g][k][d]
  double* sBlk = sBig;                     // staging view   [cap_inc][ISTR]
  double* sR = sBig + IL::sr_offset(p);    // [cap_slots][NV] per-slot K.u products / [cap_inc][NV] nodal forces
  double* sF = sR;
  double* sSig = sR;                       // [cap_te][NGP][6] sqrt(w) sigma (do_bts): shares the space of sF
  double* sC = sR + IL::rs_doubles(p, do_bts);  // [cap_te][NGP][36] tangent of every (touched element, gp) (per_gp)
  double* sXe = sC + IL::extra_doubles(p, per_gp, CSTR);  // [cap_te][NNE][DIM] element-local coordinates (HEXREF)
  long long* sBptr = reinterpret_cast<long long*>(sXe + IL::xe_doubles(p));  // [cap_owned]
  int* sSlotBase = reinterpret_cast<int*>(sBptr + p.cap_owned);                                 // [cap_owned+1]
  int* sFinc = sSlotBase + (p.cap_owned + 1);                                                   // [cap_owned+1]
  unsigned* sRec = reinterpret_cast<unsigned*>(sFinc + (p.cap_owned + 1));                      // [cap_slots+1]
  unsigned* sHeavy = sRec + (p.cap_slots + 1);                                                  // [cap_heavy]
  unsigned char* sLbuf = reinterpret_cast<unsigned char*>(sHeavy + p.cap_heavy);                // [2][lc_bytes]
  const int lc_bytes = (p.cap_te * NNE + 3) & ~3;
  unsigned short* sEnt = reinterpret_cast<unsigned short*>(sLbuf + 2 * lc_bytes);               // [cap_ent]

  // offset of a staged block from its gather-list entry
  auto blk_off = [&](int src) { return COLORED ? src * BLK : (src / NNE) * ISTR + (src % NNE) * BLK; };

  // ---------------- prologue: tables (once per CTA), the first cluster's inputs ----------------
  {
    const ElemTable& tab = c_tab[El::ID];
    for (int t = tid; t < NGP * GROW; t += THREADS) {
      const int g = t / GROW, r = t - g * GROW;
      sdN[g * TSTR + r] = tab.dN[t];
    }
    if (tid < NGP) sW[tid] = tab.w[tid];
  }
  constexpr int RT = (256 + THREADS - 1) / THREADS;  // cap_tn <= 256
  int node_r[RT];
  auto load_node_ids = [&](const ClusterHdr& h) {
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int t = tid + r * THREADS;
      node_r[r] = (t < h.n_tn) ? p.cl_tn_node[h.tn0 + t] : -1;
    }
  };
  // coordinates, dofs and local connectivity of one cluster -> buffer b (one cp.async group)
  auto fetch_inputs = [&](const ClusterHdr& h, int b) {
    double* dX = sXbuf + b * xu_doubles;
    double* dU = dX + p.cap_tn * XSTR;
    unsigned char* dL = sLbuf + b * lc_bytes;
    const unsigned char* lc = p.cl_lconn + (int64_t)h.te0 * NNE;
    if constexpr (NNE % 4 == 0) {
      for (int t = tid; t < h.n_te * NNE / 4; t += THREADS) cp_async<4>(dL + 4 * t, lc + 4 * t);
    } else {
      for (int t = tid; t < h.n_te * NNE; t += THREADS) dL[t] = lc[t];
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int t = tid + r * THREADS;
      const int node = node_r[r];
      if (node >= 0) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) cp_async<8>(dX + t * XSTR + d, a.coords + (int64_t)node * DIM + d);
        if (fuse_ku) {
#pragma unroll
          for (int v = 0; v < DIM; ++v) cp_async<8>(dU + t * DIM + v, a.U + (int64_t)v * p.n_nodes + node);
        }
      }
    }
    cp_async_commit();
  };
  // general tangent: global element of "this thread's" geometry task (task = tid) and the copy of its 36 tangent
  // entries into shared memory; issued one cluster ahead, once the block phase is done with the previous ones
  [[maybe_unused]] int my_te_elem = -1, nxt_te_elem = -1;
  [[maybe_unused]] auto load_te_elem = [&](const ClusterHdr& h) {
    return (tid < h.n_te * NGP) ? p.cl_te_elem[h.te0 + tid / NGP] : -1;
  };
  [[maybe_unused]] auto fetch_tangent = [&](const ClusterHdr& h) {
    // consecutive lanes copy consecutive 16-byte pieces of one (element, gp) tangent: 288 contiguous bytes per task
    // (the element ids are L1 hits: load_te_elem read them a phase ago)
    constexpr int PIECES = CSTR / 2;
    for (int idx = tid; idx < h.n_te * NGP * PIECES; idx += THREADS) {
      const int task = idx / PIECES, k = idx - task * PIECES;
      const int le = task / NGP, g = task - le * NGP;
      const int64_t e = p.cl_te_elem[h.te0 + le];
      cp_async<16>(sC + (long)task * CSTR + 2 * k, tan_src + CSTR * ((int64_t)g * p.n_elems + e) + 2 * k);
    }
  };
  ClusterHdr cur = load_hdr(p.cl_hdr, blockIdx.x);
  load_node_ids(cur);
  fetch_inputs(cur, 0);
  if constexpr (GEN) {
    if (per_gp || do_bts) my_te_elem = load_te_elem(cur);  // (also warms L1 for fetch_tangent)
    if (per_gp) fetch_tangent(cur);
    cp_async_commit();
  }
  cp_async_wait_group<0>();
  __syncthreads();  // tables and the first inputs are visible

  FDK_CLK_DECL
  int buf = 0;
  for (int c = blockIdx.x; c < p.n_clusters; c += gridDim.x, buf ^= 1) {
    const int q0 = cur.q0, n_owned = cur.n_owned, n_te = cur.n_te, n_inc = cur.n_inc;
    const int inc0 = cur.inc0, h0 = cur.h0, n_heavy = cur.n_heavy, n_slots = cur.n_slots;
    const int64_t slot0 = cur.slot0;
    const double* sX = sXbuf + buf * xu_doubles;
    const double* sU = sX + p.cap_tn * XSTR;
    const unsigned char* sLconn = sLbuf + buf * lc_bytes;
    const int c_next = c + gridDim.x;
    const bool has_next = c_next < p.n_clusters;
    ClusterHdr nxt = cur;
    if (has_next) nxt = load_hdr(p.cl_hdr, c_next);  // consumed after phase 1
    unsigned my_desc = 0;
    if (it < n_inc) my_desc = p.inc_desc[inc0 + it];  // consumed in phase 2
    [[maybe_unused]] int my_pos[NH];                   // staging positions of this thread's blocks (after phase 2)
    if constexpr (COLORED) {
      if (it < n_inc) {
#pragma unroll
        for (int j = 0; j < NH; ++j)
          my_pos[j] = ((((it >> 2) * 2 + j) << 4) + (int)p.blk_slot[(int64_t)(inc0 + it) * NNE + col(j)]);
      }
    }
    int my_node = 0;  // row node of this thread's share of the residual reduction (consumed at the very end)
    if ((fuse_ku || do_bts) && (tid >> 3) < n_owned * NV) my_node = p.cl_node[q0 + (tid >> 3) / NV];
    [[maybe_unused]] int64_t my_gid = 0;  // its global id, for the fused exchange (a second dependent load, hidden by phase 1)
    if constexpr (DIST) {
      if ((fuse_ku || do_bts) && (tid >> 3) < n_owned * NV) my_gid = a.node_gid[my_node];
    }
    [[maybe_unused]] unsigned my_fdst = 0;  // node-major rank of the incidence: where its nodal force is parked
    if (do_bts && part == 0 && it < n_inc) my_fdst = p.inc_fdst[inc0 + it];

    // gather descriptors nobody reads between the closing barrier of the previous cluster and this cluster's gather:
    // issued by the threads that have no geometry task (600 tasks on 1024 threads), so they cost the others nothing
    auto fetch_desc = [&](int first, int stride) {
      const unsigned* rec = p.slot_rec + slot0 + c;
      for (int t = first; t <= n_slots; t += stride) cp_async<4>(sRec + t, rec + t);
      for (int t = first; t < n_owned; t += stride) cp_async<8>(sBptr + t, p.cl_bptr + q0 + t);
      // gather lists: staging positions (coloured layout) or block ids it * nne + j; even offset: 4-byte aligned
      const unsigned short* esrc = (COLORED ? p.ent_pos : p.ent_src) + cur.ent0;
      const int n_ent = n_inc * NNE + n_owned;
      for (int t = first; t < (n_ent + 1) / 2; t += stride) cp_async<4>(sEnt + 2 * t, esrc + 2 * t);
      for (int t = first; t < n_heavy; t += stride) cp_async<4>(sHeavy + t, p.heavy_slot + h0 + t);
    };
    const int n_task = n_te * GCH;
    const bool desc_early = n_task + 128 <= THREADS;  // uniform over the CTA
    if (desc_early && tid >= n_task) fetch_desc(tid - n_task, THREADS - n_task);

    double acc[NH][BLK];
    [[maybe_unused]] double f[NV];

#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch) {  // chunks of GCH Gauss points (one chunk unless IL::GCH < NGP)
    // ---------------- phase 1: sqrt(w) dN/dx per (touched element, Gauss point) ----------------
    for (int task = tid; task < n_te * GCH; task += THREADS) {
      const int le = task / GCH, gl = task - le * GCH, g = ch * GCH + gl;
      const unsigned char* lc = sLconn + le * NNE;
      if constexpr (IL::MONO && GCH == NGP) {
        // ---- monomial basis of the trilinear element.  A nodal field is c0 + c1 xi + c2 eta + c3 zeta + c4 xi eta +
        // c5 eta zeta + c6 xi zeta + c7 xi eta zeta with c_m = 1/8 sum_k h_m(k) x_k (h_m: products of the nodes' signs).
        // The 32 tasks of a warp are the 8 Gauss points of 4 elements: twelve of its lanes first compute the seven
        // coefficient vectors of those elements, one (element, component) each -- 8 coordinate loads instead of the 24
        // every task made -- and park them in shared memory; after a warp-level sync every task reads its element's 21
        // coefficients, forms J in 27 FMA (instead of 72 from 24 coordinates and 24 table entries) and builds the
        // reference gradients of the eight shape functions from the Gauss point's signs: no table loads at all.
        const int lane = tid & 31;
        {
          const int left = n_te * GCH - (task - lane);  // tasks of this warp's round: the first `left` lanes are here
          if (lane < 12 && (lane / 3) * 8 < left) {
            const int el = le - (lane >> 3) + lane / 3;  // first element of the warp + lane / 3
            const int d = lane - (lane / 3) * 3;
            const unsigned char* l8 = sLconn + el * NNE;
            double c[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const double x = sX[(int)l8[k] * XSTR + d];
              const double s0 = HexRef::bit(k, 0) ? 1.0 : -1.0, s1 = HexRef::bit(k, 1) ? 1.0 : -1.0, s2 = HexRef::bit(k, 2) ? 1.0 : -1.0;
              c[0] += s0 * x;
              c[1] += s1 * x;
              c[2] += s2 * x;
              c[3] += (s0 * s1) * x;
              c[4] += (s1 * s2) * x;
              c[5] += (s0 * s2) * x;
              c[6] += (s0 * s1 * s2) * x;
            }
            double2* co = reinterpret_cast<double2*>(sXe + el * IL::XESTR + d * 8);  // [component][8]: 7 coefficients + pad
            co[0] = make_double2(0.125 * c[0], 0.125 * c[1]);
            co[1] = make_double2(0.125 * c[2], 0.125 * c[3]);
            co[2] = make_double2(0.125 * c[4], 0.125 * c[5]);
            co[3] = make_double2(0.125 * c[6], 0.0);
          }
          __syncwarp(left >= 32 ? 0xffffffffu : ((1u << left) - 1u));
        }
        constexpr double GA = 0.5773502691896258;
        const double xi = (g & 4) ? GA : -GA, et = (g & 2) ? GA : -GA, ze = (g & 1) ? GA : -GA;
        double J[3][3];
        {
          const double xe_ = xi * et, ez = et * ze, xz = xi * ze;
#pragma unroll
          for (int x = 0; x < 3; ++x) {  // one component at a time: eight doubles live
            const double2* c2 = reinterpret_cast<const double2*>(sXe + le * IL::XESTR + x * 8);
            const double2 c01 = c2[0], c23 = c2[1], c45 = c2[2], c6_ = c2[3];
            // c0..c6 = coefficients of xi, eta, zeta, xi eta, eta zeta, xi zeta, xi eta zeta
            J[0][x] = fma(c6_.x, ez, fma(c45.y, ze, fma(c23.y, et, c01.x)));
            J[1][x] = fma(c6_.x, xz, fma(c45.x, ze, fma(c23.y, xi, c01.y)));
            J[2][x] = fma(c6_.x, xe_, fma(c45.y, xi, fma(c45.x, et, c23.x)));
          }
        }
        double iJ[3][3];
        const double det = invert<3>(J, iJ);
        const double s = sqrt(sW[g] * fabs(det));
        if constexpr (GEN) {
          if (do_bts) {  // sqrt(w) sigma of this Gauss point: with sqrt(w) dN/dx it gives w B^T sigma
            const int64_t e = task == tid ? my_te_elem : p.cl_te_elem[cur.te0 + le];
            const double* sp = a.stress_gp + 6 * ((int64_t)g * p.n_elems + e);
            double* so = sSig + (long)task * SGS;
#pragma unroll
            for (int k = 0; k < 6; ++k) so[k] = s * sp[k];
          }
        }
        {
          const double s8 = 0.125 * s;
#pragma unroll
          for (int x = 0; x < 3; ++x)
#pragma unroll
            for (int r = 0; r < 3; ++r) iJ[x][r] *= s8;
        }
        // 8 dN_k / dxi_r = s_kr f_a(k) f_b(k), f_d(k) = 1 + s_kd xi_d: two values per axis, four products per direction
        const double fx[2] = {1.0 - xi, 1.0 + xi}, fy[2] = {1.0 - et, 1.0 + et}, fz[2] = {1.0 - ze, 1.0 + ze};
        double2* out = reinterpret_cast<double2*>(sG + le * ESTR + gl * GSTR);
#pragma unroll
        for (int k = 0; k < NNE; k += 2) {
          double v[6];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int b0 = HexRef::bit(k + h, 0), b1 = HexRef::bit(k + h, 1), b2 = HexRef::bit(k + h, 2);
            const double A = (b0 ? 1.0 : -1.0) * (fy[b1] * fz[b2]);
            const double B = (b1 ? 1.0 : -1.0) * (fx[b0] * fz[b2]);
            const double C = (b2 ? 1.0 : -1.0) * (fx[b0] * fy[b1]);
#pragma unroll
            for (int x = 0; x < 3; ++x) v[3 * h + x] = fma(iJ[x][2], C, fma(iJ[x][1], B, iJ[x][0] * A));
          }
#pragma unroll
          for (int t = 0; t < 3; ++t) out[(k * 3) / 2 + t] = make_double2(v[2 * t], v[2 * t + 1]);
        }
        continue;
      }
      const double* dN = sdN + g * TSTR;
      int ln[NNE];
      if constexpr (HEXREF) {
        // lane (le, g) copies node g of its element into the element-local array; its 8 lanes sit in one warp
        // (task = le * 8 + g, 8 | 32), so a warp-level sync publishes the copy.  Then position j works on node pi_g(j).
        {
          const double* xk = sX + (int)lc[g] * XSTR;
          const double2 x01 = *reinterpret_cast<const double2*>(xk);
          const double x2 = xk[2];
          double* xe = sXe + le * IL::XESTR + g * DIM;
          xe[0] = x01.x;
          xe[1] = x01.y;
          xe[2] = x2;
        }
        {
          const int left = n_te * GCH - (task - (tid & 31));  // tasks of this warp's round: the first `left` lanes are here
          __syncwarp(left >= 32 ? 0xffffffffu : ((1u << left) - 1u));
        }
        const int gxv = HexRef::gx(g);
#pragma unroll
        for (int j = 0; j < NNE; ++j) ln[j] = HexRef::perm((int)((HexRef::CODES >> (4 * j)) & 7u), gxv);
      } else if constexpr (NNE % 4 == 0) {
        const unsigned* lc4 = reinterpret_cast<const unsigned*>(lc);
#pragma unroll
        for (int q = 0; q < NNE / 4; ++q) {
          const unsigned w4 = lc4[q];
#pragma unroll
          for (int k = 0; k < 4; ++k) ln[4 * q + k] = (w4 >> (8 * k)) & 0xFF;
        }
      } else {
#pragma unroll
        for (int k = 0; k < NNE; ++k) ln[k] = lc[k];
      }
      double J[DIM][DIM];
#pragma unroll
      for (int r = 0; r < DIM; ++r)
#pragma unroll
        for (int x = 0; x < DIM; ++x) J[r][x] = 0.0;
#pragma unroll
      for (int k = 0; k < NNE; k += 2) {
        double X[2][DIM];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if constexpr (HEXREF) {
            const double* xk = sXe + le * IL::XESTR + ln[k + h] * DIM;
#pragma unroll
            for (int x = 0; x < DIM; ++x) X[h][x] = xk[x];
          } else {
            const double* xk = sX + ln[k + h] * XSTR;
            const double2 x01 = *reinterpret_cast<const double2*>(xk);
            X[h][0] = x01.x;
            X[h][1] = x01.y;
            if constexpr (DIM == 3) X[h][2] = xk[2];
          }
        }
#pragma unroll
        for (int r = 0; r < DIM; ++r) {
          double2 dn;
          if constexpr (HEXREF) dn = make_double2(HexRef::dn(r, k), HexRef::dn(r, k + 1));  // compile-time operands
          else dn = *reinterpret_cast<const double2*>(dN + r * NNE + k);                     // nodes k, k+1
#pragma unroll
          for (int x = 0; x < DIM; ++x) J[r][x] = fma(dn.y, X[1][x], fma(dn.x, X[0][x], J[r][x]));
        }
      }
      double iJ[DIM][DIM];
      const double det = invert<DIM>(J, iJ);
      const double s = sqrt(sW[g] * fabs(det));
#pragma unroll
      for (int x = 0; x < DIM; ++x)
#pragma unroll
        for (int r = 0; r < DIM; ++r) iJ[x][r] *= s;
      if constexpr (GEN) {
        if (do_bts) {  // sqrt(w) sigma of this Gauss point: with sqrt(w) dN/dx it gives w B^T sigma
          const int64_t e = task == tid ? my_te_elem : p.cl_te_elem[cur.te0 + le];
          const double* sp = a.stress_gp + 6 * ((int64_t)g * p.n_elems + e);
          double* so = sSig + (long)task * SGS;
#pragma unroll
          for (int k = 0; k < 6; ++k) so[k] = s * sp[k];
        }
      }
      // G[k][x] = sum_r iJ[x][r] dN[r][k], stored [k][x] as 128-bit pairs
      double2* out = reinterpret_cast<double2*>(sG + le * ESTR + gl * GSTR);
#pragma unroll
      for (int k = 0; k < NNE; k += 2) {
        double2 dn[DIM];
#pragma unroll
        for (int r = 0; r < DIM; ++r) {
          if constexpr (HEXREF) dn[r] = make_double2(HexRef::dn(r, k), HexRef::dn(r, k + 1));
          else dn[r] = *reinterpret_cast<const double2*>(dN + r * NNE + k);
        }
        double v[2 * DIM];  // [h][x]
#pragma unroll
        for (int x = 0; x < DIM; ++x) {
          double v0 = iJ[x][0] * dn[0].x, v1 = iJ[x][0] * dn[0].y;
#pragma unroll
          for (int r = 1; r < DIM; ++r) {
            v0 = fma(iJ[x][r], dn[r].x, v0);
            v1 = fma(iJ[x][r], dn[r].y, v1);
          }
          v[x] = v0;
          v[DIM + x] = v1;
        }
#pragma unroll
        for (int t = 0; t < DIM; ++t) out[(k * DIM) / 2 + t] = make_double2(v[2 * t], v[2 * t + 1]);
      }
    }
    if (ch == 0) {
      if (has_next) load_node_ids(nxt);  // consumed after phase 2
      if constexpr (GEN) {
        if (has_next && (per_gp || do_bts)) nxt_te_elem = load_te_elem(nxt);
      }
    }
    __syncthreads();                   // B2: geometry (of this chunk) complete
    FDK_CLK(2)

    // ---------------- the rest of the gather descriptors (land during phase 2) ----------------
    if (ch == 0) {
      if (!desc_early) fetch_desc(tid, THREADS);
      for (int t = tid; t < n_owned; t += THREADS) {  // read by the residual reduction of the previous cluster:
        cp_async<4>(sSlotBase + t, p.cl_slot_loc + q0 + t);  // only now, after every warp has passed it
        if (do_bts) cp_async<4>(sFinc + t, p.cl_finc_loc + q0 + t);
      }
      cp_async_commit();
      if (tid == 0) {
        sSlotBase[n_owned] = n_slots;
        sFinc[n_owned] = n_inc;
      }
    }

    // ---------------- phase 2: NH column blocks of one incidence per thread ----------------
    if (ch == 0) {
#pragma unroll
      for (int j = 0; j < NH; ++j)
#pragma unroll
        for (int b = 0; b < BLK; ++b) acc[j][b] = 0.0;
#pragma unroll
      for (int v = 0; v < NV; ++v) f[v] = 0.0;
    }
    [[maybe_unused]] const unsigned p2_lanes = __ballot_sync(0xffffffffu, it < n_inc);  // this warp's working lanes
    if (it < n_inc) {
      const int le = my_desc & 0xFFF, i = my_desc >> 12;
      const double* gi_p = sG + le * ESTR + i * DIM;
      const double* gj_p = sG + le * ESTR + col(0) * DIM;
      [[maybe_unused]] const int ci = (int)((HexRef::CODES >> (4 * i)) & 7u);  // node code of the row node
      [[maybe_unused]] const double* ge_p = sG + le * ESTR;
      if constexpr (!GEN) {
        constexpr int UNR = HEXREF ? NGP : P2_UNROLL;  // reflected layout: the Gauss point must be a compile-time value
#pragma unroll UNR
        for (int gl = 0; gl < GCH; ++gl) {
          [[maybe_unused]] const int g = ch * GCH + gl;  // Gauss point; gl = its row in the chunk's geometry
          double gi[DIM];
          double gj[NH][DIM];
          if constexpr (HEXREF) {
            // node i sits at position pi_g(i); the thread's column pair (2 part, 2 part + 1) at pair slot
            // part ^ (iy + 2 iz), in swapped order when ix ^ iy (all of it folded at compile time but `part`, `ci`)
            const int gxv = HexRef::gx(g);
            const double* gp = ge_p + gl * GSTR;
            const double* gip = gp + HexRef::perm(ci, gxv) * DIM;
#pragma unroll
            for (int d = 0; d < DIM; ++d) gi[d] = gip[d];
            const double2* g2 = reinterpret_cast<const double2*>(gp + ((part ^ (gxv >> 1)) * 2) * DIM);
            const double2 v0 = g2[0], v1 = g2[1], v2 = g2[2];
            const bool swp = ((gxv ^ (gxv >> 1)) & 1) != 0;
            gj[swp ? 1 : 0][0] = v0.x;
            gj[swp ? 1 : 0][1] = v0.y;
            gj[swp ? 1 : 0][2] = v1.x;
            gj[swp ? 0 : 1][0] = v1.y;
            gj[swp ? 0 : 1][1] = v2.x;
            gj[swp ? 0 : 1][2] = v2.y;
          } else {
#pragma unroll
          for (int d = 0; d < DIM; ++d) gi[d] = gi_p[gl * GSTR + d];
          if constexpr ((PART_UNIFORM || ADJ) && (NH * DIM) % 2 == 0) {  // contiguous columns: 128-bit loads
            const double2* g2 = reinterpret_cast<const double2*>(gj_p + gl * GSTR);
#pragma unroll
            for (int t = 0; t < NH * DIM / 2; ++t) {
              const double2 v = g2[t];
              gj[(2 * t) / DIM][(2 * t) % DIM] = v.x;
              gj[(2 * t + 1) / DIM][(2 * t + 1) % DIM] = v.y;
            }
          } else {
#pragma unroll
            for (int j = 0; j < NH; ++j)
#pragma unroll
              for (int d = 0; d < DIM; ++d) gj[j][d] = gj_p[gl * GSTR + (col(j) - col(0)) * DIM + d];
          }
          }
#pragma unroll
          for (int j = 0; j < NH; ++j)
#pragma unroll
            for (int cc = 0; cc < DIM; ++cc)
#pragma unroll
              for (int aa = 0; aa < DIM; ++aa) acc[j][cc * DIM + aa] = fma(gi[cc], gj[j][aa], acc[j][cc * DIM + aa]);
        }
      } else {
        // K_ij += B_i^T C_g B_j with B of Voigt order [xx, yy, zz, xy, xz, yz] (engineering shears): first
        // t = B_i^T C_g (3 x 6), one tangent column (6 contiguous entries, Fortran order) at a time, then
        // acc_j += t B_j.  sqrt(w) sits in both gradients.
        const double* c_p = per_gp ? sC + (long)le * NGP * CSTR : nullptr;
        const double* s_p = sSig + (long)le * NGP * SGS;
#pragma unroll 1
        for (int g = 0; g < NGP; ++g) {
          double gi[3];
          [[maybe_unused]] int gxv = 0;
          if constexpr (HEXREF) {
            gxv = HexRef::gx(g);
            const double* gip = ge_p + g * GSTR + HexRef::perm(ci, gxv) * 3;
#pragma unroll
            for (int d = 0; d < 3; ++d) gi[d] = gip[d];
          } else {
#pragma unroll
            for (int d = 0; d < 3; ++d) gi[d] = gi_p[g * GSTR + d];
          }
          if (do_bts && part == 0) {  // nodal force of this incidence: B_i^T (w sigma)
            const double* ws = s_p + g * SGS;
            f[0] += ws[0] * gi[0] + ws[3] * gi[1] + ws[4] * gi[2];
            f[1] += ws[1] * gi[1] + ws[3] * gi[0] + ws[5] * gi[2];
            f[2] += ws[2] * gi[2] + ws[4] * gi[0] + ws[5] * gi[1];
          }
          if constexpr (R1) {
            // K_ij += lam' g_i (x) g_j + mu' (g_j (x) g_i + (g_i . g_j) 1) - kappa (n^ g_i) (x) (n^ g_j); sqrt(w) sits in g
            const double2* r2 = reinterpret_cast<const double2*>(c_p + g * CSTR);
            const double2 q0 = r2[0], q1 = r2[1];
            const double lamp = q0.x, mup = q0.y, kap = q1.x;
            if (!__any_sync(p2_lanes, kap != 0.0)) {
              // every Gauss point this warp works on in this step is elastic (kappa is exactly 0 there, and the plastic
              // zone of a structure is spatially coherent, so whole warps are): the flow direction is neither loaded nor
              // multiplied -- 48 instead of 102 DFMA and 6 fewer operands per thread and Gauss point
              static_assert(ADJ || !R1, "structured tangent path: adjacent column blocks");
              double gjv[NH * 3];
              if constexpr (HEXREF) {
                const double2* g2 = reinterpret_cast<const double2*>(ge_p + g * GSTR + ((part ^ (gxv >> 1)) * 2) * 3);
                const double2 v0 = g2[0], v1 = g2[1], v2 = g2[2];
                const bool swp = ((gxv ^ (gxv >> 1)) & 1) != 0;
                gjv[0] = swp ? v1.y : v0.x;
                gjv[1] = swp ? v2.x : v0.y;
                gjv[2] = swp ? v2.y : v1.x;
                gjv[3] = swp ? v0.x : v1.y;
                gjv[4] = swp ? v0.y : v2.x;
                gjv[5] = swp ? v1.x : v2.y;
              } else {
                const double2* g2 = reinterpret_cast<const double2*>(gj_p + g * GSTR);
#pragma unroll
                for (int q = 0; q < NH * 3 / 2; ++q) {
                  const double2 v = g2[q];
                  gjv[2 * q] = v.x;
                  gjv[2 * q + 1] = v.y;
                }
              }
              double li[3], mi[3];
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                li[d] = lamp * gi[d];
                mi[d] = mup * gi[d];
              }
#pragma unroll
              for (int j = 0; j < NH; ++j) {
                const double* gj = gjv + 3 * j;
                const double dot = mi[0] * gj[0] + mi[1] * gj[1] + mi[2] * gj[2];
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
#pragma unroll
                  for (int aa = 0; aa < 3; ++aa) acc[j][cc * 3 + aa] += fma(li[cc], gj[aa], mi[aa] * gj[cc]);
                  acc[j][cc * 3 + cc] += dot;
                }
              }
              continue;
            }
            const double2 q2 = r2[2], q3 = r2[3], q4 = r2[4];
            const double n0 = q1.y, n1 = q2.x, n2 = q2.y, n3 = q3.x, n4 = q3.y, n5 = q4.x;
            double li[3], mi[3], ki[3];
            {
              const double a0 = n0 * gi[0] + n3 * gi[1] + n4 * gi[2];
              const double a1 = n3 * gi[0] + n1 * gi[1] + n5 * gi[2];
              const double a2 = n4 * gi[0] + n5 * gi[1] + n2 * gi[2];
              ki[0] = -kap * a0;
              ki[1] = -kap * a1;
              ki[2] = -kap * a2;
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                li[d] = lamp * gi[d];
                mi[d] = mup * gi[d];
              }
            }
            double gjv[NH * 3];
            if constexpr (HEXREF) {
              const double2* g2 = reinterpret_cast<const double2*>(ge_p + g * GSTR + ((part ^ (gxv >> 1)) * 2) * 3);
              const double2 v0 = g2[0], v1 = g2[1], v2 = g2[2];
              const bool swp = ((gxv ^ (gxv >> 1)) & 1) != 0;
              gjv[0] = swp ? v1.y : v0.x;
              gjv[1] = swp ? v2.x : v0.y;
              gjv[2] = swp ? v2.y : v1.x;
              gjv[3] = swp ? v0.x : v1.y;
              gjv[4] = swp ? v0.y : v2.x;
              gjv[5] = swp ? v1.x : v2.y;
            } else {
              static_assert(ADJ || !R1, "structured tangent path: adjacent column blocks");
              const double2* g2 = reinterpret_cast<const double2*>(gj_p + g * GSTR);
#pragma unroll
              for (int q = 0; q < NH * 3 / 2; ++q) {
                const double2 v = g2[q];
                gjv[2 * q] = v.x;
                gjv[2 * q + 1] = v.y;
              }
            }
#pragma unroll
            for (int j = 0; j < NH; ++j) {
              const double* gj = gjv + 3 * j;
              const double b0 = n0 * gj[0] + n3 * gj[1] + n4 * gj[2];
              const double b1 = n3 * gj[0] + n1 * gj[1] + n5 * gj[2];
              const double b2 = n4 * gj[0] + n5 * gj[1] + n2 * gj[2];
              const double bj[3] = {b0, b1, b2};
              const double dot = mi[0] * gj[0] + mi[1] * gj[1] + mi[2] * gj[2];
#pragma unroll
              for (int cc = 0; cc < 3; ++cc) {
#pragma unroll
                for (int aa = 0; aa < 3; ++aa)
                  acc[j][cc * 3 + aa] += fma(li[cc], gj[aa], fma(mi[aa], gj[cc], ki[cc] * bj[aa]));
                acc[j][cc * 3 + cc] += dot;
              }
            }
            continue;
          }
          double t[3][6];
#pragma unroll
          for (int sc = 0; sc < 6; ++sc) {
            double c0, c1, c2, c3, c4, c5;  // C[0..5][sc]
            if (per_gp) {
              const double2* cc2 = reinterpret_cast<const double2*>(c_p + g * CSTR + 6 * sc);
              const double2 u0 = cc2[0], u1 = cc2[1], u2 = cc2[2];
              c0 = u0.x; c1 = u0.y; c2 = u1.x; c3 = u1.y; c4 = u2.x; c5 = u2.y;
            } else {
              c0 = a.C[0 * 6 + sc]; c1 = a.C[1 * 6 + sc]; c2 = a.C[2 * 6 + sc];
              c3 = a.C[3 * 6 + sc]; c4 = a.C[4 * 6 + sc]; c5 = a.C[5 * 6 + sc];
            }
            t[0][sc] = gi[0] * c0 + gi[1] * c3 + gi[2] * c4;
            t[1][sc] = gi[1] * c1 + gi[0] * c3 + gi[2] * c5;
            t[2][sc] = gi[2] * c2 + gi[0] * c4 + gi[1] * c5;
          }
          [[maybe_unused]] double gjv[NH * 3];
          if constexpr (HEXREF) {  // pair slot part ^ (iy + 2 iz); order swapped when ix ^ iy (run-time here)
            const double2* g2 = reinterpret_cast<const double2*>(ge_p + g * GSTR + ((part ^ (gxv >> 1)) * 2) * 3);
            const double2 v0 = g2[0], v1 = g2[1], v2 = g2[2];
            const bool swp = ((gxv ^ (gxv >> 1)) & 1) != 0;
            gjv[0] = swp ? v1.y : v0.x;
            gjv[1] = swp ? v2.x : v0.y;
            gjv[2] = swp ? v2.y : v1.x;
            gjv[3] = swp ? v0.x : v1.y;
            gjv[4] = swp ? v0.y : v2.x;
            gjv[5] = swp ? v1.x : v2.y;
          } else if constexpr (ADJ) {
            const double2* g2 = reinterpret_cast<const double2*>(gj_p + g * GSTR);
#pragma unroll
            for (int q = 0; q < NH * 3 / 2; ++q) {
              const double2 v = g2[q];
              gjv[2 * q] = v.x;
              gjv[2 * q + 1] = v.y;
            }
          }
#pragma unroll
          for (int j = 0; j < NH; ++j) {
            double gj[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) gj[d] = ADJ ? gjv[j * 3 + d] : gj_p[g * GSTR + (col(j) - col(0)) * 3 + d];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
              acc[j][cc * 3 + 0] += t[cc][0] * gj[0] + t[cc][3] * gj[1] + t[cc][4] * gj[2];
              acc[j][cc * 3 + 1] += t[cc][1] * gj[1] + t[cc][3] * gj[0] + t[cc][5] * gj[2];
              acc[j][cc * 3 + 2] += t[cc][2] * gj[2] + t[cc][4] * gj[0] + t[cc][5] * gj[1];
            }
          }
        }
      }
    }
    if (ch + 1 < NCH) __syncthreads();  // the next chunk's geometry overwrites this one
    }  // chunks
    if (has_next) fetch_inputs(nxt, buf ^ 1);  // lands during the gather
    else cp_async_commit();
    __syncthreads();  // B3: everyone is done reading the geometry; the region becomes the staging array
    FDK_CLK(3)
    if (it < n_inc) {
#pragma unroll
      for (int j = 0; j < NH; ++j) {
        double* sp = COLORED ? sBlk + my_pos[j] * BLK : sBlk + it * ISTR + col(j) * BLK;
#pragma unroll
        for (int b = 0; b < BLK; ++b) sp[b] = acc[j][b];
      }
      if constexpr (GEN) {
        if (do_bts && part == 0) {
#pragma unroll
          for (int v = 0; v < NV; ++v) sF[my_fdst * NV + v] = f[v];
        }
      }
    }
    if constexpr (GEN) {  // every warp is past the block phase: the tangent buffer may take the next cluster's entries
      if (per_gp && has_next) fetch_tangent(nxt);
      cp_async_commit();
      cp_async_wait_group<2>();  // this cluster's descriptors (two younger groups may still be in flight)
    } else {
      cp_async_wait_group<1>();  // this cluster's descriptors (the next cluster's inputs may still be in flight)
    }
    __syncthreads();           // B4: staging and descriptors complete
    FDK_CLK(4)

    // ---------------- phase 3a: heavy slots, all loads of a slot entry in flight together ----------------
    if (n_heavy > 0) {  // uniform over the CTA
      for (int t = tid; t < n_heavy * BLK; t += THREADS) {
        const int h = t / BLK, b = t - h * BLK;
        const int s = sHeavy[h];
        const unsigned r0 = sRec[s], r1 = sRec[s + 1];
        const int e0 = r0 & 0xFFFF;
        const int e1 = (int)(r1 & 0xFFFF) - (((r0 ^ r1) >> 24) ? 1 : 0);
        double v = 0.0;
        for (int eb = e0; eb < e1; eb += 8) {
          int src[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) src[k] = sEnt[eb + k < e1 ? eb + k : e0];
          double x[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) x[k] = sBlk[blk_off(src[k]) + b];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (eb + k < e1) v += x[k];
        }
        // in place: only this thread touches column b of the entries of slot h
        const int src0 = sEnt[e0];
        sBlk[blk_off(src0) + b] = v;
      }
      __syncthreads();
      FDK_CLK(5)
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
      const double* bp_[HEAVY_T];
#pragma unroll
      for (int t = 0; t < HEAVY_T; ++t) {
        const int src = sEnt[e0 + (t < cnt ? t : 0)];
        bp_[t] = sBlk + blk_off(src);
      }
      double S[BLK];
#pragma unroll
      for (int b = 0; b < BLK; ++b) S[b] = bp_[0][b];
#pragma unroll
      for (int t = 1; t < HEAVY_T; ++t) {
        if (t < cnt) {
#pragma unroll
          for (int b = 0; b < BLK; ++b) S[b] += bp_[t][b];
        }
      }
      double Kb[BLK];
      if constexpr (GEN) {
#pragma unroll
        for (int b = 0; b < BLK; ++b) Kb[b] = S[b];
      } else {  // K_IJ = lambda S + mu S^T + mu tr(S) 1
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
      }
#pragma unroll
      for (int cc = 0; cc < NV; ++cc) {
        double* row = a.K + ((int64_t)cc * NV * p.blk_nnz + (int64_t)NV * bp);
#pragma unroll
        for (int aa = 0; aa < NV; ++aa) __stcs(row + (int64_t)aa * deg + pcol, Kb[cc * NV + aa]);
      }
      if (fuse_ku) {  // this slot's share of (K U)_I
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
    cp_async_wait_group<0>();  // the next cluster's inputs: published by the barrier below
    __syncthreads();           // B6: gather done (staging free), sR complete, next inputs visible
    FDK_CLK(6)

    // ---------------- residual: D_I = -sum over the slots of row I, 8 lanes per (node, component) ----------------
    if (fuse_ku || do_bts) {
      const int* sBase = fuse_ku ? sSlotBase : sFinc;  // per-slot K.u products or per-incidence nodal forces
      const int n_out = n_owned * NV * 8;
      for (int t0 = 0; t0 < n_out; t0 += THREADS) {  // whole warps take part in the shuffles
        const int t = t0 + tid;
        const int sub = t & 7, o = t >> 3;
        const int n = o / NV, v = o - n * NV;
        double sum = 0.0;
        if (t < n_out) {
          const int k1 = sBase[n + 1];
          for (int k = sBase[n] + sub; k < k1; k += 8) sum += sR[k * NV + v];
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        if (t < n_out && sub == 0) {
          const int node = (t0 == 0) ? my_node : p.cl_node[q0 + n];
          a.D[(int64_t)v * p.n_nodes + node] = -sum;
          if constexpr (DIST) {  // fused exchange: straight into every rank's global vector over NVLink
            const int64_t gid = (t0 == 0) ? my_gid : a.node_gid[node];
            for (int r = 0; r < a.n_dst; ++r) {
              double* dst = a.D_dst[r] + (int64_t)v * a.n_dst_nodes + gid;
              *dst = -sum;  // weak store: visible to the peers once the kernel and the barrier behind it are through
            }
          }
        }
      }
    }
    FDK_CLK(7)
    if constexpr (GEN) {
      if (do_bts) __syncthreads();  // the next geometry phase overwrites the parked forces with sqrt(w) sigma
    }
    // no closing barrier (otherwise): the next cluster's phase 1 writes only the geometry view (below sR), its
    // descriptor copies are issued after its B2, which every warp reaches after finishing this reduction
    cur = nxt;
    if constexpr (GEN) my_te_elem = nxt_te_elem;
  }  // cluster loop
  FDK_CLK_FLUSH
}

// does the balanced kernel fit this plan (threads per cluster, shared memory)?  Otherwise the caller uses k_assemble.
template <class El, int THREADS, int TPI, int PHYS = PHYS_ISO>
bool assemble_iso_fits(const AsmArgs& a) {
  using IL = IsoLayout<El, TPI>;
  constexpr bool GENP = PHYS == PHYS_GENERAL || PHYS == PHYS_R1;
  const bool bts = GENP && (a.compute & FDK_VECTOR) && !a.fuse_ku;
  const bool staged = PHYS == PHYS_R1 || (PHYS == PHYS_GENERAL && a.tangent_gp != nullptr);
  return TPI * a.p.cap_inc <= THREADS &&
         IL::smem_bytes(a.p, staged, bts, PHYS == PHYS_R1 ? J2_R1 : IL::CSTR) <= 227 * 1024;
}

template <class El, int THREADS, int TPI, int PHYS = PHYS_ISO, bool DIST = false>
int launch_assemble_iso(AsmArgs& a, cudaStream_t stream) {
  using IL = IsoLayout<El, TPI>;
  const fdk_plan& p = a.p;
  FDK_REQUIRE(TPI * p.cap_inc <= THREADS, FDK_ECAP, "cluster with %d incidences exceeds %d threads / %d", p.cap_inc,
              THREADS, TPI);
  FDK_REQUIRE(p.cap_te < 4096 && p.cap_tn <= 256 && p.cap_owned < 255 && p.cap_ent < 65536 && p.cap_slots < 65535,
              FDK_ECAP, "cluster capacity overflow (te=%d tn=%d owned=%d ent=%d slots=%d)", p.cap_te, p.cap_tn,
              p.cap_owned, p.cap_ent, p.cap_slots);
  FDK_REQUIRE(p.nvar == IL::NV, FDK_EINVAL, "plan nvar %d does not match the operator (%d)", p.nvar, IL::NV);
  FDK_REQUIRE(!IL::COLORED || (p.blk_slot && p.ent_pos), FDK_EINVAL,
              "the plan carries no block colouring (fdk_plan_color_blocks)");
  constexpr bool GENP = PHYS == PHYS_GENERAL || PHYS == PHYS_R1;
  const bool bts = GENP && (a.compute & FDK_VECTOR) && !a.fuse_ku;
  const bool staged = PHYS == PHYS_R1 || (PHYS == PHYS_GENERAL && a.tangent_gp != nullptr);
  const size_t smem = IL::smem_bytes(p, staged, bts, PHYS == PHYS_R1 ? J2_R1 : IL::CSTR);
  FDK_REQUIRE(smem <= 227 * 1024, FDK_ECAP, "cluster needs %zu bytes of shared memory (> 227 KB)", smem);
  if (p.n_clusters == 0) return 0;
  if (int rc = ensure_device_tables()) return rc;
  auto kern = DIST ? k_assemble_iso<El, THREADS, TPI, PHYS, true> : k_assemble_iso<El, THREADS, TPI, PHYS, false>;
  static thread_local size_t smem_set = 0;
  if (smem > smem_set) {
    FDK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  // persistent CTAs: as many as are resident at once (one per SM for the 1024-thread variants)
  static thread_local int resident = 0;
  static thread_local size_t resident_smem = 0;
  if (resident == 0 || resident_smem != smem) {  // the occupancy depends on the plan's shared-memory footprint
    resident_smem = smem;
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

}  // namespace fdk
