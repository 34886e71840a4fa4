/*
 * fdk.h -- C ABI of libfdk (fedoo_b200/_fdk.so): B200-native (sm_100a) kernels for
 * fedoo's global-operator assembly hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _h;
 *   - every function returns 0 on success or a negative FDK_E* / a positive
 *     cudaError_t; fdk_last_error_string() describes the last failure on the
 *     calling thread;
 *   - nothing here allocates or frees caller buffers; the symbolic functions
 *     allocate and release their own temporary workspace;
 *   - all numeric kernels are launched on the given stream and return
 *     asynchronously; they are deterministic (no floating-point atomics): every
 *     CSR value and every entry of the global vector is written exactly once.
 *   - dof ordering is variable-major: dof = var * n_nodes + node
 *     (fedoo/core/problem.py:89-91); Gauss-point ordering is gp-major:
 *     gp_index = gp * n_elems + elem (fedoo/core/mesh.py:1137-1140); Voigt order
 *     [xx, yy, zz, xy, xz, yz] with engineering shear strains
 *     (fedoo/core/modelingspace.py:289-316).
 *
 * Each entry point names the reference interface it replaces.
 */
#ifndef FDK_H
#define FDK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fdk_stream_t; /* cudaStream_t */

enum fdk_elem_type { FDK_HEX8 = 0, FDK_TET4 = 1, FDK_TET10 = 2, FDK_QUAD4 = 3 };

enum fdk_error {
  FDK_OK = 0,
  FDK_EINVAL = -1,   /* bad argument */
  FDK_ECAP = -2,     /* a cluster exceeds the shared-memory capacity of the kernel */
  FDK_EOVERFLOW = -3 /* an index does not fit its integer type */
};

/* compute flags of fdk_assemble_* (fedoo/core/assembly.py:143-151: compute = all|matrix|vector) */
enum fdk_compute { FDK_MATRIX = 1, FDK_VECTOR = 2, FDK_ALL = 3 };

const char* fdk_last_error_string(void);
int fdk_version(void);

/* Runtime options (process-wide): "fuse_ku" (default 1): for a linear law with K and D both requested the
 * residual is taken from the assembled rows, D = -K_row . U, instead of a second B^T sigma integration;
 * "mma" (default 1): hex8 + isotropic law in the generic cluster kernel: element matrices by FP64 tensor-core
 * DMMA.m8n8k4; "iso4" (default 1): hex8 + isotropic law + matrix requested (residual fused): the balanced
 * 1024-thread kernel (4 threads per incidence) instead of the generic one; "j2_continuum_tangent" (default 0):
 * fdk_j2_update returns the continuum elastoplastic tangent L - (L:n)(n:L)/(n:L:n + R') at the end state (what
 * simcoon's cutting-plane EPICP umat returns) instead of the consistent tangent of the radial return. */
int fdk_set_option(const char* key, int value);
int fdk_get_option(const char* key, int* value);

/* Diagnostic (builds with -DFDK_PHASE_CLOCKS only, FDK_EINVAL otherwise): SM cycles spent between the barriers
 * of the cluster kernel, summed over all CTAs since the last reset; out_h receives the first n (<= 16) counters. */
int fdk_debug_phase_clocks(unsigned long long* out_h, int n, int reset);

/* element table accessors (host): what the kernels integrate with.
 * Replaces fedoo/lib_elements/{hexahedron,tetrahedron,quadrangle}.py tables
 * (hexahedron.py:22-27,134-135,178-248; tetrahedron.py:21-61,72-97,106-208;
 * quadrangle.py:24-26,125-167).  out_dN_h: [ngp][dim][nne], out_N_h: [ngp][nne]. */
int fdk_element_info(int elem_type, int* nne, int* ngp, int* dim);
int fdk_element_table(int elem_type, double* out_w_h, double* out_N_h, double* out_dN_h);

/* ------------------------------------------------------------------------- *
 * Symbolic phase (one-time): CSR pattern, bit-exact with the reference's
 * scipy-built one.  Replaces _BlocSparse.tocsr symbolic part
 * (fedoo/core/_sparsematrix.py:225-284: key = row*n_cols+col, unique, bincount)
 * and scipy.sparse.bmat tiling (:310-315).
 * ------------------------------------------------------------------------- */

/* Step 1: number of distinct (I,J) node pairs sharing an element.
 * conn: [n_elems][nne] int32.  keys_out: caller buffer of n_elems*nne*nne uint64
 * receiving the sorted unique keys I*n_nodes+J in its first *blk_nnz_h entries. */
int fdk_sym_block_keys(int n_nodes, int64_t n_elems, int nne, const int32_t* conn, uint64_t* keys_out,
                       int64_t* blk_nnz_h, fdk_stream_t stream);

/* Step 2: block CSR from the unique keys: blk_indptr [n_nodes+1] int64, blk_indices [blk_nnz] int32. */
int fdk_sym_block_csr(int n_nodes, int64_t blk_nnz, const uint64_t* keys, int64_t* blk_indptr, int32_t* blk_indices,
                      fdk_stream_t stream);

/* Step 3: nvar x nvar tiling -> global CSR of shape (nvar*n_nodes + n_global_dof)^2.
 * index_bytes = 4 (int32) or 8 (int64): the caller applies scipy's get_index_dtype rule
 * (int32 iff max(nnz, n_rows) <= 2^31-1).  indptr has nvar*n_nodes+n_global_dof+1 entries. */
int fdk_sym_expand_csr(int n_nodes, int nvar, int n_global_dof, int64_t blk_nnz, const int64_t* blk_indptr,
                       const int32_t* blk_indices, int index_bytes, void* indptr, void* indices, fdk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Assembly plan: node clusters (one CTA each) with everything the numeric
 * kernels need, built once per mesh by the host (fedoo_b200/plan.py).
 * ------------------------------------------------------------------------- */
typedef struct fdk_plan {
  int32_t elem_type;  /* enum fdk_elem_type */
  int32_t n_nodes;
  int64_t n_elems;
  int32_t n_clusters;
  int32_t nvar;       /* variables per node of the assembled operator (dim for elasticity, 1 for heat) */
  int64_t blk_nnz;    /* nnz of the node-node block pattern */
  /* capacities = max over clusters (sizes the dynamic shared memory) */
  int32_t cap_te, cap_tn, cap_inc, cap_owned, cap_slots, cap_ent, cap_heavy;
  int32_t threads;    /* CTA size the clusters were sized for (2 * cap_inc <= threads) */
  const int32_t* cl_hdr;       /* [n_clusters][16] packed per-cluster header: q0, n_owned, te0, n_te, tn0,
                                  n_tn, inc0, n_inc, heavy0, n_heavy, slot0 (lo, hi), n_slots, ent0, 0, 0  */
  const int32_t* cl_node_ptr;  /* [n_clusters+1] range of owned nodes (cluster order)                    */
  const int32_t* cl_node;      /* [n_owned]  global id of the q-th owned node                             */
  const int64_t* cl_bptr;      /* [n_owned]  blk_indptr[cl_node[q]]                                       */
  const int64_t* cl_slot_ptr;  /* [n_owned+1] exclusive cumsum of block-row lengths, cluster order        */
  const int32_t* cl_finc_ptr;  /* [n_owned+1] exclusive cumsum of incidences per owned node               */
  const int32_t* cl_slot_loc;  /* [n_owned] cl_slot_ptr relative to the first slot of the node's cluster  */
  const int32_t* cl_finc_loc;  /* [n_owned] cl_finc_ptr relative to the first incidence of the cluster    */
  const int32_t* cl_inc_ptr;   /* [n_clusters+1] range of incidences (= threads), element-major order     */
  const uint16_t* inc_desc;    /* [n_inc] local touched-element index | local node << 12                  */
  const uint16_t* ent_src;     /* gather lists: for every cluster, at ent0, its (incidence, local column
                                  node) blocks sorted by CSR slot, as local_thread * nne + j; one unused
                                  gap entry closes every block row; ent0 is even                          */
  const uint16_t* inc_fdst;    /* [n_inc] node-major rank of the incidence inside its cluster             */
  const int32_t* cl_te_ptr;    /* [n_clusters+1] range of touched elements                                */
  const int32_t* cl_te_elem;   /* global element id of each touched element                               */
  const uint8_t* cl_te_own;    /* 1 if this cluster is the unique owner of the touched element            */
  const uint8_t* cl_lconn;     /* [n_te_total][nne] local (cluster) index of each element node            */
  const uint32_t* te_desc;     /* [n_te_total] cluster-local index of the element's first incidence (the
                                  incidences of one element are consecutive: element-major thread order)
                                  | mask << 16, bit i of mask set: local node i is owned by the cluster   */
  const int32_t* cl_tn_ptr;    /* [n_clusters+1] range of touched nodes                                   */
  const int32_t* cl_tn_node;   /* global node id of each touched node                                     */
  const uint32_t* slot_rec;    /* per cluster n_slots+1 records at index cl_slot_ptr[q0] + cluster:
                                  first gather-list entry of the slot (entries are slot-sorted, one gap entry
                                  after every block row) | cluster-local touched-node index of the slot's
                                  column node << 16 | cluster-local index of the owner (row) node << 24;
                                  the last record of a cluster is an end sentinel with owner 0xFF          */
  const int32_t* cl_heavy_ptr; /* [n_clusters+1] range of heavy slots (more than 4 contributions)         */
  const uint32_t* heavy_slot;  /* cluster-local slot index of each heavy slot                             */
  const uint8_t* blk_slot;     /* hex8 balanced kernel (fdk_plan_color_blocks), may be NULL: [n_inc][nne]
                                  slot (0..15) of block (incidence, j) inside the 16-block staging window of
                                  its producer group (incidence / 4) * 2 + j / 4                          */
  const uint16_t* ent_pos;     /* same indexing as ent_src: staging POSITION of the entry's block         */
} fdk_plan;

/* Fills blk_slot / ent_pos (device arrays sized like inc_desc * nne and ent_src) for a hex8 plan: a bank-conflict-free
 * layout of the staging array of the balanced kernel (csrc/fdk_color.cuh).  One-time, per plan; the two pointers of
 * *plan are ignored on input. */
int fdk_plan_color_blocks(const fdk_plan* plan, uint8_t* blk_slot, uint16_t* ent_pos, fdk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Numeric assembly (per step).  Replaces Assembly.assemble_global_mat
 * (fedoo/core/assembly.py:143-470) + _BlocSparse.addToBlocATB / tocsr numeric part
 * (fedoo/core/_sparsematrix.py:55-174, 286-315) + the vector branch
 * (fedoo/core/assembly.py:400-411) for StressEquilibrium (small strain,
 * fedoo/weakform/stress_equilibrium.py:92-145) and HeatEquation
 * (fedoo/weakform/heat_equation.py:78-119,168-187).
 *
 * coords: [n_nodes][dim] double (fedoo Mesh.nodes layout).
 * K_values: CSR data of the tiled pattern (nvar*nvar*blk_nnz doubles) or NULL.
 * D: global vector (nvar*n_nodes doubles) or NULL; D = -int B^T sigma.
 * ------------------------------------------------------------------------- */

/* Isotropic linear elasticity, uniform (lambda, mu): K from the closed form
 * K[(c,I),(a,J)] = sum_g w (lambda G_cI G_aJ + mu G_aI G_cJ + mu d_ca G_I.G_J)
 * (= ElasticIsotrop tangent, fedoo/constitutivelaw/elastic_isotrop.py:35-68; for plane
 * stress pass lambda = nu E/(1-nu^2)).  The residual uses sigma = C eps(U) computed on
 * the fly (fused update, fedoo/weakform/stress_equilibrium.py:485-524,589-601 +
 * fedoo/constitutivelaw/elastic_anisotropic.py:36-56) unless stress_gp != NULL, in
 * which case sigma is read from stress_gp [6][n_gp] (column-major (6,N), gp-major
 * columns = sv["Stress"] layout). */
int fdk_assemble_elastic_iso(const fdk_plan* plan, int compute, const double* coords, double lambda, double mu,
                             const double* U, const double* stress_gp, double* K_values, double* D,
                             fdk_stream_t stream);

/* The same with the exchange of the residual FUSED into the kernel (multi-GPU, one process per GPU): every owned
 * entry of D is also stored at var * n_global_nodes + node_gid[node] of n_dst destination vectors -- one NVLink
 * multicast address of a symmetric allocation (NVSwitch replicates each store into all GPUs' copies) or the peers'
 * own addresses (n_dst <= 8, host array of device pointers).  node_gid: int64 [plan->n_nodes] rank-local node ->
 * global node.  The caller synchronises the ranks afterwards (a device-side barrier on the same stream).  Served by
 * the balanced hex8 kernel only (matrix requested, residual from U): FDK_EINVAL otherwise.  No reference
 * counterpart (the reference is single-process). */
int fdk_assemble_elastic_iso_dist(const fdk_plan* plan, int compute, const double* coords, double lambda, double mu,
                                  const double* U, double* K_values, double* D, double* const* D_dst_h, int n_dst,
                                  const int64_t* node_gid, int64_t n_global_nodes, fdk_stream_t stream);

/* General tangent: C is 6x6 row-major uniform (tangent_gp == NULL) or per Gauss point
 * tangent_gp [(6,6,N) Fortran order: C_ij of GP n at i + 6 j + 36 n] = sv["TangentMatrix"]
 * (fedoo/constitutivelaw/simcoon_umat.py:556-580, elastic_anisotropic.py:19-31).
 * K = sum_g w B^T C_g B (full, no symmetry assumption). Residual as above. */
int fdk_assemble_elastic_general(const fdk_plan* plan, int compute, const double* coords, const double* C_h,
                                 const double* tangent_gp, const double* U, const double* stress_gp,
                                 double* K_values, double* D, fdk_stream_t stream);

/* Heat equation: K_IJ = sum_g w grad N_I . k grad N_J + d_IJ (rho c/dt) sum_g w N_I sum_J N_J
 * (row-sum lumping, fedoo/core/_sparsematrix.py:91-98); D_I = -sum_g w [grad N_I . k grad T
 * + (rho c/dt) N_I (T_g - Tstart_g)].  cond_h: 3x3 row-major host array. rho_c_over_dt = 0 for
 * the steady problem. */
int fdk_assemble_heat(const fdk_plan* plan, int compute, const double* coords, const double* cond_h,
                      double rho_c_over_dt, const double* T, const double* T_start, double* K_values, double* D,
                      fdk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Gauss-point state update.  Replaces Assembly.get_gp_results / get_grad_disp
 * (fedoo/core/assembly.py:1045-1112,1285-1336), _comp_linear_strain
 * (fedoo/weakform/stress_equilibrium.py:589-601) and ElasticAnisotropic.update
 * (fedoo/constitutivelaw/elastic_anisotropic.py:36-56).
 * conn: [n_elems][nne] int32. Outputs (any may be NULL): grad_gp [9][n_gp] rows a*3+b = du_a/dx_b
 * (row-major (9,N)), strain_gp / stress_gp column-major (6,N).  C_h 6x6 row-major host,
 * or tangent_gp per GP as above. */
int fdk_gp_strain_stress(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                         const double* U, const double* C_h, const double* tangent_gp, double* grad_gp,
                         double* strain_gp, double* stress_gp, fdk_stream_t stream);

/* The same with the small-strain F-bar option of StressEquilibrium (wf.fbar = True,
 * fedoo/weakform/stress_equilibrium.py:213-214,527-540): the volumetric part of grad u is replaced by its mean over the
 * element's Gauss points before strain and stress are formed.  fbar_center: n_elems doubles of device scratch (returns
 * those means).  3-D elements only. */
int fdk_gp_strain_stress_fbar(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                              const double* U, const double* C_h, const double* tangent_gp, double* fbar_center,
                              double* grad_gp, double* strain_gp, double* stress_gp, fdk_stream_t stream);

/* Residual alone, D = -int B^T sigma (compute = "vector" of Assembly.assemble_global_mat, fedoo/core/assembly.py:400-411;
 * asked for at every Newton sub-iteration, fedoo/problem/non_linear.py:400-404) without the cluster plan of the matrix
 * kernels: element forces into fe_scratch (n_elems * nne * dim doubles), then each node sums its incidences in the order
 * of node_inc (node_ptr int64 [n_nodes+1], node_inc int32 = element * nne + local node): no atomics, reproducible.
 * sigma from stress_gp (6,N), or recomputed from U with C_h (6x6 host) or tangent_gp (6,6,N).  Writes the nvar * n_nodes
 * nodal entries of D. */
int fdk_residual_elastic(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                         const double* C_h, const double* tangent_gp, const double* U, const double* stress_gp,
                         const int64_t* node_ptr, const int32_t* node_inc, double* fe_scratch, double* D,
                         fdk_stream_t stream);

/* The same for the heat equation (fedoo/weakform/heat_equation.py:78-119,168-187): D_I = -sum_g w [grad N_I . (cond
 * grad T) + (rho c / dt) N_I (T_g - T_start,g)]; cond_h 3x3 row-major host; T_start NULL = zero start temperature (the reference's __temp_start = 0, heat_equation.py:140-147). */
int fdk_residual_heat(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                      const double* cond_h, double rho_c_over_dt, const double* T, const double* T_start,
                      const int64_t* node_ptr, const int32_t* node_inc, double* fe_scratch, double* D,
                      fdk_stream_t stream);

/* Rows of HIGH-VALENCE nodes (csrc/fdk_rows.cuh): the cluster plan leaves out the nodes whose incident elements alone
 * exceed a cluster's shared memory (unstructured tet meshes have a few; the reference's util/meshes/octet_truss_quad.msh
 * does) and this entry assembles their block rows, one CTA per node, any valence.  rows[n_rows]: the nodes; node_ptr /
 * node_inc: incidences element * nne + local node, element-ascending; K_values in the tiled layout of fdk_plan;
 * law: isotropic (lam, mu), or C_h (6x6 row-major host) or tangent_gp (6,6,N).  Vector: -int B^T stress_gp, or
 * -K_row . U when no stress is given (exact for every tangent, the state being linear in U).  Same arithmetic and the
 * same reference lines as fdk_assemble_elastic_iso / _general. */
int fdk_assemble_rows_elastic(int elem_type, int n_rows, const int32_t* rows, int n_nodes, int64_t n_elems,
                              const int32_t* conn, const double* coords, const int64_t* node_ptr, const int32_t* node_inc,
                              const int64_t* blk_indptr, const int32_t* blk_indices, int64_t blk_nnz, int max_row_degree,
                              int isotropic, double lam, double mu, const double* C_h, const double* tangent_gp,
                              int compute, const double* U, const double* stress_gp, double* K_values, double* D,
                              fdk_stream_t stream);

/* Heat equation on a tet4 mesh, matrix and / or residual in ONE launch of the row-owner kernel (csrc/fdk_heat_tet4.cuh):
 * K_IJ = sum_e V_e grad N_I . (cond grad N_J) + delta_IJ (rho c / dt) V_e / 4 (lumped capacity, the reference's
 * mat_lumping = [False, True], fedoo/weakform/heat_equation.py:78-119,168-227, core/_sparsematrix.py:91-98) and D as in
 * fdk_residual_heat.  One thread per node row walks its incidences: node_ptr [n_nodes + 1] into inc_rec, an array of
 * int32 pairs {element * 4 + local node, positions (4 x u8, little endian) of the columns conn[e][0..3] inside the
 * node's row of the block pattern blk_indptr / blk_indices (fdk_sym_block_csr)}, element-ascending per node;
 * max_row_degree <= 255.  K_values [blk_nnz] in the order of the pattern (nvar = 1: the CSR itself).  No cluster plan
 * is needed.  T_start NULL = 0.  rows [n_rows] (device) restricts the launch to the listed node rows -- a rank's owned
 * nodes in a multi-GPU partition; the other rows of K_values and D are left untouched; NULL = every row (n_rows =
 * n_nodes or 0). */
int fdk_assemble_heat_tet4(int compute, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                           const double* cond_h, double rho_c_over_dt, const double* T, const double* T_start,
                           const int64_t* node_ptr, const int32_t* inc_rec, const int64_t* blk_indptr,
                           int max_row_degree, int n_rows, const int32_t* rows, double* K_values, double* D,
                           fdk_stream_t stream);

/* The heat residual from GIVEN Gauss-point fields, as the reference's weak forms pass them through assembly.sv
 * (fedoo/weakform/heat_equation.py:99-117 "grad v . (K TempGradient)", :178-186 "(rho c / dt) v (Temp - Temp_start)"):
 * D_I = -sum_g w [grad N_I . flux_g + N_I src_g]; flux_gp [3][n_gp] row-major (NULL = 0), src_gp [n_gp] (NULL = 0),
 * gp-major columns g * n_elems + e.  This is the entry the adapter under the real fedoo.Assembly calls. */
int fdk_residual_heat_gp(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                         const double* flux_gp, const double* src_gp, const int64_t* node_ptr, const int32_t* node_inc,
                         double* fe_scratch, double* D, fdk_stream_t stream);

/* Thermal state: temp_gp [n_gp] and temp_gradient_gp [3][n_gp] (row-major)
 * (fedoo/weakform/heat_equation.py:64-70,149-152). */
int fdk_gp_temperature(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                       const double* T, double* temp_gp, double* temp_gradient_gp, fdk_stream_t stream);

/* Finite-strain kinematics the reference computes in its own Python before handing over to simcoon
 * (fedoo/weakform/stress_equilibrium.py:542-586, _comp_F / _comp_Fbar): F = 1 + grad u at every Gauss point and, with
 * fbar != 0, the F-bar form F (J_mean / J)^(1/3), J = det F, J_mean = the mean of J over the element's Gauss points.
 * F_gp [n_gp][9], entry i + 3 j of Gauss point g * n_elems + e = F_ij (the memory of the reference's Fortran-ordered
 * (3, 3, N) array sv["F"]).  quad4: F_33 = 1. */
int fdk_gp_deformation_gradient(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                                const double* U, int fbar, double* F_gp, fdk_stream_t stream);

/* J2 plasticity (isotropic power-law hardening sigma_Y + k p^m), backward-Euler radial
 * return + consistent tangent, one thread per Gauss point.  Follows the Simcoon("EPICP")
 * state-variable protocol of fedoo/constitutivelaw/simcoon_umat.py:463-580 (props
 * [E, nu, alpha, sigmaY, k, m], statev [T, p, EP(6)] :103-113) and the algorithm sketch of
 * fedoo/constitutivelaw/elasto_plasticity.py:275-376.  All arrays column-major with gp-major
 * columns: strain_gp (6,N) total strain, statev_start (8,N) -> stress_gp (6,N), statev (8,N),
 * tangent_gp (6,6,N) or NULL. */
int fdk_j2_update(int64_t n_gp, const double* props_h, const double* strain_gp, const double* statev_start,
                  double* stress_gp, double* statev, double* tangent_gp, fdk_stream_t stream);

/* The same update returning the tangent in its STRUCTURED form: both J2 tangents (consistent / continuum) are
 * C = lam' 1(x)1 + 2 mu' I_sym - kappa n^(x)n^ with n^ the unit deviatoric stress direction; tangent_r1 holds, per
 * Gauss point (gp-major), the 10 doubles [lam', mu', kappa, n^ (6, stress Voigt), 0] -- 80 bytes instead of the 288 of
 * the (6,6,N) array of the reference's protocol (fedoo/constitutivelaw/simcoon_umat.py:556-580).
 * fdk_j2_tangent_expand rebuilds the (6,6,N) array (for sv["TangentMatrix"] readers and the other kernels);
 * fdk_assemble_elastic_r1 assembles K = int B^T C B straight from the structured form (hex8, 3D, plan built with
 * small = True; balanced kernel, 10 instead of 36 tangent operands per element and Gauss point in shared memory),
 * and D = -int B^T stress_gp when FDK_VECTOR is set. */
int fdk_j2_update_r1(int64_t n_gp, const double* props_h, const double* strain_gp, const double* statev_start,
                     double* stress_gp, double* statev, double* tangent_r1, fdk_stream_t stream);
int fdk_j2_tangent_expand(int64_t n_gp, const double* tangent_r1, double* tangent_gp, fdk_stream_t stream);
/* The update with the strain taken straight from the dof vector U (variable-major): geometry, grad u, Voigt strain and
 * radial return in one pass -- the strain array StressEquilibrium.update would write for the law to read back
 * (weakform/stress_equilibrium.py:191-217, constitutivelaw/simcoon_umat.py:561-576) never goes through HBM.  3D
 * elements; tangent_gp and / or tangent_r1 may be NULL. */
int fdk_j2_update_from_dofs(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                            const double* U, const double* props_h, const double* statev_start, double* stress_gp,
                            double* statev, double* tangent_gp, double* tangent_r1, fdk_stream_t stream);

int fdk_assemble_elastic_r1(const fdk_plan* plan, int compute, const double* coords, const double* tangent_r1,
                            const double* stress_gp, double* K_values, double* D, fdk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * What the callers of the assembly do with K, on the device.  Replaces the
 * host-side elimination + solve of Problem.solve (fedoo/core/problem.py:277-298: MatCB^T A MatCB,
 * B - A Xbc; fedoo/core/base.py:521-537: scipy cg with M = diag(1 / A.diagonal())) for pure
 * Dirichlet conditions.  CSR with float64 values; index_bytes = 4 (int32) or 8 (int64) for BOTH
 * indptr and indices (scipy's convention); free_mask: one byte per dof, 0 = imposed, or NULL.
 * ------------------------------------------------------------------------- */

/* y = A x restricted to the free dofs: y[r] = sum over free columns c of A[r,c] x[c] for free rows, 0 otherwise. */
int fdk_csr_spmv(int64_t n_rows, int64_t nnz, const void* indptr, const void* indices, int index_bytes,
                 const double* data, const double* x, const uint8_t* free_mask, double* y, fdk_stream_t stream);

/* The same product and solve for a matrix in the TILED pattern of fdk_sym_expand_csr without global dofs, given its
 * block pattern (blk_indptr int64 [n_nodes+1], blk_indices int32 [blk_nnz]): the column list of a block row is read
 * once for its nvar x nvar scalar rows (nvar = 1, 2, 3).  Results are identical up to summation order. */
int fdk_bcsr_spmv(int n_nodes, int nvar, int64_t blk_nnz, const int64_t* blk_indptr, const int32_t* blk_indices,
                  const double* data, const double* x, const uint8_t* free_mask, double* y, fdk_stream_t stream);
int fdk_bcsr_pcg_jacobi(int n_nodes, int nvar, int64_t blk_nnz, const int64_t* blk_indptr, const int32_t* blk_indices,
                        const void* indptr, const void* indices, int index_bytes, const double* data, const double* b,
                        double* x, const uint8_t* free_mask, double rtol, int max_iter, int check_every, double* work,
                        int* iters_h, double* relres_h, fdk_stream_t stream);

/* Periodic boundary conditions with global (non-nodal) dofs, on the device.  Replaces the MPC
 * elimination of Problem.solve (fedoo/core/problem.py:277-298, MatCB built by fedoo/core/boundary_conditions.py:560-700)
 * for the constraints PeriodicBC generates (fedoo/constraint/periodic_bc.py:910-1800):
 *     x[slave_s] = x[master_s] + sum_k coef[s][k] * x[n_nodal + k],   k < n_glob ("MeanStrain" dofs E_xx..E_yz)
 * Vectors carry n_nodal + n_glob entries.  The matrix keeps its nodal size: the global rows/columns that
 * Assembly.assemble_global_mat appends by resize (fedoo/core/assembly.py:192-197,459-460) are empty. */
typedef struct fdk_mpc {
  int64_t n_nodal;         /* nvar * n_nodes */
  int32_t n_glob;          /* trailing global dofs */
  int64_t n_slave;
  const int32_t* slave;    /* (n_slave) eliminated dof */
  const int32_t* master;   /* (n_slave) dof it follows; never itself a slave */
  const double* coef;      /* (n_slave, n_glob) row-major */
  int64_t n_master;        /* distinct masters */
  const int32_t* mst_dof;  /* (n_master) */
  const int32_t* mst_ptr;  /* (n_master + 1) into mst_slv */
  const int32_t* mst_slv;  /* slave ordinals grouped by master (fixed summation order) */
  double* scratch;         /* FDK_MPC_SCRATCH_DOUBLES doubles of device scratch (partial sums of the fold; one map per
                              stream at a time) */
} fdk_mpc;
#define FDK_MPC_SCRATCH_DOUBLES (9 * 9 * 148) /* n_glob <= 9, n_rhs <= 9 */

/* x <- T x: fills the slave entries from their masters and the global dofs. */
int fdk_mpc_expand(const fdk_mpc* mpc, double* x, fdk_stream_t stream);
/* q <- T^T q: slave rows folded into their masters and into the global rows (overwritten), slave entries cleared. */
int fdk_mpc_fold(const fdk_mpc* mpc, double* q, fdk_stream_t stream);
/* Jacobi-PCG on T^T A T: b already folded, free_mask (n_nodal + n_glob bytes) = 0 on imposed dofs AND on slaves,
 * x returns the independent dofs (call fdk_mpc_expand on it); work = fdk_pcg_work_doubles(n_nodal + n_glob). */
int fdk_bcsr_pcg_jacobi_mpc(int n_nodes, int nvar, int64_t blk_nnz, const int64_t* blk_indptr,
                            const int32_t* blk_indices, const void* indptr, const void* indices, int index_bytes,
                            const double* data, const double* b, double* x, const uint8_t* free_mask, double rtol,
                            int max_iter, int check_every, double* work, const fdk_mpc* mpc, int* iters_h,
                            double* relres_h, fdk_stream_t stream);

/* n_rhs systems sharing one matrix, solved in lockstep (the load cases of fedoo/homogen/tangent_stiffness.py:97-153):
 * b, x are (n_nodal + n_glob, n_rhs) row-major; every iteration reads K once for all products.  mpc may be NULL
 * (Dirichlet conditions only; free_mask then has n_nodal bytes).  b already folded; x returns the EXPANDED solutions;
 * relres_h has n_rhs entries; work = fdk_pcg_multi_work_doubles(n_nodal + n_glob, n_rhs).  n_rhs = 3 or 6. */
int64_t fdk_pcg_multi_work_doubles(int64_t n, int n_rhs);
int fdk_bcsr_pcg_jacobi_multi(int n_nodes, int nvar, int64_t blk_nnz, const int64_t* blk_indptr,
                              const int32_t* blk_indices, const void* indptr, const void* indices, int index_bytes,
                              const double* data, int n_rhs, const double* b, double* x, const uint8_t* free_mask,
                              double rtol, int max_iter, int check_every, double* work, const fdk_mpc* mpc, int* iters_h,
                              double* relres_h, fdk_stream_t stream);

/* diag[r] = A[r,r] (0 if not stored); columns sorted within a row (the pattern of fdk_sym_expand_csr is). */
int fdk_csr_diagonal(int64_t n_rows, const void* indptr, const void* indices, int index_bytes, const double* data,
                     double* diag, fdk_stream_t stream);

/* Jacobi-preconditioned conjugate gradient for A x = b on the free dofs, x = 0 initially and on the imposed
 * dofs; stops when ||r|| <= rtol ||b|| (checked every check_every iterations) or after max_iter iterations.
 * work: fdk_pcg_work_doubles(n) doubles of device scratch.  iters_h / relres_h (host, may be NULL) receive the
 * iteration count and ||r|| / ||b||.  Synchronises the stream (it reads the residual norm back). */
int64_t fdk_pcg_work_doubles(int64_t n);
int fdk_pcg_jacobi(int64_t n, int64_t nnz, const void* indptr, const void* indices, int index_bytes, const double* data,
                   const double* b, double* x, const uint8_t* free_mask, double rtol, int max_iter, int check_every,
                   double* work, int* iters_h, double* relres_h, fdk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Results extraction.  Replaces Mesh.convert_data GaussPoint -> Node / Element
 * (fedoo/core/mesh.py:1149-1160,1267-1308) and StressTensorList.von_mises (fedoo/util/voigt_tensors.py:270-283)
 * as used by Problem.get_results (fedoo/core/output.py:120-330).
 * field: value of component c at Gauss point n (gp-major, n = g * n_elems + e) at field[n * gp_stride + c * comp_stride]
 * (a (ncomp, N) column-major array has gp_stride = ncomp, comp_stride = 1; a row-major one gp_stride = 1,
 * comp_stride = N); ncomp <= 6.  von_mises != 0: the field has 6 Voigt stress components and the ONE converted
 * quantity is their von Mises norm, taken at the Gauss points first (output.py:190-197).
 * ------------------------------------------------------------------------- */

/* node value = (1 / #elements around the node) sum over those elements and their Gauss points of P[i][g] * value,
 * P_h = pinv(shape functions at the Gauss points), host, [nne][ngp] row-major.  node_ptr [n_nodes+1] / node_inc
 * (element * nne + local node) list the incidences of every node.  out: [ncomp_out][n_nodes] row-major. */
int fdk_gp_to_node(int nne, int ngp, int n_nodes, int64_t n_elems, const int64_t* node_ptr, const int32_t* node_inc,
                   const double* P_h, const double* field, int ncomp, int64_t comp_stride, int64_t gp_stride,
                   int von_mises, double* out, fdk_stream_t stream);

/* element value = mean over the element's Gauss points.  out: [ncomp_out][n_elems] row-major. */
int fdk_gp_to_element(int ngp, int64_t n_elems, const double* field, int ncomp, int64_t comp_stride, int64_t gp_stride,
                      int von_mises, double* out, fdk_stream_t stream);

/* out[n] = von Mises norm of the 6 Voigt stress components at Gauss point n. */
int fdk_gp_von_mises(int64_t n_gp, const double* field, int64_t comp_stride, int64_t gp_stride, double* out,
                     fdk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Multi-GPU helpers: pack / unpack-add of owned or halo entries around an NCCL
 * exchange of the global vector (no reference counterpart: the reference is
 * single-process).
 * ------------------------------------------------------------------------- */
int fdk_gather_f64(int64_t n, const int64_t* index, const double* src, double* dst, fdk_stream_t stream);
/* dst[seg_dst[s] + k] = src[seg_src[s] + k] for k < seg_len[s], s < n_seg (device arrays; max_len = max seg_len):
 * pack / unpack when the owned dofs of every rank are contiguous runs (slab partitions). */
int fdk_copy_segments(int n_seg, const int64_t* seg_src, const int64_t* seg_dst, const int64_t* seg_len,
                      int64_t max_len, const double* src, double* dst, fdk_stream_t stream);
int fdk_scatter_add_f64(int64_t n, const int64_t* index, const double* src, double* dst, fdk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FDK_H */
