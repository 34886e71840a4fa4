// fdk_api.cu -- the C ABI of libfdk (see include/fdk.h).  Single translation unit: the kernels
// live in the .cuh files included below.
#include <cstdlib>

#include "fdk_assemble.cuh"
#include "fdk_assemble_iso.cuh"
#include "fdk_color.cuh"
#include "fdk_gp.cuh"
#include "fdk_heat_tet4.cuh"
#include "fdk_results.cuh"
#include "fdk_rows.cuh"
#include "fdk_solve.cuh"
#include "fdk_symbolic.cuh"

using namespace fdk;

namespace {

// Runtime options (fdk_set_option); defaults from the environment at load time.
//   fuse_ku : take the residual of a linear law from the assembled rows (D = -K_row . U) when K and D are
//             both requested (default 1; FDK_NO_FUSE_KU=1 -> 0: always integrate B^T sigma)
//   mma     : hex8 + isotropic law: form the element matrices with FP64 tensor-core DMMA instead of
//             CUDA-core FMAs (default 1; FDK_MMA=0 -> 0).  Measured on B200: 19.7 vs 24.0 ms @ 8 M.
int env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return e == nullptr ? dflt : (e[0] == '1');
}
int g_opt_fuse = env_flag("FDK_NO_FUSE_KU", 0) ? 0 : 1;
int g_opt_mma = env_flag("FDK_MMA", 1);
//   iso4    : hex8 + isotropic law + matrix requested (residual fused): the balanced 1024-thread kernel of
//             fdk_assemble_iso.cuh (default 1; FDK_ISO4=0 -> 0: k_assemble, with or without mma)
int g_opt_iso4 = env_flag("FDK_ISO4", 1);
//   j2_continuum_tangent : fdk_j2_update returns the continuum elastoplastic tangent instead of the consistent one
int g_opt_j2_continuum = env_flag("FDK_J2_CONTINUUM_TANGENT", 0);

int check_plan(const fdk_plan* p) {
  FDK_REQUIRE(p != nullptr, FDK_EINVAL, "plan is NULL");
  int nne, ngp, dim;
  if (int rc = elem_dims(p->elem_type, &nne, &ngp, &dim)) return rc;
  FDK_REQUIRE(p->n_clusters >= 0 && p->n_nodes >= 0 && p->n_elems >= 0, FDK_EINVAL, "negative size in plan");
  if (p->n_clusters > 0)
    FDK_REQUIRE(p->cl_hdr && p->cl_node_ptr && p->cl_node && p->cl_bptr && p->cl_slot_ptr && p->cl_inc_ptr && p->inc_desc &&
                    p->cl_te_ptr && p->cl_te_elem && p->cl_lconn && p->cl_tn_ptr && p->cl_tn_node &&
                    p->cl_finc_ptr && p->ent_src && p->inc_fdst && p->slot_rec && p->cl_heavy_ptr && p->te_desc && p->cl_slot_loc && p->cl_finc_loc,
                FDK_EINVAL, "plan has NULL arrays");
  return 0;
}

int check_io(int compute, const double* coords, const double* K, const double* D) {
  FDK_REQUIRE(compute >= 1 && compute <= 3, FDK_EINVAL, "compute must be FDK_MATRIX, FDK_VECTOR or FDK_ALL");
  FDK_REQUIRE(coords != nullptr, FDK_EINVAL, "coords is NULL");
  FDK_REQUIRE(!(compute & FDK_MATRIX) || K != nullptr, FDK_EINVAL, "K_values is NULL but the matrix is requested");
  FDK_REQUIRE(!(compute & FDK_VECTOR) || D != nullptr, FDK_EINVAL, "D is NULL but the vector is requested");
  return 0;
}

}  // namespace

extern "C" {

const char* fdk_last_error_string(void) { return g_err; }
int fdk_version(void) { return 110; }

int fdk_set_option(const char* key, int value) {
  FDK_REQUIRE(key != nullptr, FDK_EINVAL, "option key is NULL");
  if (strcmp(key, "fuse_ku") == 0) { g_opt_fuse = value != 0; return 0; }
  if (strcmp(key, "mma") == 0) { g_opt_mma = value != 0; return 0; }
  if (strcmp(key, "iso4") == 0) { g_opt_iso4 = value != 0; return 0; }
  if (strcmp(key, "j2_continuum_tangent") == 0) { g_opt_j2_continuum = value != 0; return 0; }
  set_error("unknown option '%s' (known: fuse_ku, mma, iso4, j2_continuum_tangent)", key);
  return FDK_EINVAL;
}

int fdk_get_option(const char* key, int* value) {
  FDK_REQUIRE(key != nullptr && value != nullptr, FDK_EINVAL, "NULL argument");
  if (strcmp(key, "fuse_ku") == 0) { *value = g_opt_fuse; return 0; }
  if (strcmp(key, "mma") == 0) { *value = g_opt_mma; return 0; }
  if (strcmp(key, "iso4") == 0) { *value = g_opt_iso4; return 0; }
  if (strcmp(key, "j2_continuum_tangent") == 0) { *value = g_opt_j2_continuum; return 0; }
  set_error("unknown option '%s' (known: fuse_ku, mma, iso4, j2_continuum_tangent)", key);
  return FDK_EINVAL;
}

int fdk_debug_phase_clocks(unsigned long long* out_h, int n, int reset) {
#ifdef FDK_PHASE_CLOCKS
  FDK_REQUIRE(out_h != nullptr && n >= 0 && n <= 16, FDK_EINVAL, "bad arguments");
  unsigned long long tmp[16];
  FDK_CUDA(cudaDeviceSynchronize());
  FDK_CUDA(cudaMemcpyFromSymbol(tmp, g_phase_clk, sizeof(tmp)));
  for (int i = 0; i < n; ++i) out_h[i] = tmp[i];
  if (reset) {
    for (int i = 0; i < 16; ++i) tmp[i] = 0;
    FDK_CUDA(cudaMemcpyToSymbol(g_phase_clk, tmp, sizeof(tmp)));
  }
  return 0;
#else
  (void)out_h; (void)n; (void)reset;
  set_error("libfdk was built without -DFDK_PHASE_CLOCKS");
  return FDK_EINVAL;
#endif
}

int fdk_element_info(int elem_type, int* nne, int* ngp, int* dim) { return elem_dims(elem_type, nne, ngp, dim); }

int fdk_element_table(int elem_type, double* w, double* N, double* dN) {
  int nne, ngp, dim;
  if (int rc = elem_dims(elem_type, &nne, &ngp, &dim)) return rc;
  const ElemTable& t = host_table(elem_type);
  for (int g = 0; g < ngp; ++g) w[g] = t.w[g];
  for (int i = 0; i < ngp * nne; ++i) N[i] = t.N[i];
  for (int i = 0; i < ngp * dim * nne; ++i) dN[i] = t.dN[i];
  return 0;
}

int fdk_sym_block_keys(int n_nodes, int64_t n_elems, int nne, const int32_t* conn, uint64_t* keys_out,
                       int64_t* blk_nnz_h, fdk_stream_t stream) {
  FDK_REQUIRE(n_nodes >= 0 && n_elems >= 0 && nne > 0 && nne <= MAX_NNE, FDK_EINVAL, "bad sizes");
  FDK_REQUIRE(blk_nnz_h != nullptr, FDK_EINVAL, "blk_nnz_h is NULL");
  FDK_REQUIRE(n_elems == 0 || (conn && keys_out), FDK_EINVAL, "NULL buffer");
  return sym_block_keys(n_nodes, n_elems, nne, conn, keys_out, blk_nnz_h, (cudaStream_t)stream);
}

int fdk_sym_block_csr(int n_nodes, int64_t blk_nnz, const uint64_t* keys, int64_t* blk_indptr, int32_t* blk_indices,
                      fdk_stream_t stream) {
  FDK_REQUIRE(n_nodes >= 0 && blk_nnz >= 0 && blk_indptr, FDK_EINVAL, "bad arguments");
  return sym_block_csr(n_nodes, blk_nnz, keys, blk_indptr, blk_indices, (cudaStream_t)stream);
}

int fdk_sym_expand_csr(int n_nodes, int nvar, int n_global_dof, int64_t blk_nnz, const int64_t* blk_indptr,
                       const int32_t* blk_indices, int index_bytes, void* indptr, void* indices, fdk_stream_t stream) {
  FDK_REQUIRE(n_nodes >= 0 && nvar > 0 && n_global_dof >= 0 && indptr, FDK_EINVAL, "bad arguments");
  return sym_expand_csr(n_nodes, nvar, n_global_dof, blk_nnz, blk_indptr, blk_indices, index_bytes, indptr, indices,
                        (cudaStream_t)stream);
}

int fdk_plan_color_blocks(const fdk_plan* plan, uint8_t* blk_slot, uint16_t* ent_pos, fdk_stream_t stream) {
  if (int rc = check_plan(plan)) return rc;
  FDK_REQUIRE(plan->elem_type == FDK_HEX8, FDK_EINVAL, "block colouring is defined for hex8 plans");
  FDK_REQUIRE(plan->cap_inc * 8 <= COLOR_MAX_BLOCKS, FDK_ECAP, "cluster with %d incidences exceeds the colouring capacity",
              plan->cap_inc);
  if (plan->n_clusters == 0) return 0;
  FDK_REQUIRE(blk_slot && ent_pos, FDK_EINVAL, "NULL output");
  k_color_blocks<<<plan->n_clusters, 32, 0, (cudaStream_t)stream>>>(*plan, blk_slot, ent_pos);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

int fdk_assemble_elastic_iso(const fdk_plan* plan, int compute, const double* coords, double lambda, double mu,
                             const double* U, const double* stress_gp, double* K_values, double* D,
                             fdk_stream_t stream) {
  if (int rc = check_plan(plan)) return rc;
  if (int rc = check_io(compute, coords, K_values, D)) return rc;
  FDK_REQUIRE(!(compute & FDK_VECTOR) || U || stress_gp, FDK_EINVAL, "the vector needs U or stress_gp");
  AsmArgs a{};
  a.p = *plan;
  a.coords = coords;
  a.U = U;
  a.stress_gp = stress_gp;
  a.K = K_values;
  a.D = D;
  a.lam = lambda;
  a.mu = mu;
  a.compute = compute;
  // linear law and both outputs requested: the residual -int B^T C eps(U) equals -(K U) row by row,
  // so it is taken from the assembled rows in the gather phase instead of a second B^T sigma pass
  a.fuse_ku = (compute == FDK_ALL && U != nullptr && stress_gp == nullptr && g_opt_fuse) ? 1 : 0;
  a.no_mma = g_opt_mma ? 0 : 1;
  // headline case: matrix requested and no B^T sigma pass -> balanced kernel (fdk_assemble_iso.cuh)
  const bool bts = (compute & FDK_VECTOR) && !a.fuse_ku;
  // (the template also compiles and passes parity for tet4 / tet10 / quad4 and for 16-node hex8 clusters, but was
  // measured slower than k_assemble on tet10 -- 64 vs 55 ms at 5 M elements -- and is unmeasured on the others)
  if (g_opt_iso4 && (compute & FDK_MATRIX) && !bts && plan->elem_type == FDK_HEX8 && plan->threads == Hex8::THREADS &&
      plan->blk_slot && plan->ent_pos && assemble_iso_fits<Hex8, 1024, 4>(a))
    return launch_assemble_iso<Hex8, 1024, 4>(a, (cudaStream_t)stream);
  // tet10 plans built for 1024-thread CTAs (fedoo_b200/plan.py, _CAPS_BIG): the balanced organisation with 5 threads
  // per incidence (two column blocks each, as for hex8) -- 32 warps per SM instead of the generic kernel's 8
  if (g_opt_iso4 && (compute & FDK_MATRIX) && !bts && plan->elem_type == FDK_TET10 && plan->threads == 1024 &&
      assemble_iso_fits<Tet10, 1024, 5>(a))
    return launch_assemble_iso<Tet10, 1024, 5>(a, (cudaStream_t)stream);
  if (g_opt_iso4 && (compute & FDK_MATRIX) && !bts && plan->elem_type == FDK_TET10 && plan->threads == 768 &&
      assemble_iso_fits<Tet10, 768, 5>(a))
    return launch_assemble_iso<Tet10, 768, 5>(a, (cudaStream_t)stream);
  FDK_REQUIRE(plan->threads != 1024 && plan->threads != 768, FDK_ECAP,
              "plan built for %d-thread clusters does not fit the balanced kernel", plan->threads);
  return dispatch_assemble<PHYS_ISO>(a, (cudaStream_t)stream);
}

int fdk_assemble_elastic_iso_dist(const fdk_plan* plan, int compute, const double* coords, double lambda, double mu,
                                  const double* U, double* K_values, double* D, double* const* D_dst_h, int n_dst,
                                  const int64_t* node_gid, int64_t n_global_nodes, fdk_stream_t stream) {
  if (int rc = check_plan(plan)) return rc;
  if (int rc = check_io(compute, coords, K_values, D)) return rc;
  FDK_REQUIRE(compute == FDK_ALL && U != nullptr, FDK_EINVAL, "the fused exchange needs compute = all and U");
  FDK_REQUIRE(n_dst >= 1 && n_dst <= 8 && D_dst_h && node_gid && n_global_nodes > 0, FDK_EINVAL, "bad destinations");
  AsmArgs a{};
  a.p = *plan;
  a.coords = coords;
  a.U = U;
  a.K = K_values;
  a.D = D;
  a.lam = lambda;
  a.mu = mu;
  a.compute = compute;
  a.fuse_ku = 1;
  for (int r = 0; r < n_dst; ++r) {
    FDK_REQUIRE(D_dst_h[r] != nullptr, FDK_EINVAL, "NULL destination");
    a.D_dst[r] = D_dst_h[r];
  }
  a.n_dst = n_dst;
  a.node_gid = node_gid;
  a.n_dst_nodes = n_global_nodes;
  const bool served = plan->elem_type == FDK_HEX8 && plan->threads == Hex8::THREADS && plan->blk_slot &&
                      plan->ent_pos && (assemble_iso_fits<Hex8, 1024, 4>(a));
  FDK_REQUIRE(served, FDK_EINVAL, "the fused exchange is served by the balanced hex8 kernel (32-node clusters) only");
  return launch_assemble_iso<Hex8, 1024, 4, PHYS_ISO, true>(a, (cudaStream_t)stream);
}

int fdk_assemble_elastic_general(const fdk_plan* plan, int compute, const double* coords, const double* C_h,
                                 const double* tangent_gp, const double* U, const double* stress_gp,
                                 double* K_values, double* D, fdk_stream_t stream) {
  if (int rc = check_plan(plan)) return rc;
  if (int rc = check_io(compute, coords, K_values, D)) return rc;
  FDK_REQUIRE(C_h || tangent_gp, FDK_EINVAL, "need C_h or tangent_gp");
  FDK_REQUIRE(!(compute & FDK_VECTOR) || U || stress_gp, FDK_EINVAL, "the vector needs U or stress_gp");
  AsmArgs a{};
  a.p = *plan;
  a.coords = coords;
  a.U = U;
  a.stress_gp = stress_gp;
  a.tangent_gp = tangent_gp;
  a.K = K_values;
  a.D = D;
  if (C_h)
    for (int i = 0; i < 36; ++i) a.C[i] = C_h[i];
  a.compute = compute;
  a.fuse_ku =
      (compute == FDK_ALL && U != nullptr && stress_gp == nullptr && tangent_gp == nullptr && g_opt_fuse) ? 1 : 0;
  // 3D hex8 on 16-node clusters (plans built with small = True), matrix requested, residual fused or integrated from
  // a given stress: balanced kernel with the tangent staged in shared memory (fdk_assemble_iso.cuh)
  const bool vec_ok = !(compute & FDK_VECTOR) || a.fuse_ku || stress_gp != nullptr;
  if (g_opt_iso4 && (compute & FDK_MATRIX) && vec_ok && plan->elem_type == FDK_HEX8 &&
      plan->threads == Hex8::THREADS / 2 && plan->blk_slot && plan->ent_pos &&
      assemble_iso_fits<Hex8, 512, 4, PHYS_GENERAL>(a))
    return launch_assemble_iso<Hex8, 512, 4, PHYS_GENERAL>(a, (cudaStream_t)stream);
  return dispatch_assemble<PHYS_GENERAL>(a, (cudaStream_t)stream);
}

int fdk_assemble_rows_elastic(int elem_type, int n_rows, const int32_t* rows, int n_nodes, int64_t n_elems,
                              const int32_t* conn, const double* coords, const int64_t* node_ptr, const int32_t* node_inc,
                              const int64_t* blk_indptr, const int32_t* blk_indices, int64_t blk_nnz, int max_row_degree,
                              int isotropic, double lam, double mu, const double* C_h, const double* tangent_gp,
                              int compute, const double* U, const double* stress_gp, double* K_values, double* D,
                              fdk_stream_t stream) {
  if (n_rows == 0) return 0;
  FDK_REQUIRE(rows && conn && coords && node_ptr && node_inc && blk_indptr && blk_indices, FDK_EINVAL, "NULL argument");
  FDK_REQUIRE((compute & ~FDK_ALL) == 0 && compute != 0, FDK_EINVAL, "compute must be FDK_MATRIX, FDK_VECTOR or both");
  FDK_REQUIRE(!(compute & FDK_MATRIX) || K_values, FDK_EINVAL, "matrix requested without K_values");
  FDK_REQUIRE(!(compute & FDK_VECTOR) || (D && (U || stress_gp)), FDK_EINVAL, "the vector needs D and U or stress_gp");
  FDK_REQUIRE(isotropic || C_h || tangent_gp, FDK_EINVAL, "need an isotropic law, C_h or tangent_gp");
  RowArgs a{};
  a.n_rows = n_rows;
  a.rows = rows;
  a.n_nodes = n_nodes;
  a.n_elems = n_elems;
  a.conn = conn;
  a.coords = coords;
  a.node_ptr = node_ptr;
  a.node_inc = node_inc;
  a.blk_indptr = blk_indptr;
  a.blk_indices = blk_indices;
  a.blk_nnz = blk_nnz;
  a.max_deg = max_row_degree;
  a.lam = lam;
  a.mu = mu;
  if (C_h)
    for (int i = 0; i < 36; ++i) a.C[i] = C_h[i];
  a.tangent_gp = tangent_gp;
  a.U = U;
  a.stress_gp = stress_gp;
  a.compute = compute;
  a.fuse_ku = ((compute & FDK_VECTOR) && stress_gp == nullptr) ? 1 : 0;  // linear in U: D_I = -K_row . U exactly
  a.K = K_values;
  a.D = D;
  cudaStream_t st = (cudaStream_t)stream;
#define FDK_ROWS(El) (isotropic ? launch_assemble_rows<El, PHYS_ISO>(a, st) : launch_assemble_rows<El, PHYS_GENERAL>(a, st))
  switch (elem_type) {
    case FDK_HEX8: return FDK_ROWS(Hex8);
    case FDK_TET4: return FDK_ROWS(Tet4);
    case FDK_TET10: return FDK_ROWS(Tet10);
    case FDK_QUAD4: return FDK_ROWS(Quad4);
  }
#undef FDK_ROWS
  set_error("unknown element type %d", elem_type);
  return FDK_EINVAL;
}

int fdk_assemble_elastic_r1(const fdk_plan* plan, int compute, const double* coords, const double* tangent_r1,
                            const double* stress_gp, double* K_values, double* D, fdk_stream_t stream) {
  if (int rc = check_plan(plan)) return rc;
  if (int rc = check_io(compute, coords, K_values, D)) return rc;
  FDK_REQUIRE(tangent_r1 != nullptr, FDK_EINVAL, "tangent_r1 is NULL");
  FDK_REQUIRE(compute & FDK_MATRIX, FDK_EINVAL, "the structured-tangent kernel assembles the matrix (residual alone: fdk_residual_elastic)");
  FDK_REQUIRE(!(compute & FDK_VECTOR) || stress_gp, FDK_EINVAL, "the vector needs stress_gp");
  FDK_REQUIRE(plan->elem_type == FDK_HEX8 && plan->nvar == 3 && plan->threads == Hex8::THREADS / 2 && plan->blk_slot &&
                  plan->ent_pos,
              FDK_EINVAL, "structured J2 tangent: hex8, 3 dofs per node, plan built with small = True");
  AsmArgs a{};
  a.p = *plan;
  a.coords = coords;
  a.stress_gp = stress_gp;
  a.tangent_r1 = tangent_r1;
  a.K = K_values;
  a.D = D;
  a.compute = compute;
  a.fuse_ku = 0;
  FDK_REQUIRE((assemble_iso_fits<Hex8, 512, 4, PHYS_R1>(a)), FDK_ECAP, "plan does not fit the balanced kernel");
  return launch_assemble_iso<Hex8, 512, 4, PHYS_R1>(a, (cudaStream_t)stream);
}

int fdk_j2_update_r1(int64_t n_gp, const double* props_h, const double* strain_gp, const double* statev_start,
                     double* stress_gp, double* statev, double* tangent_r1, fdk_stream_t stream) {
  FDK_REQUIRE(props_h && strain_gp && statev_start && stress_gp && statev && tangent_r1, FDK_EINVAL, "NULL argument");
  FDK_REQUIRE(n_gp >= 0, FDK_EINVAL, "negative n_gp");
  if (n_gp == 0) return 0;
  J2Args a{};
  a.n_gp = n_gp;
  a.E = props_h[0];
  a.nu = props_h[1];
  a.sigY = props_h[3];
  a.k = props_h[4];
  a.m = props_h[5];
  a.strain = strain_gp;
  a.statev0 = statev_start;
  a.stress = stress_gp;
  a.statev = statev;
  a.tangent = nullptr;
  a.tangent_r1 = tangent_r1;
  a.continuum = g_opt_j2_continuum;
  k_j2_update<<<(unsigned)((n_gp + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

int fdk_j2_update_from_dofs(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                            const double* U, const double* props_h, const double* statev_start, double* stress_gp,
                            double* statev, double* tangent_gp, double* tangent_r1, fdk_stream_t stream) {
  FDK_REQUIRE(conn && coords && U && props_h && statev_start && stress_gp && statev, FDK_EINVAL, "NULL argument");
  FDK_REQUIRE(n_elems >= 0, FDK_EINVAL, "negative n_elems");
  if (n_elems == 0) return 0;
  J2UArgs a{};
  a.n_nodes = n_nodes;
  a.n_elems = n_elems;
  a.conn = conn;
  a.coords = coords;
  a.U = U;
  a.j2.E = props_h[0];
  a.j2.nu = props_h[1];
  a.j2.sigY = props_h[3];
  a.j2.k = props_h[4];
  a.j2.m = props_h[5];
  a.j2.statev0 = statev_start;
  a.j2.stress = stress_gp;
  a.j2.statev = statev;
  a.j2.tangent = tangent_gp;
  a.j2.tangent_r1 = tangent_r1;
  a.j2.continuum = g_opt_j2_continuum;
  switch (elem_type) {
    case FDK_HEX8: a.j2.n_gp = n_elems * Hex8::NGP; return launch_j2_update_u<Hex8>(a, (cudaStream_t)stream);
    case FDK_TET4: a.j2.n_gp = n_elems * Tet4::NGP; return launch_j2_update_u<Tet4>(a, (cudaStream_t)stream);
    case FDK_TET10: a.j2.n_gp = n_elems * Tet10::NGP; return launch_j2_update_u<Tet10>(a, (cudaStream_t)stream);
  }
  set_error("fdk_j2_update_from_dofs: 3D element types only (got %d)", elem_type);
  return FDK_EINVAL;
}

int fdk_j2_tangent_expand(int64_t n_gp, const double* tangent_r1, double* tangent_gp, fdk_stream_t stream) {
  FDK_REQUIRE(n_gp >= 0 && (n_gp == 0 || (tangent_r1 && tangent_gp)), FDK_EINVAL, "bad argument");
  if (n_gp == 0) return 0;
  k_j2_tangent_expand<<<(unsigned)((n_gp + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_gp, tangent_r1, tangent_gp);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

int fdk_assemble_heat(const fdk_plan* plan, int compute, const double* coords, const double* cond_h,
                      double rho_c_over_dt, const double* T, const double* T_start, double* K_values, double* D,
                      fdk_stream_t stream) {
  if (int rc = check_plan(plan)) return rc;
  if (int rc = check_io(compute, coords, K_values, D)) return rc;
  FDK_REQUIRE(cond_h != nullptr, FDK_EINVAL, "cond_h is NULL");
  FDK_REQUIRE(!(compute & FDK_VECTOR) || T, FDK_EINVAL, "the vector needs T");
  AsmArgs a{};
  a.p = *plan;
  a.coords = coords;
  a.U = T;
  a.U2 = T_start;
  a.K = K_values;
  a.D = D;
  for (int i = 0; i < 9; ++i) a.cond[i] = cond_h[i];
  a.rcdt = rho_c_over_dt;
  a.compute = compute;
  return dispatch_assemble<PHYS_HEAT>(a, (cudaStream_t)stream);
}

static int gp_strain_stress_impl(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                                 const double* U, const double* C_h, const double* tangent_gp, double* fbar_center,
                                 double* grad_gp, double* strain_gp, double* stress_gp, fdk_stream_t stream) {
  FDK_REQUIRE(conn && coords && U, FDK_EINVAL, "NULL input");
  FDK_REQUIRE(!stress_gp || C_h || tangent_gp, FDK_EINVAL, "stress needs C_h or tangent_gp");
  GpArgs a{};
  a.n_nodes = n_nodes;
  a.n_elems = n_elems;
  a.conn = conn;
  a.coords = coords;
  a.U = U;
  a.tangent_gp = tangent_gp;
  a.grad_gp = grad_gp;
  a.strain_gp = strain_gp;
  a.stress_gp = stress_gp;
  a.fbar_center = fbar_center;
  if (C_h)
    for (int i = 0; i < 36; ++i) a.C[i] = C_h[i];
  if (fbar_center != nullptr) {
    FDK_REQUIRE(elem_type != FDK_QUAD4, FDK_EINVAL, "F-bar is available for the 3-D elements");
    int rc = FDK_EINVAL;
    switch (elem_type) {
      case FDK_HEX8: rc = launch_gp_fbar_center<Hex8>(a, (cudaStream_t)stream); break;
      case FDK_TET4: rc = launch_gp_fbar_center<Tet4>(a, (cudaStream_t)stream); break;
      case FDK_TET10: rc = launch_gp_fbar_center<Tet10>(a, (cudaStream_t)stream); break;
    }
    if (rc) return rc;
  }
  switch (elem_type) {
    case FDK_HEX8: return launch_gp_strain_stress<Hex8>(a, (cudaStream_t)stream);
    case FDK_TET4: return launch_gp_strain_stress<Tet4>(a, (cudaStream_t)stream);
    case FDK_TET10: return launch_gp_strain_stress<Tet10>(a, (cudaStream_t)stream);
    case FDK_QUAD4: return launch_gp_strain_stress<Quad4>(a, (cudaStream_t)stream);
  }
  set_error("unknown element type %d", elem_type);
  return FDK_EINVAL;
}

int fdk_gp_strain_stress(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                         const double* U, const double* C_h, const double* tangent_gp, double* grad_gp,
                         double* strain_gp, double* stress_gp, fdk_stream_t stream) {
  return gp_strain_stress_impl(elem_type, n_nodes, n_elems, conn, coords, U, C_h, tangent_gp, nullptr, grad_gp, strain_gp,
                               stress_gp, stream);
}

int fdk_gp_strain_stress_fbar(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                              const double* U, const double* C_h, const double* tangent_gp, double* fbar_center,
                              double* grad_gp, double* strain_gp, double* stress_gp, fdk_stream_t stream) {
  FDK_REQUIRE(fbar_center != nullptr, FDK_EINVAL, "NULL scratch for the element means");
  return gp_strain_stress_impl(elem_type, n_nodes, n_elems, conn, coords, U, C_h, tangent_gp, fbar_center, grad_gp,
                               strain_gp, stress_gp, stream);
}

int fdk_residual_elastic(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                         const double* C_h, const double* tangent_gp, const double* U, const double* stress_gp,
                         const int64_t* node_ptr, const int32_t* node_inc, double* fe_scratch, double* D,
                         fdk_stream_t stream) {
  FDK_REQUIRE(conn && coords && node_ptr && node_inc && fe_scratch && D, FDK_EINVAL, "NULL argument");
  FDK_REQUIRE(stress_gp || (U && (C_h || tangent_gp)), FDK_EINVAL, "need stress_gp, or U with C_h or tangent_gp");
  ResArgs a{};
  a.n_nodes = n_nodes;
  a.n_elems = n_elems;
  a.conn = conn;
  a.coords = coords;
  a.U = U;
  a.stress_gp = stress_gp;
  a.tangent_gp = tangent_gp;
  a.fe = fe_scratch;
  if (C_h)
    for (int i = 0; i < 36; ++i) a.C[i] = C_h[i];
  switch (elem_type) {
    case FDK_HEX8: return launch_residual<Hex8>(a, node_ptr, node_inc, D, (cudaStream_t)stream);
    case FDK_TET4: return launch_residual<Tet4>(a, node_ptr, node_inc, D, (cudaStream_t)stream);
    case FDK_TET10: return launch_residual<Tet10>(a, node_ptr, node_inc, D, (cudaStream_t)stream);
    case FDK_QUAD4: return launch_residual<Quad4>(a, node_ptr, node_inc, D, (cudaStream_t)stream);
  }
  set_error("unknown element type %d", elem_type);
  return FDK_EINVAL;
}

int fdk_residual_heat(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                      const double* cond_h, double rho_c_over_dt, const double* T, const double* T_start,
                      const int64_t* node_ptr, const int32_t* node_inc, double* fe_scratch, double* D,
                      fdk_stream_t stream) {
  FDK_REQUIRE(conn && coords && cond_h && T && node_ptr && node_inc && fe_scratch && D, FDK_EINVAL, "NULL argument");
  ResHeatArgs a{};
  a.n_nodes = n_nodes;
  a.n_elems = n_elems;
  a.conn = conn;
  a.coords = coords;
  a.T = T;
  a.T_start = T_start;
  a.rcdt = rho_c_over_dt;
  a.fe = fe_scratch;
  for (int i = 0; i < 9; ++i) a.cond[i] = cond_h[i];
  switch (elem_type) {
    case FDK_HEX8: return launch_residual_heat<Hex8>(a, node_ptr, node_inc, D, (cudaStream_t)stream);
    case FDK_TET4: return launch_residual_heat<Tet4>(a, node_ptr, node_inc, D, (cudaStream_t)stream);
    case FDK_TET10: return launch_residual_heat<Tet10>(a, node_ptr, node_inc, D, (cudaStream_t)stream);
    case FDK_QUAD4: return launch_residual_heat<Quad4>(a, node_ptr, node_inc, D, (cudaStream_t)stream);
  }
  set_error("unknown element type %d", elem_type);
  return FDK_EINVAL;
}

int fdk_assemble_heat_tet4(int compute, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                           const double* cond_h, double rho_c_over_dt, const double* T, const double* T_start,
                           const int64_t* node_ptr, const int32_t* inc_rec, const int64_t* blk_indptr,
                           int max_row_degree, int n_rows, const int32_t* rows, double* K_values, double* D,
                           fdk_stream_t stream) {
  FDK_REQUIRE(conn && coords && cond_h && node_ptr && inc_rec && blk_indptr, FDK_EINVAL, "NULL argument");
  FDK_REQUIRE(rows ? (n_rows >= 0 && n_rows <= n_nodes) : (n_rows == n_nodes || n_rows == 0), FDK_EINVAL,
              "n_rows: the length of rows, or n_nodes (0 is accepted too) when rows is NULL");
  FDK_REQUIRE((compute & ~FDK_ALL) == 0 && compute != 0, FDK_EINVAL, "compute must be FDK_MATRIX, FDK_VECTOR or both");
  FDK_REQUIRE(!(compute & FDK_MATRIX) || K_values, FDK_EINVAL, "matrix requested without K_values");
  FDK_REQUIRE(!(compute & FDK_VECTOR) || (D && T), FDK_EINVAL, "vector requested without D / T");
  HeatTet4Args a{};
  a.n_nodes = n_nodes;
  a.n_elems = n_elems;
  a.conn = conn;
  a.coords = coords;
  a.node_ptr = node_ptr;
  a.inc_rec = reinterpret_cast<const int2*>(inc_rec);
  a.blk_indptr = blk_indptr;
  a.T = (compute & FDK_VECTOR) ? T : nullptr;
  a.T_start = T_start;
  a.rcdt = rho_c_over_dt;
  a.max_deg = max_row_degree;
  a.rows = rows;
  a.n_rows = rows ? n_rows : n_nodes;
  a.K = (compute & FDK_MATRIX) ? K_values : nullptr;
  a.D = (compute & FDK_VECTOR) ? D : nullptr;
  for (int i = 0; i < 9; ++i) a.cond[i] = cond_h[i];
  return launch_heat_tet4(a, (cudaStream_t)stream);
}

int fdk_gp_deformation_gradient(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                                const double* U, int fbar, double* F_gp, fdk_stream_t stream) {
  FDK_REQUIRE(conn && coords && U && F_gp, FDK_EINVAL, "NULL argument");
  DefGradArgs a{};
  a.n_nodes = n_nodes;
  a.n_elems = n_elems;
  a.conn = conn;
  a.coords = coords;
  a.U = U;
  a.fbar = fbar;
  a.F_gp = F_gp;
  switch (elem_type) {
    case FDK_HEX8: return launch_gp_defgrad<Hex8>(a, (cudaStream_t)stream);
    case FDK_TET4: return launch_gp_defgrad<Tet4>(a, (cudaStream_t)stream);
    case FDK_TET10: return launch_gp_defgrad<Tet10>(a, (cudaStream_t)stream);
    case FDK_QUAD4: return launch_gp_defgrad<Quad4>(a, (cudaStream_t)stream);
  }
  set_error("unknown element type %d", elem_type);
  return FDK_EINVAL;
}

int fdk_residual_heat_gp(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                         const double* flux_gp, const double* src_gp, const int64_t* node_ptr, const int32_t* node_inc,
                         double* fe_scratch, double* D, fdk_stream_t stream) {
  FDK_REQUIRE(conn && coords && node_ptr && node_inc && fe_scratch && D, FDK_EINVAL, "NULL argument");
  ResHeatGpArgs a{};
  a.n_nodes = n_nodes;
  a.n_elems = n_elems;
  a.conn = conn;
  a.coords = coords;
  a.flux_gp = flux_gp;
  a.src_gp = src_gp;
  a.fe = fe_scratch;
  switch (elem_type) {
    case FDK_HEX8: return launch_residual_heat_gp<Hex8>(a, node_ptr, node_inc, D, (cudaStream_t)stream);
    case FDK_TET4: return launch_residual_heat_gp<Tet4>(a, node_ptr, node_inc, D, (cudaStream_t)stream);
    case FDK_TET10: return launch_residual_heat_gp<Tet10>(a, node_ptr, node_inc, D, (cudaStream_t)stream);
    case FDK_QUAD4: return launch_residual_heat_gp<Quad4>(a, node_ptr, node_inc, D, (cudaStream_t)stream);
  }
  set_error("unknown element type %d", elem_type);
  return FDK_EINVAL;
}

int fdk_gp_temperature(int elem_type, int n_nodes, int64_t n_elems, const int32_t* conn, const double* coords,
                       const double* T, double* temp_gp, double* temp_gradient_gp, fdk_stream_t stream) {
  FDK_REQUIRE(conn && coords && T, FDK_EINVAL, "NULL input");
  GpArgs a{};
  a.n_nodes = n_nodes;
  a.n_elems = n_elems;
  a.conn = conn;
  a.coords = coords;
  a.U = T;
  a.temp_gp = temp_gp;
  a.temp_grad_gp = temp_gradient_gp;
  switch (elem_type) {
    case FDK_HEX8: return launch_gp_temperature<Hex8>(a, (cudaStream_t)stream);
    case FDK_TET4: return launch_gp_temperature<Tet4>(a, (cudaStream_t)stream);
    case FDK_TET10: return launch_gp_temperature<Tet10>(a, (cudaStream_t)stream);
    case FDK_QUAD4: return launch_gp_temperature<Quad4>(a, (cudaStream_t)stream);
  }
  set_error("unknown element type %d", elem_type);
  return FDK_EINVAL;
}

int fdk_j2_update(int64_t n_gp, const double* props_h, const double* strain_gp, const double* statev_start,
                  double* stress_gp, double* statev, double* tangent_gp, fdk_stream_t stream) {
  FDK_REQUIRE(props_h && strain_gp && statev_start && stress_gp && statev, FDK_EINVAL, "NULL argument");
  FDK_REQUIRE(n_gp >= 0, FDK_EINVAL, "negative n_gp");
  if (n_gp == 0) return 0;
  J2Args a{};
  a.n_gp = n_gp;
  a.E = props_h[0];
  a.nu = props_h[1];
  a.sigY = props_h[3];
  a.k = props_h[4];
  a.m = props_h[5];
  a.strain = strain_gp;
  a.statev0 = statev_start;
  a.stress = stress_gp;
  a.statev = statev;
  a.tangent = tangent_gp;
  a.continuum = g_opt_j2_continuum;
  k_j2_update<<<(unsigned)((n_gp + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

int fdk_csr_spmv(int64_t n_rows, int64_t nnz, const void* indptr, const void* indices, int index_bytes,
                 const double* data, const double* x, const uint8_t* free_mask, double* y, fdk_stream_t stream) {
  FDK_REQUIRE(n_rows >= 0 && nnz >= 0, FDK_EINVAL, "negative size");
  if (n_rows == 0) return 0;
  FDK_REQUIRE(indptr && x && y && (nnz == 0 || (indices && data)), FDK_EINVAL, "NULL argument");
  FDK_REQUIRE(index_bytes == 4 || index_bytes == 8, FDK_EINVAL, "index_bytes must be 4 or 8");
  const int lanes = pick_lanes(n_rows, nnz);
  if (index_bytes == 4)
    return launch_spmv<int32_t, true>(n_rows, (const int32_t*)indptr, (const int32_t*)indices, data, x, free_mask, y, lanes,
                                (cudaStream_t)stream);
  return launch_spmv<int64_t, true>(n_rows, (const int64_t*)indptr, (const int64_t*)indices, data, x, free_mask, y, lanes,
                              (cudaStream_t)stream);
}

int fdk_bcsr_spmv(int n_nodes, int nvar, int64_t blk_nnz, const int64_t* blk_indptr, const int32_t* blk_indices,
                  const double* data, const double* x, const uint8_t* free_mask, double* y, fdk_stream_t stream) {
  FDK_REQUIRE(n_nodes >= 0 && blk_nnz >= 0, FDK_EINVAL, "negative size");
  if (n_nodes == 0) return 0;
  FDK_REQUIRE(blk_indptr && blk_indices && data && x && y, FDK_EINVAL, "NULL argument");
  BlockPattern b;
  b.n_nodes = n_nodes; b.nvar = nvar; b.blk_nnz = blk_nnz; b.blk_indptr = blk_indptr; b.blk_indices = blk_indices;
  return launch_bspmv<true>(b, data, x, free_mask, y, (cudaStream_t)stream);
}

int fdk_bcsr_pcg_jacobi(int n_nodes, int nvar, int64_t blk_nnz, const int64_t* blk_indptr, const int32_t* blk_indices,
                        const void* indptr, const void* indices, int index_bytes, const double* data, const double* b,
                        double* x, const uint8_t* free_mask, double rtol, int max_iter, int check_every, double* work,
                        int* iters_h, double* relres_h, fdk_stream_t stream) {
  FDK_REQUIRE(n_nodes >= 0 && blk_nnz >= 0 && nvar >= 1 && nvar <= 3 && max_iter >= 0 && rtol >= 0.0, FDK_EINVAL,
              "bad size or tolerance");
  if (iters_h) *iters_h = 0;
  if (relres_h) *relres_h = 0.0;
  if (n_nodes == 0) return 0;
  FDK_REQUIRE(blk_indptr && blk_indices && indptr && indices && data && b && x && work, FDK_EINVAL, "NULL argument");
  FDK_REQUIRE(index_bytes == 4 || index_bytes == 8, FDK_EINVAL, "index_bytes must be 4 or 8");
  if (check_every < 1) check_every = 1;
  BlockPattern bp;
  bp.n_nodes = n_nodes; bp.nvar = nvar; bp.blk_nnz = blk_nnz; bp.blk_indptr = blk_indptr; bp.blk_indices = blk_indices;
  const int64_t n = (int64_t)nvar * n_nodes, nnz = (int64_t)nvar * nvar * blk_nnz;
  if (index_bytes == 4)
    return pcg_jacobi<int32_t>(n, nnz, (const int32_t*)indptr, (const int32_t*)indices, data, b, x, free_mask, rtol,
                               max_iter, check_every, work, iters_h, relres_h, (cudaStream_t)stream, &bp);
  return pcg_jacobi<int64_t>(n, nnz, (const int64_t*)indptr, (const int64_t*)indices, data, b, x, free_mask, rtol,
                             max_iter, check_every, work, iters_h, relres_h, (cudaStream_t)stream, &bp);
}

static int mpc_from_abi(const fdk_mpc* m, MpcMap* out) {
  FDK_REQUIRE(m != nullptr, FDK_EINVAL, "NULL constraint map");
  FDK_REQUIRE(m->n_nodal >= 0 && m->n_glob >= 0 && m->n_slave >= 0 && m->n_master >= 0, FDK_EINVAL, "negative size");
  FDK_REQUIRE(m->n_nodal + m->n_glob <= 0x7fffffffLL, FDK_EOVERFLOW, "constraint map indices are int32");
  FDK_REQUIRE(m->n_glob <= MPC_MAX_GLOB, FDK_ECAP, "at most 9 global dofs");
  FDK_REQUIRE(m->n_glob == 0 || m->scratch != nullptr, FDK_EINVAL, "NULL scratch (FDK_MPC_SCRATCH_DOUBLES doubles)");
  if (m->n_slave > 0)
    FDK_REQUIRE(m->slave && m->master && (m->n_glob == 0 || m->coef) && m->mst_dof && m->mst_ptr && m->mst_slv, FDK_EINVAL,
                "NULL constraint array");
  out->n_nodal = m->n_nodal; out->n_glob = m->n_glob; out->n_slave = m->n_slave; out->slave = m->slave;
  out->master = m->master; out->coef = m->coef; out->n_master = m->n_master; out->mst_dof = m->mst_dof;
  out->mst_ptr = m->mst_ptr; out->mst_slv = m->mst_slv; out->scratch = m->scratch;
  return 0;
}

int fdk_mpc_expand(const fdk_mpc* mpc, double* x, fdk_stream_t stream) {
  MpcMap m;
  if (int rc = mpc_from_abi(mpc, &m)) return rc;
  FDK_REQUIRE(x != nullptr || m.n_slave == 0, FDK_EINVAL, "NULL argument");
  return mpc_expand(m, x, (cudaStream_t)stream);
}

int fdk_mpc_fold(const fdk_mpc* mpc, double* q, fdk_stream_t stream) {
  MpcMap m;
  if (int rc = mpc_from_abi(mpc, &m)) return rc;
  FDK_REQUIRE(q != nullptr || (m.n_slave == 0 && m.n_glob == 0), FDK_EINVAL, "NULL argument");
  return mpc_fold(m, q, nullptr, false, (cudaStream_t)stream);
}

int fdk_bcsr_pcg_jacobi_mpc(int n_nodes, int nvar, int64_t blk_nnz, const int64_t* blk_indptr,
                            const int32_t* blk_indices, const void* indptr, const void* indices, int index_bytes,
                            const double* data, const double* b, double* x, const uint8_t* free_mask, double rtol,
                            int max_iter, int check_every, double* work, const fdk_mpc* mpc, int* iters_h,
                            double* relres_h, fdk_stream_t stream) {
  FDK_REQUIRE(n_nodes >= 0 && blk_nnz >= 0 && nvar >= 1 && nvar <= 3 && max_iter >= 0 && rtol >= 0.0, FDK_EINVAL,
              "bad size or tolerance");
  if (iters_h) *iters_h = 0;
  if (relres_h) *relres_h = 0.0;
  if (n_nodes == 0) return 0;
  MpcMap m;
  if (int rc = mpc_from_abi(mpc, &m)) return rc;
  FDK_REQUIRE(m.n_nodal == (int64_t)nvar * n_nodes, FDK_EINVAL, "constraint map / matrix size mismatch");
  FDK_REQUIRE(blk_indptr && blk_indices && indptr && indices && data && b && x && work && free_mask, FDK_EINVAL,
              "NULL argument");
  FDK_REQUIRE(index_bytes == 4 || index_bytes == 8, FDK_EINVAL, "index_bytes must be 4 or 8");
  if (check_every < 1) check_every = 1;
  BlockPattern bp;
  bp.n_nodes = n_nodes; bp.nvar = nvar; bp.blk_nnz = blk_nnz; bp.blk_indptr = blk_indptr; bp.blk_indices = blk_indices;
  const int64_t n = (int64_t)nvar * n_nodes, nnz = (int64_t)nvar * nvar * blk_nnz;
  if (index_bytes == 4)
    return pcg_jacobi<int32_t>(n, nnz, (const int32_t*)indptr, (const int32_t*)indices, data, b, x, free_mask, rtol,
                               max_iter, check_every, work, iters_h, relres_h, (cudaStream_t)stream, &bp, &m);
  return pcg_jacobi<int64_t>(n, nnz, (const int64_t*)indptr, (const int64_t*)indices, data, b, x, free_mask, rtol,
                             max_iter, check_every, work, iters_h, relres_h, (cudaStream_t)stream, &bp, &m);
}

int64_t fdk_pcg_multi_work_doubles(int64_t n, int n_rhs) { return pcg_multi_work_doubles(n, n_rhs); }

int fdk_bcsr_pcg_jacobi_multi(int n_nodes, int nvar, int64_t blk_nnz, const int64_t* blk_indptr,
                              const int32_t* blk_indices, const void* indptr, const void* indices, int index_bytes,
                              const double* data, int n_rhs, const double* b, double* x, const uint8_t* free_mask,
                              double rtol, int max_iter, int check_every, double* work, const fdk_mpc* mpc, int* iters_h,
                              double* relres_h, fdk_stream_t stream) {
  FDK_REQUIRE(n_nodes >= 0 && blk_nnz >= 0 && nvar >= 1 && nvar <= 3 && max_iter >= 0 && rtol >= 0.0, FDK_EINVAL,
              "bad size or tolerance");
  FDK_REQUIRE(n_rhs == 3 || n_rhs == 6, FDK_EINVAL, "n_rhs must be 3 or 6");
  if (iters_h) *iters_h = 0;
  if (relres_h)
    for (int k = 0; k < n_rhs; ++k) relres_h[k] = 0.0;
  if (n_nodes == 0) return 0;
  MpcMap m;
  if (mpc != nullptr) {
    if (int rc = mpc_from_abi(mpc, &m)) return rc;
    FDK_REQUIRE(m.n_nodal == (int64_t)nvar * n_nodes, FDK_EINVAL, "constraint map / matrix size mismatch");
    FDK_REQUIRE(free_mask != nullptr, FDK_EINVAL, "the constrained solve needs the mask of the independent dofs");
  }
  FDK_REQUIRE(blk_indptr && blk_indices && indptr && indices && data && b && x && work, FDK_EINVAL, "NULL argument");
  FDK_REQUIRE(index_bytes == 4 || index_bytes == 8, FDK_EINVAL, "index_bytes must be 4 or 8");
  if (check_every < 1) check_every = 1;
  BlockPattern bp;
  bp.n_nodes = n_nodes; bp.nvar = nvar; bp.blk_nnz = blk_nnz; bp.blk_indptr = blk_indptr; bp.blk_indices = blk_indices;
  const int64_t n = (int64_t)nvar * n_nodes;
  const MpcMap* mp = mpc != nullptr ? &m : nullptr;
#define FDK_MULTI(IDX_, R_)                                                                                          \
  return pcg_jacobi_multi<IDX_, R_>(n, (const IDX_*)indptr, (const IDX_*)indices, data, b, x, free_mask, rtol, max_iter, \
                                    check_every, work, iters_h, relres_h, (cudaStream_t)stream, bp, mp)
  if (index_bytes == 4) {
    if (n_rhs == 3) FDK_MULTI(int32_t, 3);
    FDK_MULTI(int32_t, 6);
  }
  if (n_rhs == 3) FDK_MULTI(int64_t, 3);
  FDK_MULTI(int64_t, 6);
#undef FDK_MULTI
}

int fdk_csr_diagonal(int64_t n_rows, const void* indptr, const void* indices, int index_bytes, const double* data,
                     double* diag, fdk_stream_t stream) {
  FDK_REQUIRE(n_rows >= 0, FDK_EINVAL, "negative size");
  if (n_rows == 0) return 0;
  FDK_REQUIRE(indptr && indices && data && diag, FDK_EINVAL, "NULL argument");
  FDK_REQUIRE(index_bytes == 4 || index_bytes == 8, FDK_EINVAL, "index_bytes must be 4 or 8");
  const unsigned grid = (unsigned)((n_rows + 255) / 256);
  if (index_bytes == 4)
    k_csr_diagonal<int32_t><<<grid, 256, 0, (cudaStream_t)stream>>>(n_rows, (const int32_t*)indptr,
                                                                      (const int32_t*)indices, data, diag);
  else
    k_csr_diagonal<int64_t><<<grid, 256, 0, (cudaStream_t)stream>>>(n_rows, (const int64_t*)indptr,
                                                                      (const int64_t*)indices, data, diag);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

int64_t fdk_pcg_work_doubles(int64_t n) { return 5 * n + 2 * RED_BLOCKS + S_COUNT; }

int fdk_pcg_jacobi(int64_t n, int64_t nnz, const void* indptr, const void* indices, int index_bytes, const double* data,
                   const double* b, double* x, const uint8_t* free_mask, double rtol, int max_iter, int check_every,
                   double* work, int* iters_h, double* relres_h, fdk_stream_t stream) {
  FDK_REQUIRE(n >= 0 && nnz >= 0 && max_iter >= 0 && rtol >= 0.0, FDK_EINVAL, "bad size or tolerance");
  if (iters_h) *iters_h = 0;
  if (relres_h) *relres_h = 0.0;
  if (n == 0) return 0;
  FDK_REQUIRE(indptr && indices && data && b && x && work, FDK_EINVAL, "NULL argument");
  FDK_REQUIRE(index_bytes == 4 || index_bytes == 8, FDK_EINVAL, "index_bytes must be 4 or 8");
  if (check_every < 1) check_every = 1;
  if (index_bytes == 4)
    return pcg_jacobi<int32_t>(n, nnz, (const int32_t*)indptr, (const int32_t*)indices, data, b, x, free_mask, rtol,
                               max_iter, check_every, work, iters_h, relres_h, (cudaStream_t)stream);
  return pcg_jacobi<int64_t>(n, nnz, (const int64_t*)indptr, (const int64_t*)indices, data, b, x, free_mask, rtol,
                             max_iter, check_every, work, iters_h, relres_h, (cudaStream_t)stream);
}

int fdk_gp_to_node(int nne, int ngp, int n_nodes, int64_t n_elems, const int64_t* node_ptr, const int32_t* node_inc,
                   const double* P_h, const double* field, int ncomp, int64_t comp_stride, int64_t gp_stride,
                   int von_mises, double* out, fdk_stream_t stream) {
  FDK_REQUIRE(nne > 0 && nne <= MAX_NNE && ngp > 0 && ngp <= MAX_NGP && n_nodes >= 0 && n_elems >= 0, FDK_EINVAL, "bad sizes");
  FDK_REQUIRE(von_mises ? ncomp == 6 : (ncomp >= 1 && ncomp <= 6), FDK_EINVAL, "ncomp must be 1..6 (6 for von Mises)");
  if (n_nodes == 0) return 0;
  FDK_REQUIRE(node_ptr && node_inc && P_h && field && out, FDK_EINVAL, "NULL argument");
  ConvArgs a{};
  a.nne = nne; a.ngp = ngp; a.n_nodes = n_nodes; a.ncomp = ncomp; a.von_mises = von_mises; a.n_elems = n_elems;
  a.comp_stride = comp_stride; a.gp_stride = gp_stride; a.node_ptr = node_ptr; a.node_inc = node_inc;
  a.field = field; a.out = out;
  for (int i = 0; i < nne * ngp; ++i) a.P[i] = P_h[i];
  k_gp_to_node<6><<<(unsigned)((n_nodes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

int fdk_gp_to_element(int ngp, int64_t n_elems, const double* field, int ncomp, int64_t comp_stride, int64_t gp_stride,
                      int von_mises, double* out, fdk_stream_t stream) {
  FDK_REQUIRE(ngp > 0 && ngp <= MAX_NGP && n_elems >= 0, FDK_EINVAL, "bad sizes");
  FDK_REQUIRE(von_mises ? ncomp == 6 : (ncomp >= 1 && ncomp <= 6), FDK_EINVAL, "ncomp must be 1..6 (6 for von Mises)");
  if (n_elems == 0) return 0;
  FDK_REQUIRE(field && out, FDK_EINVAL, "NULL argument");
  ConvArgs a{};
  a.ngp = ngp; a.ncomp = ncomp; a.von_mises = von_mises; a.n_elems = n_elems;
  a.comp_stride = comp_stride; a.gp_stride = gp_stride; a.field = field; a.out = out;
  k_gp_to_element<6><<<(unsigned)((n_elems + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

int fdk_gp_von_mises(int64_t n_gp, const double* field, int64_t comp_stride, int64_t gp_stride, double* out,
                     fdk_stream_t stream) {
  FDK_REQUIRE(n_gp >= 0, FDK_EINVAL, "negative n_gp");
  if (n_gp == 0) return 0;
  FDK_REQUIRE(field && out, FDK_EINVAL, "NULL argument");
  k_gp_von_mises<<<(unsigned)((n_gp + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_gp, field, comp_stride, gp_stride, out);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

int fdk_gather_f64(int64_t n, const int64_t* index, const double* src, double* dst, fdk_stream_t stream) {
  if (n <= 0) return 0;
  FDK_REQUIRE(index && src && dst, FDK_EINVAL, "NULL argument");
  k_gather_f64<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, index, src, dst);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

int fdk_copy_segments(int n_seg, const int64_t* seg_src, const int64_t* seg_dst, const int64_t* seg_len,
                      int64_t max_len, const double* src, double* dst, fdk_stream_t stream) {
  FDK_REQUIRE(n_seg >= 0 && max_len >= 0, FDK_EINVAL, "negative size");
  if (n_seg == 0 || max_len == 0) return 0;
  FDK_REQUIRE(seg_src && seg_dst && seg_len && src && dst, FDK_EINVAL, "NULL argument");
  FDK_REQUIRE(n_seg <= 65535, FDK_EINVAL, "too many segments");
  int64_t bx = (max_len + 255) / 256;
  if (bx > 1024) bx = 1024;
  k_copy_segments<<<dim3((unsigned)bx, (unsigned)n_seg), 256, 0, (cudaStream_t)stream>>>(n_seg, seg_src, seg_dst, seg_len,
                                                                                      src, dst);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

int fdk_scatter_add_f64(int64_t n, const int64_t* index, const double* src, double* dst, fdk_stream_t stream) {
  if (n <= 0) return 0;
  FDK_REQUIRE(index && src && dst, FDK_EINVAL, "NULL argument");
  k_scatter_add_f64<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, index, src, dst);
  FDK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
