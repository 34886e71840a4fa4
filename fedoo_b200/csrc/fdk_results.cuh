// fdk_results.cuh -- results extraction on the device: Gauss-point fields to nodes / elements and
// the von Mises stress, what Problem.get_results / Mesh.convert_data do after every solve.
//
// Reference: fedoo/core/mesh.py:1149-1160 (GP -> node matrix: pinv of the shape functions at the Gauss points, divided
// by the number of elements around the node), :1297-1303 (GP -> element: mean over the Gauss points),
// fedoo/util/voigt_tensors.py:270-283 (von Mises), fedoo/core/output.py:190-197 (Stress_vm is taken at the Gauss
// points, then converted).  The reference multiplies by an (n_nodes x n_gp) CSR matrix; here one thread per node
// walks the node's incidences (deterministic order, no atomics).
#pragma once
#include "fdk_common.cuh"

namespace fdk {

struct ConvArgs {
  int nne, ngp, n_nodes, ncomp, von_mises;
  int64_t n_elems;
  int64_t comp_stride, gp_stride;  // field value (gp n, component c) at field[n * gp_stride + c * comp_stride]
  const int64_t* node_ptr;         // [n_nodes + 1]
  const int32_t* node_inc;         // incidences of every node: element * nne + local node
  const double* field;
  double* out;                     // [ncomp_out][n_nodes] or [ncomp_out][n_elems]
  double P[MAX_NNE * MAX_NGP];     // pinv(N_gp): [local node][gp]
};

__device__ __forceinline__ double von_mises6(const double (&s)[6]) {
  const double a = s[0] - s[1], b = s[1] - s[2], c = s[0] - s[2];
  return sqrt(0.5 * (a * a + b * b + c * c + 6.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5])));
}

// MAXC components handled per thread (6 covers the Voigt tensors; wider fields are converted in slices by the host)
template <int MAXC>
__global__ void __launch_bounds__(256) k_gp_to_node(const __grid_constant__ ConvArgs a) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= a.n_nodes) return;
  const int64_t k0 = a.node_ptr[I], k1 = a.node_ptr[I + 1];
  const int nc_out = a.von_mises ? 1 : a.ncomp;
  double acc[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) acc[c] = 0.0;
  for (int64_t k = k0; k < k1; ++k) {
    const int32_t inc = a.node_inc[k];
    const int64_t e = inc / a.nne;
    const int i = inc - (int)e * a.nne;
    for (int g = 0; g < a.ngp; ++g) {
      const double w = a.P[i * a.ngp + g];
      const double* f = a.field + ((int64_t)g * a.n_elems + e) * a.gp_stride;
      if (a.von_mises) {
        double s[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) s[c] = f[c * a.comp_stride];
        acc[0] = fma(w, von_mises6(s), acc[0]);
      } else {
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (c < a.ncomp) acc[c] = fma(w, f[c * a.comp_stride], acc[c]);
      }
    }
  }
  const double inv = k1 > k0 ? 1.0 / (double)(k1 - k0) : 0.0;
  for (int c = 0; c < nc_out; ++c) a.out[(int64_t)c * a.n_nodes + I] = acc[c] * inv;
}

template <int MAXC>
__global__ void __launch_bounds__(256) k_gp_to_element(const __grid_constant__ ConvArgs a) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n_elems) return;
  const int nc_out = a.von_mises ? 1 : a.ncomp;
  double acc[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) acc[c] = 0.0;
  for (int g = 0; g < a.ngp; ++g) {
    const double* f = a.field + ((int64_t)g * a.n_elems + e) * a.gp_stride;
    if (a.von_mises) {
      double s[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) s[c] = f[c * a.comp_stride];
      acc[0] += von_mises6(s);
    } else {
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < a.ncomp) acc[c] += f[c * a.comp_stride];
    }
  }
  for (int c = 0; c < nc_out; ++c) a.out[(int64_t)c * a.n_elems + e] = acc[c] / (double)a.ngp;
}

__global__ void __launch_bounds__(256) k_gp_von_mises(int64_t n_gp, const double* __restrict__ field,
                                                      int64_t comp_stride, int64_t gp_stride, double* __restrict__ out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_gp) return;
  double s[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) s[c] = field[n * gp_stride + c * comp_stride];
  out[n] = von_mises6(s);
}

}  // namespace fdk
