// fdk_common.cuh -- shared declarations of libfdk (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

#include "fdk.h"

namespace fdk {

// ---- error plumbing ---------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define FDK_CUDA(call)                                   \
  do {                                                   \
    cudaError_t _e = (call);                             \
    if (_e != cudaSuccess) return ::fdk::cuda_fail(_e, #call); \
  } while (0)

#define FDK_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::fdk::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

// ---- element traits ------------------------------------------------------------------
// THREADS = CTA size of the cluster kernels = 2 x max incidences (node, element) per cluster
// (phase 2 runs two threads per incidence, half a block row each); half-size variant: THREADS / 2.
template <int ID_, int NNE_, int NGP_, int DIM_, int THREADS_>
struct ElemTraits {
  static constexpr int ID = ID_, NNE = NNE_, NGP = NGP_, DIM = DIM_, THREADS = THREADS_;
};
using Hex8 = ElemTraits<FDK_HEX8, 8, 8, 3, 512>;
using Tet4 = ElemTraits<FDK_TET4, 4, 4, 3, 512>;
using Tet10 = ElemTraits<FDK_TET10, 10, 15, 3, 256>;
using Quad4 = ElemTraits<FDK_QUAD4, 4, 4, 2, 512>;

constexpr int MAX_NGP = 15, MAX_NNE = 10, MAX_DIM = 3;
constexpr int J2_R1 = 10;  // doubles per Gauss point of the structured J2 tangent (csrc/fdk_gp.cuh)

// Gauss weights, shape functions and reference derivatives at the Gauss points.
struct ElemTable {
  double w[MAX_NGP];
  double N[MAX_NGP * MAX_NNE];             // [g][k], row stride = nne
  double dN[MAX_NGP * MAX_DIM * MAX_NNE];  // [g][d][k], strides dim*nne, nne
};

// host-side table (computed once) and upload to the current device's constant memory.
const ElemTable& host_table(int elem_type);
int ensure_device_tables();
int elem_dims(int elem_type, int* nne, int* ngp, int* dim);

}  // namespace fdk
