// fdk_tables.cuh -- element tables (shape functions / reference derivatives / Gauss rules).
//
// Restated from the element definitions of the reference (mathematical content only):
//   hex8  : fedoo/lib_elements/hexahedron.py:22-27 (2x2x2 Gauss, xi slowest), :134-135 (w = 1),
//           :178-248 (trilinear shape functions, node order bottom face CCW then top face)
//   tet4  : fedoo/lib_elements/tetrahedron.py:21-29 (4 points), :72-73 (w = 1/24), :106-130
//           (N = [eta, zeta, 1-xi-eta-zeta, xi])
//   tet10 : fedoo/lib_elements/tetrahedron.py:35-61 (15 points), :76-97 (weights), :133-208
//   quad4 : fedoo/lib_elements/quadrangle.py:24-26 (2x2 Gauss order (-,-),(+,-),(+,+),(-,+)), :125-167
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstring>

#include "fdk_common.cuh"

namespace fdk {

__constant__ ElemTable c_tab[4];

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}

int elem_dims(int elem_type, int* nne, int* ngp, int* dim) {
  switch (elem_type) {
    case FDK_HEX8: *nne = 8; *ngp = 8; *dim = 3; return 0;
    case FDK_TET4: *nne = 4; *ngp = 4; *dim = 3; return 0;
    case FDK_TET10: *nne = 10; *ngp = 15; *dim = 3; return 0;
    case FDK_QUAD4: *nne = 4; *ngp = 4; *dim = 2; return 0;
  }
  set_error("unknown element type %d", elem_type);
  return FDK_EINVAL;
}

namespace detail {

inline void fill_hex8(ElemTable& t) {
  const double a = 0.5773502691896258;
  const double nd[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                           {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
  int g = 0;
  for (int sx = -1; sx <= 1; sx += 2)
    for (int sy = -1; sy <= 1; sy += 2)
      for (int sz = -1; sz <= 1; sz += 2, ++g) {
        const double x[3] = {sx * a, sy * a, sz * a};
        t.w[g] = 1.0;
        for (int k = 0; k < 8; ++k) {
          const double f0 = 1 + nd[k][0] * x[0], f1 = 1 + nd[k][1] * x[1], f2 = 1 + nd[k][2] * x[2];
          t.N[g * 8 + k] = 0.125 * f0 * f1 * f2;
          t.dN[(g * 3 + 0) * 8 + k] = 0.125 * nd[k][0] * f1 * f2;
          t.dN[(g * 3 + 1) * 8 + k] = 0.125 * nd[k][1] * f0 * f2;
          t.dN[(g * 3 + 2) * 8 + k] = 0.125 * nd[k][2] * f0 * f1;
        }
      }
}

inline void fill_quad4(ElemTable& t) {
  const double a = 1.0 / std::sqrt(3.0);
  const double nd[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
  const double gp[4][2] = {{-a, -a}, {a, -a}, {a, a}, {-a, a}};
  for (int g = 0; g < 4; ++g) {
    t.w[g] = 1.0;
    for (int k = 0; k < 4; ++k) {
      const double f0 = 1 + nd[k][0] * gp[g][0], f1 = 1 + nd[k][1] * gp[g][1];
      t.N[g * 4 + k] = 0.25 * f0 * f1;
      t.dN[(g * 2 + 0) * 4 + k] = 0.25 * nd[k][0] * f1;
      t.dN[(g * 2 + 1) * 4 + k] = 0.25 * nd[k][1] * f0;
    }
  }
}

inline void fill_tet4(ElemTable& t) {
  const double a = 0.1381966011250105, b = 0.5854101966249685;
  const double gp[4][3] = {{a, a, a}, {a, a, b}, {a, b, a}, {b, a, a}};
  const double d[3][4] = {{0, 0, -1, 1}, {1, 0, -1, 0}, {0, 1, -1, 0}};
  for (int g = 0; g < 4; ++g) {
    t.w[g] = 1.0 / 24;
    const double xi = gp[g][0], eta = gp[g][1], zeta = gp[g][2];
    const double N[4] = {eta, zeta, 1 - xi - eta - zeta, xi};
    for (int k = 0; k < 4; ++k) {
      t.N[g * 4 + k] = N[k];
      for (int r = 0; r < 3; ++r) t.dN[(g * 3 + r) * 4 + k] = d[r][k];
    }
  }
}

inline void fill_tet10(ElemTable& t) {
  const double a = 0.25, b1 = 0.3197936278296299, b2 = 0.09197107805272303, c1 = 0.040619116511110234,
               c2 = 0.724086765841831, d = 0.05635083268962915, e = 0.4436491673103708;
  const double gp[15][3] = {{a, a, a},    {b1, b1, b1}, {b1, b1, c1}, {b1, c1, b1}, {c1, b1, b1},
                            {b2, b2, b2}, {b2, b2, c2}, {b2, c2, b2}, {c2, b2, b2}, {d, d, e},
                            {d, e, d},    {e, d, d},    {d, e, e},    {e, d, e},    {e, e, d}};
  const double f1 = 0.011511367871045397, f2 = 0.01198951396316977;
  const double w[15] = {8.0 / 405, f1, f1, f1, f1, f2, f2, f2, f2, 5.0 / 567, 5.0 / 567,
                        5.0 / 567, 5.0 / 567, 5.0 / 567, 5.0 / 567};
  for (int g = 0; g < 15; ++g) {
    t.w[g] = w[g];
    const double xi = gp[g][0], eta = gp[g][1], zeta = gp[g][2];
    const double m = 1 - xi - eta - zeta;
    const double N[10] = {eta * (2 * eta - 1), zeta * (2 * zeta - 1), m * (1 - 2 * xi - 2 * eta - 2 * zeta),
                          xi * (2 * xi - 1),   4 * eta * zeta,        4 * zeta * m,
                          4 * eta * m,         4 * xi * eta,          4 * xi * zeta,
                          4 * xi * m};
    const double D[3][10] = {
        {0.0, 0.0, 1 - 4 * m, -1 + 4 * xi, 0.0, -4 * zeta, -4 * eta, 4 * eta, 4 * zeta, 4 * (m - xi)},
        {-1 + 4 * eta, 0.0, 1 - 4 * m, 0.0, 4 * zeta, -4 * zeta, 4 * (m - eta), 4 * xi, 0.0, -4 * xi},
        {0.0, -1 + 4 * zeta, 1 - 4 * m, 0.0, 4 * eta, 4 * (m - zeta), -4 * eta, 0.0, 4 * xi, -4 * xi}};
    for (int k = 0; k < 10; ++k) {
      t.N[g * 10 + k] = N[k];
      for (int r = 0; r < 3; ++r) t.dN[(g * 3 + r) * 10 + k] = D[r][k];
    }
  }
}

struct HostTables {
  ElemTable t[4];
  HostTables() {
    std::memset(t, 0, sizeof(t));
    fill_hex8(t[FDK_HEX8]);
    fill_tet4(t[FDK_TET4]);
    fill_tet10(t[FDK_TET10]);
    fill_quad4(t[FDK_QUAD4]);
  }
};

}  // namespace detail

const ElemTable& host_table(int elem_type) {
  static detail::HostTables tabs;
  return tabs.t[elem_type];
}

int ensure_device_tables() {
  static bool done[64] = {false};
  int dev = 0;
  FDK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && done[dev]) return 0;
  host_table(0);
  ElemTable tmp[4];
  for (int i = 0; i < 4; ++i) tmp[i] = host_table(i);
  FDK_CUDA(cudaMemcpyToSymbol(c_tab, tmp, sizeof(tmp)));
  if (dev >= 0 && dev < 64) done[dev] = true;
  return 0;
}

}  // namespace fdk
