// fp64_peaks.cu -- microbenchmarks that size the assembly kernels on B200 (sm_100a):
//   1. DFMA throughput (CUDA-core FP64)
//   2. DMMA m8n8k4 / m16n8k8 / m16n8k16 throughput (FP64 tensor-core path)
//   3. both at once (do they share a pipe?)
//   4. shared-memory LDS.64 / LDS.128 bandwidth
//   5. streaming store bandwidth (what the K write can reach)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
        "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void k_dmma884(double* out, double a, double b) {
  double d[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma884(d[i][0], d[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1688(double* out, double x) {
  double d[4][4];
  double a[4] = {x, x + 1, x + 2, x + 3}, b[2] = {x, -x};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = threadIdx.x + i + j;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmma1688(d[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, double x) {
  double d[4][4];
  double a[8], b[4] = {x, -x, x * 2, x * 3};
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = x + i;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = threadIdx.x + i + j;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmma16816(d[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// half of the warps DFMA, half DMMA
__global__ void k_mixed(double* out, double a, double b) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp & 1) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
  } else {
    double d[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(d[i][0], d[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory read bandwidth: every lane reads its own address (no broadcast), conflict-free
template <int VEC>
__global__ void k_lds(double* out) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double s = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = ((warp * 8 + j + it) * 32 * VEC + lane * VEC) & 8191;
      if (VEC == 1) {
        s += sm[idx];
      } else {
        const double2 v = *reinterpret_cast<const double2*>(sm + idx);
        s += v.x + v.y;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_store(double* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) __stcs(out + i, (double)i);
}
__global__ void k_store2(double2* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    __stcs(out + i, make_double2((double)i, 1.0));
}

template <class F>
float time_ms(F f, int rep = 5) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < rep; ++r) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs\n", p.name, sms);
  double* out;
  CK(cudaMalloc(&out, (size_t)sms * 8 * 1024 * 8));
  for (int warps : {4, 8, 16, 32}) {
    const int threads = warps * 32;
    for (int ctas : {1, 2}) {
      if (threads * ctas > 2048) continue;
      const int grid = sms * ctas;
      float ms = time_ms([&] { k_dfma<<<grid, threads>>>(out, 1.0000001, 1e-9); });
      double fl = 2.0 * 16 * ITERS * (double)grid * threads;
      printf("DFMA      warps/CTA %2d CTAs/SM %d : %7.3f ms  %6.2f TFLOP/s\n", warps, ctas, ms, fl / ms / 1e9);
    }
  }
  for (int warps : {4, 8, 16}) {
    const int threads = warps * 32, grid = sms;
    float ms = time_ms([&] { k_dmma884<8><<<grid, threads>>>(out, 1.0000001, 1e-9); });
    double fl = 2.0 * 256 * 8 * ITERS * (double)grid * warps;
    printf("DMMA884 ILP8 warps/CTA %2d : %7.3f ms  %6.2f TFLOP/s\n", warps, ms, fl / ms / 1e9);
    ms = time_ms([&] { k_dmma884<2><<<grid, threads>>>(out, 1.0000001, 1e-9); });
    fl = 2.0 * 256 * 2 * ITERS * (double)grid * warps;
    printf("DMMA884 ILP2 warps/CTA %2d : %7.3f ms  %6.2f TFLOP/s\n", warps, ms, fl / ms / 1e9);
    ms = time_ms([&] { k_dmma1688<<<grid, threads>>>(out, 1.0000001); });
    fl = 2.0 * 1024 * 4 * ITERS * (double)grid * warps;
    printf("DMMA1688  warps/CTA %2d : %7.3f ms  %6.2f TFLOP/s\n", warps, ms, fl / ms / 1e9);
    ms = time_ms([&] { k_dmma16816<<<grid, threads>>>(out, 1.0000001); });
    fl = 2.0 * 2048 * 4 * ITERS * (double)grid * warps;
    printf("DMMA16816 warps/CTA %2d : %7.3f ms  %6.2f TFLOP/s\n", warps, ms, fl / ms / 1e9);
    ms = time_ms([&] { k_mixed<<<grid, threads>>>(out, 1.0000001, 1e-9); });
    fl = 2.0 * ITERS * (double)grid * (warps / 2) * (16.0 * 32 + 256.0 * 8);
    printf("MIXED     warps/CTA %2d : %7.3f ms  %6.2f TFLOP/s (half DFMA warps, half DMMA warps)\n", warps, ms,
           fl / ms / 1e9);
  }
  CK(cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_lds<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int warps : {8, 16, 32}) {
    const int threads = warps * 32, grid = sms;
    float ms = time_ms([&] { k_lds<1><<<grid, threads, 65536>>>(out); });
    double by = 8.0 * 8 * ITERS * (double)grid * threads;
    printf("LDS.64  warps/CTA %2d : %7.3f ms  %7.1f GB/s  = %5.1f B/clk/SM @1.965GHz\n", warps, ms, by / ms / 1e6,
           by / ms / 1e6 / sms / 1.965);
    ms = time_ms([&] { k_lds<2><<<grid, threads, 65536>>>(out); });
    by = 16.0 * 8 * ITERS * (double)grid * threads;
    printf("LDS.128 warps/CTA %2d : %7.3f ms  %7.1f GB/s  = %5.1f B/clk/SM @1.965GHz\n", warps, ms, by / ms / 1e6,
           by / ms / 1e6 / sms / 1.965);
  }
  {
    const size_t n = (size_t)2 << 30;  // 16 GiB of doubles
    double* big;
    CK(cudaMalloc(&big, n * 8));
    for (int mult : {2, 4, 8, 16}) {
      float ms = time_ms([&] { k_store<<<sms * mult, 512>>>(big, n); }, 3);
      printf("STORE  f64   grid %2dxSMs : %7.3f ms  %7.1f GB/s\n", mult, ms, n * 8.0 / ms / 1e6);
      ms = time_ms([&] { k_store2<<<sms * mult, 512>>>((double2*)big, n / 2); }, 3);
      printf("STORE  f64x2 grid %2dxSMs : %7.3f ms  %7.1f GB/s\n", mult, ms, n * 8.0 / ms / 1e6);
    }
    float ms = time_ms([&] { cudaMemsetAsync(big, 0, n * 8); }, 3);
    printf("cudaMemset 16 GiB : %7.3f ms  %7.1f GB/s\n", ms, n * 8.0 / ms / 1e6);
    cudaFree(big);
  }
  return 0;
}
