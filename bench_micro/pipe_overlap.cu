// pipe_overlap.cu -- do FP64 FMAs and shared-memory accesses overlap on B200 (sm_100a)?
// The cluster kernel's phases are each a mix of DFMA and LDS/STS; this measures, per SM,
//   (a) DFMA alone, (b) LDS.64 / LDS.128 / STS.64 alone for a few address patterns (honest: volatile asm),
//   (c) DFMA warps and LDS warps running side by side, (d) DFMA and LDS interleaved in the same warp,
//   (e) SHFL alone and next to DFMA.
// Times are SM cycles per "unit" (one warp-instruction of each kind), from clock64 inside the kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_overlap pipe_overlap.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

constexpr int ITERS = 2048;

// integer-typed so that consuming the loaded value does not touch the FP64 pipe
__device__ __forceinline__ unsigned lds64(unsigned addr) {
  unsigned x, y;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(addr) : "memory");
  return x ^ y;
}
__device__ __forceinline__ unsigned lds128(unsigned addr) {
  unsigned x, y, z, w;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr) : "memory");
  return x ^ y ^ z ^ w;
}
__device__ __forceinline__ void sts64(unsigned addr, unsigned v) {
  asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(v), "r"(v + 1) : "memory");
}

// MODE bits: 1 = DFMA, 2 = memory ops (compile time: no branches inside the timed loop).
// PAT: 0 LDS.64 distinct conflict-free, 1 LDS.64 4 distinct addresses (8 lanes share), 2 LDS.128 distinct,
//      3 LDS.128 8 distinct addresses, 4 STS.64 distinct, 5 SHFL.64 (two SHFL.32), 6 LDS.64 2-way bank conflict,
//      7 LDS.64 all lanes same address
template <int PAT, int NF, int NM, int MODE>
__device__ __forceinline__ double body(unsigned addr, double a, double b) {
  double acc[NF];
#pragma unroll
  for (int i = 0; i < NF; ++i) acc[i] = threadIdx.x + i;
  unsigned s = threadIdx.x;
#pragma unroll 2
  for (int it = 0; it < ITERS; ++it) {
    if (MODE & 1) {
#pragma unroll
      for (int i = 0; i < NF; ++i) acc[i] = fma(acc[i], a, b);
    }
    if (MODE & 2) {
#pragma unroll
      for (int j = 0; j < NM; ++j) {
        const unsigned ad = addr + (((j + it) * 272) & 2047);
        if (PAT == 0 || PAT == 1 || PAT == 6 || PAT == 7) s ^= lds64(ad);
        else if (PAT == 2 || PAT == 3) s ^= lds128(ad);
        else if (PAT == 4) sts64(ad, s + j);
        else { s ^= __shfl_xor_sync(0xffffffffu, s + j, 1 + (j & 15)); s ^= __shfl_xor_sync(0xffffffffu, s + 2 * j, 1 + (j & 15)); }
      }
    }
  }
  double r = s;
#pragma unroll
  for (int i = 0; i < NF; ++i) r += acc[i];
  return r;
}

template <int PAT, int NF, int NM, int MODE_LO, int MODE_HI>
__global__ void k_mix(double* out, long long* clk, int split, double a, double b) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (warp & 7) * 512 * 8;
  unsigned addr;
  if (PAT == 0 || PAT == 4) addr = base + lane * 8;
  else if (PAT == 1) addr = base + (lane >> 3) * 8 * 5;
  else if (PAT == 2) addr = base + lane * 16;
  else if (PAT == 3) addr = base + (lane >> 2) * 16 * 3;
  else if (PAT == 6) addr = base + lane * 16;  // 64-bit loads at stride 16 B: 2-way conflict
  else addr = base;
  __syncthreads();
  const long long t0 = clock64();
  double r;
  if (warp < split) r = body<PAT, NF, NM, MODE_LO>(addr, a, b);
  else r = body<PAT, NF, NM, MODE_HI>(addr, a, b);
  __syncthreads();
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int PAT, int NF, int NM, int MODE_LO, int MODE_HI>
double run(int warps, int split, double* out, long long* clk, int sms) {
  auto k = k_mix<PAT, NF, NM, MODE_LO, MODE_HI>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  const size_t smem = 160 * 1024;  // one CTA per SM
  k<<<sms, warps * 32, smem>>>(out, clk, split, 1.0000001, 1e-9);
  k<<<sms, warps * 32, smem>>>(out, clk, split, 1.0000001, 1e-9);
  CK(cudaDeviceSynchronize());
  long long h[256];
  CK(cudaMemcpy(h, clk, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  double m = 0;
  for (int i = 0; i < sms; ++i) m += (double)h[i];
  return m / sms / ITERS;  // cycles per loop iteration
}

template <int PAT>
void suite(const char* name, int warps, double* out, long long* clk, int sms) {
  constexpr int NF = 16, NM = 8;
  const double f = run<PAT, NF, NM, 1, 1>(warps, warps, out, clk, sms);         // all warps DFMA
  const double m = run<PAT, NF, NM, 2, 2>(warps, warps, out, clk, sms);         // all warps memory
  const double both = run<PAT, NF, NM, 3, 3>(warps, warps, out, clk, sms);      // every warp does both, interleaved
  const double side = run<PAT, NF, NM, 1, 2>(warps, warps / 2, out, clk, sms);  // half the warps each
  // per SM and iteration: `warps` warps x NF DFMA -> ideal warps*NF*2/4 cycles (16 lanes/clk/SMSP)
  printf("%-24s warps %2d | DFMA only %6.1f clk/iter (ideal %5.1f) | mem only %6.1f (%.2f clk per warp-instr) | "
         "same warp both %6.1f (sum %6.1f, max %6.1f) | half/half %6.1f (max of halves %6.1f)\n",
         name, warps, f, warps * NF * 2.0 / 4, m, m / (warps * NM), both, f + m, f > m ? f : m, side,
         (f > m ? f : m) / 2);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs; per loop iteration every active warp issues 16 DFMA and/or 8 memory instructions\n", p.name,
         sms);
  double* out;
  long long* clk;
  CK(cudaMalloc(&out, (size_t)sms * 1024 * 8));
  CK(cudaMalloc(&clk, 256 * sizeof(long long)));
  for (int warps : {8, 16, 32}) {
    suite<0>("LDS.64 distinct", warps, out, clk, sms);
    suite<1>("LDS.64 4 addresses", warps, out, clk, sms);
    suite<7>("LDS.64 1 address", warps, out, clk, sms);
    suite<6>("LDS.64 2-way conflict", warps, out, clk, sms);
    suite<2>("LDS.128 distinct", warps, out, clk, sms);
    suite<3>("LDS.128 8 addresses", warps, out, clk, sms);
    suite<4>("STS.64 distinct", warps, out, clk, sms);
    suite<5>("SHFL.64", warps, out, clk, sms);
  }
  return 0;
}
