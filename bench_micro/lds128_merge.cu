// lds128_merge.cu -- how does the B200 LSU merge 128-bit shared-memory loads whose lanes share addresses?
// The block phase of the cluster kernel reads dN/dx with LDS.128 where ~10 distinct 16-byte chunks serve 32 lanes;
// ncu charges 3.6 wavefronts per instruction.  This measures SM cycles per warp-instruction for lane -> chunk maps:
//   0  32 distinct chunks (contiguous 512 B)
//   1  8 distinct chunks, chunk = lane & 7           (every quarter-warp reads the same 8)
//   2  8 distinct chunks, chunk = lane >> 2          (4 consecutive lanes share; 2 chunks per quarter)
//   3  4 distinct chunks, chunk = lane >> 3          (one per quarter-warp)
//   4  kernel pattern A: lane = 4 * incidence + part, 3 elements of stride 108 chunks: chunk = elem * 108 + 3 * part
//   5  kernel pattern B: lane = 8 * part + incidence (a quarter-warp = one part): same chunks
//   6  like 4 with element stride 105 chunks (an odd multiple of 16 B ... 1680 B)
//   7  1 chunk for all lanes
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds128_merge lds128_merge.cu
#include <cuda_runtime.h>

#include <cstdio>

constexpr int ITERS = 4096;

__device__ __forceinline__ unsigned lds128(unsigned addr) {
  unsigned x, y, z, w;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr) : "memory");
  return x ^ y ^ z ^ w;
}

__global__ void k(int pat, unsigned* out, long long* clk) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int chunk;
  const int inc_a = lane >> 2, part_a = lane & 3, inc_b = lane & 7, part_b = lane >> 3;
  switch (pat) {
    case 0: chunk = lane; break;
    case 1: chunk = lane & 7; break;
    case 2: chunk = lane >> 2; break;
    case 3: chunk = lane >> 3; break;
    case 4: chunk = (inc_a * 3 / 8) * 108 + 3 * part_a; break;
    case 5: chunk = (inc_b * 3 / 8) * 108 + 3 * part_b; break;
    case 6: chunk = (inc_a * 3 / 8) * 105 + 3 * part_a; break;
    default: chunk = 0; break;
  }
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + chunk * 16;
  unsigned s = threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s ^= lds128(base + (((j + it) & 7) * 16 * 13));  // a few rotating row offsets
  }
  __syncthreads();
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  unsigned* out;
  long long* clk;
  cudaMalloc(&out, sms * 1024 * 4);
  cudaMalloc(&clk, 256 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const char* names[] = {"32 distinct", "8 distinct, lane & 7", "8 distinct, lane >> 2", "4 distinct, lane >> 3",
                         "kernel A (4 inc + part)", "kernel B (8 part + inc)", "kernel A, stride 105", "1 chunk"};
  for (int warps : {16, 32}) {
    for (int pat = 0; pat < 8; ++pat) {
      k<<<sms, warps * 32, 160 * 1024>>>(pat, out, clk);
      k<<<sms, warps * 32, 160 * 1024>>>(pat, out, clk);
      cudaDeviceSynchronize();
      long long h[256];
      cudaMemcpy(h, clk, sms * 8, cudaMemcpyDeviceToHost);
      double m = 0;
      for (int i = 0; i < sms; ++i) m += (double)h[i];
      printf("warps %2d  %-26s %6.2f SM-cycles per LDS.128 warp-instruction\n", warps, names[pat], m / sms / ITERS / (warps * 8));
    }
  }
  return 0;
}
