// Micro-benchmark: rate of the legacy warp-level tensor path on sm_100a, mma.sync.m16n8k8 TF32 (what a
// 3xTF32 filter bank / DCT would use), in MMA per clock per SM, for 4..16 resident warps per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_mma_tf32.bin mb_mma_tf32.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>

__global__ void __launch_bounds__(512, 1) kern(float* sink, long long* cyc, int iters) {
  const int lane = threadIdx.x & 31;
  unsigned a[4], b[2];
  float c[4][4];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + lane * 1e-3f + i);
  b[0] = __float_as_uint(0.5f + lane * 1e-3f);
  b[1] = __float_as_uint(0.25f);
  for (int u = 0; u < 4; ++u)
    for (int i = 0; i < 4; ++i) c[u][i] = 0.f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)  // 4 independent accumulator chains
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[u][0]), "+f"(c[u][1]), "+f"(c[u][2]), "+f"(c[u][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  long long t1 = clock64();
  float s = 0;
  for (int u = 0; u < 4; ++u)
    for (int i = 0; i < 4; ++i) s += c[u][i];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* sink; long long* dcyc;
  cudaMalloc(&sink, 4); cudaMalloc(&dcyc, sms * 8);
  const int iters = 20000;
  for (int warps : {4, 8, 16}) {
    kern<<<sms, warps * 32>>>(sink, dcyc, 64);
    kern<<<sms, warps * 32>>>(sink, dcyc, iters);
    cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), dcyc, sms * 8, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    double cyc = (double)h[sms / 2];
    double mma = (double)iters * 4 * warps;
    printf("%2d warps/SM: %.3f MMA(m16n8k8 tf32)/clk/SM  = %.0f FMA/clk/SM  (%.1f cycles per MMA per SM)\n", warps, mma / cyc,
           mma / cyc * 1024, cyc / mma);
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
