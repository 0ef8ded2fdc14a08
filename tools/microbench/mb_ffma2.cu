// Micro-benchmark: issue rate and pipe rate of packed FP32 (FFMA2/FADD2/FMUL2) on sm_100a,
// alone and mixed with scalar FP32, shared-memory loads and integer ALU work.
// One 1024-thread CTA per SM (8 warps per SMSP); thread 0 of each CTA reads clock64 around the loop.
// Output: warp-instructions per clock per SMSP for each instruction class.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_ffma2.bin mb_ffma2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define NCH 8
template <int MODE>
__global__ void __launch_bounds__(1024, 1) kern(float* sink, long long* cyc, int iters, float fa, float fc) {
  __shared__ float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3f;
  __syncthreads();
  float2 x[NCH];
  float y[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, 1.f + i); y[i] = 0.5f * i + threadIdx.x; }
  const float2 a2 = make_float2(fa, fa), c2 = make_float2(fc, fc);
  int ia = threadIdx.x, ib = 3;
  float ld = 0.f;
  const float* sp = sm + (threadIdx.x & 1023);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if constexpr (MODE == 0) {  // 16 scalar FFMA
#pragma unroll
        for (int i = 0; i < NCH; ++i) { x[i].x = __fmaf_rn(x[i].x, fa, fc); x[i].y = __fmaf_rn(x[i].y, fa, fc); }
      } else if constexpr (MODE == 1) {  // 8 FFMA2
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[i] = __ffma2_rn(x[i], a2, c2);
      } else if constexpr (MODE == 2) {  // 8 FADD2
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[i] = __fadd2_rn(x[i], c2);
      } else if constexpr (MODE == 3) {  // 8 FMUL2
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[i] = __fmul2_rn(x[i], a2);
      } else if constexpr (MODE == 4) {  // 8 FFMA2 with swapped + half-negated operand
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[i] = __ffma2_rn(make_float2(-x[i].y, x[i].x), a2, x[(i + 1) % NCH]);
      } else if constexpr (MODE == 5) {  // 8 FFMA2 + 8 scalar FFMA (independent)
#pragma unroll
        for (int i = 0; i < NCH; ++i) { x[i] = __ffma2_rn(x[i], a2, c2); y[i] = __fmaf_rn(y[i], fa, fc); }
      } else if constexpr (MODE == 6) {  // 8 FFMA2 + 8 LDS
#pragma unroll
        for (int i = 0; i < NCH; ++i) { x[i] = __ffma2_rn(x[i], a2, c2); ld += sp[((it + u) & 1) * 1024 + i * 32]; }
      } else if constexpr (MODE == 7) {  // 16 FFMA + 8 LDS
#pragma unroll
        for (int i = 0; i < NCH; ++i) { x[i].x = __fmaf_rn(x[i].x, fa, fc); x[i].y = __fmaf_rn(x[i].y, fa, fc); ld += sp[((it + u) & 1) * 1024 + i * 32]; }
      } else if constexpr (MODE == 8) {  // 8 FFMA2 + 8 integer ALU (LOP3)
#pragma unroll
        for (int i = 0; i < NCH; ++i) { x[i] = __ffma2_rn(x[i], a2, c2); ia = (ia ^ ib) & (ia | (ib + i)); }
      } else if constexpr (MODE == 9) {  // 16 scalar FADD
#pragma unroll
        for (int i = 0; i < NCH; ++i) { x[i].x = x[i].x + fc; x[i].y = x[i].y + fc; }
      } else if constexpr (MODE == 10) {  // 8 FFMA2 + 16 LDS (LDS-heavy)
#pragma unroll
        for (int i = 0; i < NCH; ++i) { x[i] = __ffma2_rn(x[i], a2, c2); ld += sp[((it + u) & 1) * 1024 + i * 32]; ld += sp[2048 + i * 32]; }
      } else if constexpr (MODE == 11) {  // 8 FFMA2 + 8 FADD scalar dependent on nothing (ld chain) : FADD only from LDS adds
#pragma unroll
        for (int i = 0; i < NCH; ++i) { x[i] = __ffma2_rn(x[i], a2, c2); y[i] = y[i] + fc; }
      }
    }
  }
  long long t1 = clock64();
  float s = ld + (float)ia;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += x[i].x + x[i].y + y[i];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Case { const char* name; int fp_packed, fp_scalar, other; };
static const Case cases[] = {
  {"16 FFMA (scalar)", 0, 16, 0},          {"8 FFMA2", 8, 0, 0},
  {"8 FADD2", 8, 0, 0},                    {"8 FMUL2", 8, 0, 0},
  {"8 FFMA2 swap+neg operand", 8, 0, 0},   {"8 FFMA2 + 8 FFMA", 8, 8, 0},
  {"8 FFMA2 + 8 LDS(+8 FADD)", 8, 8, 8},   {"16 FFMA + 8 LDS(+8 FADD)", 0, 24, 8},
  {"8 FFMA2 + 8x3 int ALU", 8, 0, 24},     {"16 FADD (scalar)", 0, 16, 0},
  {"8 FFMA2 + 16 LDS(+16 FADD)", 8, 16, 16}, {"8 FFMA2 + 8 FADD", 8, 8, 0},
};

template <int MODE>
void run(int sms, float* sink, long long* dcyc, int iters) {
  kern<MODE><<<sms, 1024>>>(sink, dcyc, 64, 0.999f, 1e-4f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<MODE><<<sms, 1024>>>(sink, dcyc, iters, 0.999f, 1e-4f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), dcyc, sms * 8, cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  const Case& c = cases[MODE];
  double cyc = (double)h[sms / 2];
  double per_smsp_iters = (double)iters * 4 * 8;  // 4 unrolled bodies x 8 warps per SMSP
  double n_inst = c.fp_packed + c.fp_scalar + c.other;
  double flops = (c.fp_packed * 2 + c.fp_scalar) * 2.0 * 32 * per_smsp_iters * 4 * sms;  // counts FADD as 2 too (slot-equiv)
  printf("%-30s cycles %10.0f  inst/clk/SMSP %.3f  fp-lane-ops/clk/SMSP %.2f  (%.3f ms, %.1f 'TFLOP/s' slot-equiv, clk %.0f MHz)\n",
         c.name, cyc, n_inst * per_smsp_iters / cyc, (c.fp_packed * 2 + c.fp_scalar) * 32.0 * per_smsp_iters / cyc, ms,
         flops / (ms * 1e-3) / 1e12, cyc / (ms * 1e-3) / 1e6);
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* sink; long long* dcyc; cudaMalloc(&sink, 4); cudaMalloc(&dcyc, sms * 8);
  const int iters = 20000;
  run<0>(sms, sink, dcyc, iters); run<1>(sms, sink, dcyc, iters); run<2>(sms, sink, dcyc, iters);
  run<3>(sms, sink, dcyc, iters); run<4>(sms, sink, dcyc, iters); run<5>(sms, sink, dcyc, iters);
  run<6>(sms, sink, dcyc, iters); run<7>(sms, sink, dcyc, iters); run<8>(sms, sink, dcyc, iters);
  run<9>(sms, sink, dcyc, iters); run<10>(sms, sink, dcyc, iters); run<11>(sms, sink, dcyc, iters);
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
