// Register-resident radix-2 DIT FFT building blocks with compile-time twiddles, written for the
// packed FP32 instructions of sm_100a (FFMA2 / FADD2 / FMUL2 on float2 = one complex number).
//
// Every index is a template/constexpr value, so after inlining the `float2 v[N]`
// arrays live entirely in registers and all twiddles are instruction immediates.
// Butterflies use the FMA form  a' = a + w*b,  b' = 2a - a'  (3 FFMA2 for a general
// twiddle, 2 FADD2 for w in {1, -i}, 3 for the sqrt(1/2) twiddles).
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace aad {

// ---------------------------------------------------------------- constexpr trig
constexpr double kPi = 3.141592653589793238462643383279502884;

constexpr double cx_sin_small(double x) {  // |x| <= pi/4
  double x2 = x * x, term = x, sum = x;
  for (int i = 1; i < 14; ++i) {
    term *= -x2 / double((2 * i) * (2 * i + 1));
    sum += term;
  }
  return sum;
}
constexpr double cx_cos_small(double x) {
  double x2 = x * x, term = 1.0, sum = 1.0;
  for (int i = 1; i < 14; ++i) {
    term *= -x2 / double((2 * i - 1) * (2 * i));
    sum += term;
  }
  return sum;
}
struct cx_cs {
  double c, s;
};
// cos/sin of 2*pi*num/den with exact quadrant symmetry (integer range reduction).
constexpr cx_cs cx_cossin_2pi(long long num, long long den) {
  long long r = num % den;
  if (r < 0) r += den;
  long long q = (4 * r) / den;          // quadrant
  long long rem = 4 * r - q * den;      // in [0, den): angle = (pi/2) * rem/den
  double c = 0, s = 0;
  if (rem == 0) {
    c = 1.0;
    s = 0.0;
  } else if (2 * rem <= den) {
    double a = (kPi / 2) * double(rem) / double(den);
    c = cx_cos_small(a);
    s = cx_sin_small(a);
  } else {
    double a = (kPi / 2) * double(den - rem) / double(den);
    c = cx_sin_small(a);
    s = cx_cos_small(a);
  }
  switch (q) {
    case 0: return {c, s};
    case 1: return {-s, c};
    case 2: return {-c, -s};
    default: return {s, -c};
  }
}

// ---------------------------------------------------------------- static loops
template <int I>
using ic = std::integral_constant<int, I>;

template <int B, int E, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (B < E) {
    f(ic<B>{});
    static_for<B + 1, E>(static_cast<F&&>(f));
  }
}

constexpr int bitrev(int x, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}
constexpr int ilog2(int x) {
  int r = 0;
  while ((1 << r) < x) ++r;
  return r;
}

// ---------------------------------------------------------------- packed FP32 helpers
// sm_100a executes FFMA2 / FADD2 / FMUL2 on float2 register pairs: one issue slot for two
// FP32 lane-ops (measured: same 128 lane-ops/clk/SM as scalar FFMA, half the issue slots;
// tools/microbench/mb_ffma2.cu).  Operand swap (.LO_HI), per-half negation and scalar broadcast
// (.F32) are instruction modifiers, so with a complex number held as one float2 the products
// by -i, conj() and i*w*z cost nothing extra: ptxas folds the make_float2() shuffles below.
__device__ __forceinline__ float2 pk_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 pk_sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 pk_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 pk_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
// a * s + c with the scalar s broadcast to both halves
__device__ __forceinline__ float2 pk_fma_s(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }

// complex product a * w  (2 packed instructions)
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  // a*w = w.x * (a.x, a.y) + w.y * (-a.y, a.x): swap/negate on the register pair, scalars broadcast
  const float2 t = __fmul2_rn(make_float2(-a.y, a.x), make_float2(w.y, w.y));
  return __ffma2_rn(a, make_float2(w.x, w.x), t);
}

// complex product a * exp(-2*pi*i*J/M) with a compile-time twiddle (2 packed instructions, immediates)
template <int J, int M>
__device__ __forceinline__ float2 cmul_const(float2 a) {
  if constexpr (J % M == 0) {
    return a;
  } else {
    constexpr cx_cs cs = cx_cossin_2pi(J, M);
    constexpr float wr = float(cs.c), wi = float(-cs.s);
    const float2 t = __fmul2_rn(make_float2(-a.y, a.x), make_float2(wi, wi));
    return __ffma2_rn(a, make_float2(wr, wr), t);
  }
}

// ---------------------------------------------------------------- butterflies
// W = exp(-2*pi*i*J/M) = (wr, wi) with wi = -sin.  (a, b) <- (a + W b, a - W b).
// Packed cost: 2 for W in {1, -i}, 3 otherwise (general: t = a + wr*b; a' = t + wi*(-b.y, b.x);
// b' = 2a - a').
template <int J, int M>
__device__ __forceinline__ void butterfly(float2& a, float2& b) {
  if constexpr (J == 0) {
    const float2 t = b;
    b = pk_sub(a, t);
    a = pk_add(a, t);
  } else if constexpr (4 * J == M) {  // w = -i : w*b = (b.y, -b.x)
    const float2 t = b;
    b = __fadd2_rn(a, make_float2(-t.y, t.x));
    a = __fadd2_rn(a, make_float2(t.y, -t.x));
  } else if constexpr (8 * J == M) {  // w = (1 - i)/sqrt2 : w*b = c*(b.x + b.y, b.y - b.x)
    constexpr float c = 0.70710678118654752440f;
    const float2 s = __fadd2_rn(b, make_float2(b.y, -b.x));
    b = pk_fma_s(s, -c, a);
    a = pk_fma_s(s, c, a);
  } else if constexpr (8 * J == 3 * M) {  // w = (-1 - i)/sqrt2 : w*b = -c*(b.x - b.y, b.x + b.y)
    constexpr float c = 0.70710678118654752440f;
    const float2 s = __fadd2_rn(b, make_float2(-b.y, b.x));
    b = pk_fma_s(s, c, a);
    a = pk_fma_s(s, -c, a);
  } else {
    constexpr cx_cs cs = cx_cossin_2pi(J, M);
    constexpr float wr = float(cs.c), wi = float(-cs.s);
    const float2 t = pk_fma_s(b, wr, a);
    const float2 na = __ffma2_rn(make_float2(-b.y, b.x), make_float2(wi, wi), t);
    b = __ffma2_rn(a, make_float2(2.0f, 2.0f), make_float2(-na.x, -na.y));
    a = na;
  }
}

// In-place DIT FFT of size N on v[OFF .. OFF+N).  Input must be stored in
// bit-reversed order (element n at OFF + bitrev(n)); output is in natural order.
// FIRST > 1 skips the leading stages (the caller has fused them with something else).
template <int N, int OFF, int NV, int FIRST = 1>
__device__ __forceinline__ void fft_dit(float2 (&v)[NV]) {
  constexpr int LOG2N = ilog2(N);
  static_for<FIRST, LOG2N + 1>([&](auto s_) {
    constexpr int S = decltype(s_)::value;
    constexpr int M = 1 << S;
    constexpr int H = M / 2;
    static_for<0, N / M>([&](auto g_) {
      constexpr int K = decltype(g_)::value * M;
      static_for<0, H>([&](auto j_) {
        constexpr int J = decltype(j_)::value;
        butterfly<J, M>(v[OFF + K + J], v[OFF + K + J + H]);
      });
    });
  });
}

}  // namespace aad
