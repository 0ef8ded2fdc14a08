// CQCC on the device: the sibling extractor of the reference's map that its CNN-BiLSTM is trained on
// (extract_cqcc, ASV_dl_func.py:442-481).
//
//   librosa.cqt (hop 512, C1 upwards, 12 bins per octave)  ->  |.|  ->  amplitude_to_db(ref=np.max)
//   ->  linear interpolation of every frame from the geometric CQT frequencies onto a uniform grid
//   ->  log(x^2 + 1e-12)  ->  DCT-II (ortho), first n_ceps rows.
//
// librosa's CQT is a recursion over octaves: the top octave's 12 wavelets are applied to the signal, the signal is
// halved in rate, the same 12 (rescaled) wavelets give the next octave, and so on.  librosa applies the wavelets as
// sparsified FFT bases to a rectangular-window STFT; the response of one bin to one frame is therefore a fixed
// linear functional of the frame's n_fft samples, and this file applies it in that form: the plan turns every
// sparsified basis row back into n_fft complex taps g_k[n] = sum_f B[k, f] exp(-2 pi i f n / n_fft) (in double, with
// the per-octave sqrt(sr / my_sr) and the final 1 / sqrt(length) scales folded in), and k_cqt_octave evaluates
// sum_n g_k[n] y[t hop - n_fft / 2 + n] directly.  The 2 -> 1 resampler is a 255-tap Kaiser-windowed sinc half band
// (librosa's 'soxr_hq' is a closed polyphase design; see oracle/cqcc_ref.py for what that means for parity).
//
// Kernels: k_cqt_resample (one octave down), k_cqt_octave (12 bins x frames of one octave, magnitudes + running
// utterance maximum), k_cqcc_epilogue (dB, interpolation, log, DCT).  Arithmetic is float32 like the reference's
// (complex64 CQT); tables are built in double.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/aad.h"

namespace {

constexpr int kHop = 512;
constexpr int kTaps = 255;                 // resampler: taps h[0 .. 254], centre 127, even offsets from it are zero
constexpr int kMaxOct = 12;
constexpr double kPi = 3.141592653589793238462643383279502884;
constexpr double kFminC1 = 32.70319566257483;  // librosa.note_to_hz('C1')

struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) err = cudaSetDevice(device);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ---------------------------------------------------------------------------------------------- kernels
__device__ __forceinline__ float load_sample(const void* wav, int i16, long long idx) {
  return i16 ? (float)__ldg(static_cast<const short*>(wav) + idx) * (1.0f / 32768.0f)
             : __ldg(static_cast<const float*>(wav) + idx);
}

// The resampler's taps are one fixed design (255-tap Kaiser half band): they live in constant memory so that the
// unrolled filter reads them through the uniform datapath (LDCU -> uniform-register operand of FFMA2), not through
// the shared-memory pipe.  Only the taps at odd offsets from the centre (and the centre) are non-zero:
// half[i] = h[2 i], i = 0 .. 127.  The kernel accumulates output PAIRS with packed FMAs, which need (half[i], half[i-1])
// as one aligned 64-bit operand: c_pa[t] = (R[2 t], R[2 t + 1]), c_pb[t] = (R[2 t + 1], R[2 t + 2]) over the reversed,
// zero-padded array R[j] = half[128 - j] (R[0] = R[129..] = 0), so that (half[i], half[i-1]) = (R[128 - i], R[129 - i]).
__constant__ float2 c_pa[66];
__constant__ float2 c_pb[66];
__constant__ float c_centre;

// y_out[b][n] = sqrt(2) * sum_k h[k] * y_in[b][2 n + k - 127],  n < ceil(len_in / 2); zero outside [0, len_in).
// Even k reach odd input offsets: with the inputs de-interleaved (sE[j] = in[base + 2 j + 1], sO[j] = in[base + 2 j + 2],
// base = 2 n0 - 128) output n0 + m is  sqrt(2) (sum_i half[i] sE[m + i] + c_centre sO[m + 63]).  A thread owns EIGHT
// consecutive outputs as four packed accumulators: every float4 it reads from sE feeds up to 16 FFMA2 whose tap pairs
// are uniform-register operands (16 outputs per thread was measured too: the tap pairs no longer fit the uniform
// register file and the uniform datapath saturates, 6.8 ms instead of 2.8 ms for the first stage).  sE is skewed by
// 4 words per 32 so that the stride-8 float4 reads of a quarter warp fall into 8 bank groups.  Rows of the octave buffers carry pad_in / pad_out zeros in front and n_fft zeros behind
// the signal (written here), so that every CQT frame of the lower octaves is an interior frame.
// grid (chunks of 1024 outputs, B), 128 threads.
#ifndef AAD_RS_PER
#define AAD_RS_PER 8
#endif
constexpr int kRsPer = AAD_RS_PER, kRsOut = 128 * kRsPer, kRsIn = kRsOut + 128 + 8;
// stride-8 float4 reads of a quarter warp need a skew to fall into 8 bank groups; stride 12 (48 bytes) does by itself
__device__ __forceinline__ int rs_skew(int j) { return kRsPer == 8 ? j + 4 * (j >> 5) : j; }
__global__ void __launch_bounds__(128) k_cqt_resample(const void* in, int in_i16, long long in_stride, const long long* row_off,
                                                      int pad_in, float* out, long long out_stride, int pad_out, int tail_out,
                                                      const int32_t* lengths, int shift) {
  __shared__ __align__(16) float sE[kRsIn + 4 * (kRsIn / 32) + 8];
  // sO is stored one position late and skewed like sE: the eight centre-tap inputs of a thread, sO[m0 + 63 .. m0 + 70],
  // are then two aligned float4 (as scalars the stride-8 reads of a warp were 8-way bank conflicts)
  __shared__ __align__(16) float sO[kRsIn + 4 * (kRsIn / 32) + 16];
  const int b = blockIdx.y;
  long long len0 = lengths[b];
  if (len0 < 0) len0 = 0;
  long long len_in = len0;
  for (int s = 0; s < shift; ++s) len_in = (len_in + 1) >> 1;  // length at the input octave
  const long long len_out = (len_in + 1) >> 1;
  const long long n0 = (long long)blockIdx.x * kRsOut;
  float* orow = out + (long long)b * out_stride;
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < pad_out; i += 128) orow[i] = 0.f;
  if (n0 >= len_out + tail_out) return;
  if (n0 >= len_out) {  // only zeros behind the signal in this chunk
    for (int i = threadIdx.x; i < kRsOut; i += 128)
      if (n0 + i < len_out + tail_out) orow[pad_out + n0 + i] = 0.f;
    return;
  }
  const long long base = 2 * n0 - (kTaps - 1) / 2 - 1;
  const long long irow = (row_off ? __ldg(row_off + b) : (long long)b * in_stride) + pad_in;
  // in[base + 4 q + {0, 1, 2, 3}] = sO[2 q - 1], sE[2 q], sO[2 q], sE[2 q + 1]: one aligned 16-byte load per four samples
  // when the row allows it (the library's own octave buffers always do), scalar masked loads otherwise
  const bool vec = ((irow + base) & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  for (int q = threadIdx.x; q <= kRsIn / 2; q += 128) {
    const long long s = base + 4 * q;
    float v[4];
    if (vec && s >= 0 && s + 4 <= len_in) {
      if (in_i16) {
        const short4 t = __ldg(reinterpret_cast<const short4*>(static_cast<const short*>(in) + irow + s));
        v[0] = (float)t.x * (1.0f / 32768.0f); v[1] = (float)t.y * (1.0f / 32768.0f);
        v[2] = (float)t.z * (1.0f / 32768.0f); v[3] = (float)t.w * (1.0f / 32768.0f);
      } else {
        const float4 t = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(in) + irow + s));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = (s + e >= 0 && s + e < len_in) ? load_sample(in, in_i16, irow + s + e) : 0.f;
    }
    if (q > 0) sO[rs_skew(2 * q)] = v[0];                 // sO[2 q - 1], stored one late
    if (2 * q < kRsIn) {
      sE[rs_skew(2 * q)] = v[1];
      sO[rs_skew(2 * q + 1)] = v[2];                       // sO[2 q]
    }
    if (2 * q + 1 < kRsIn) sE[rs_skew(2 * q + 1)] = v[3];
  }
  __syncthreads();
  const int m0 = kRsPer * threadIdx.x;
  float2 acc[kRsPer / 2];
#pragma unroll
  for (int p = 0; p < kRsPer / 2; ++p) acc[p] = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < (128 + kRsPer + 2) / 4; ++r) {
    const float4 v = *reinterpret_cast<const float4*>(sE + rs_skew(m0 + 4 * r));
    const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int q = 4 * r + e;  // sE index relative to m0: feeds outputs (2 p, 2 p + 1) with taps (i, i - 1), i = q - 2 p
#pragma unroll
      for (int p = 0; p < kRsPer / 2; ++p) {
        const int i = q - 2 * p;
        if (i >= 0 && i <= 128) {
          const int sidx = 128 - i;  // (half[i], half[i - 1]) = (R[sidx], R[sidx + 1])
          const float2 tp = (sidx & 1) ? c_pb[sidx >> 1] : c_pa[sidx >> 1];
          acc[p] = __ffma2_rn(make_float2(x[e], x[e]), tp, acc[p]);
        }
      }
    }
  }
  float res[kRsPer];
  {
    float cen[kRsPer];   // sO[m0 + 63 + m]
#pragma unroll
    for (int q = 0; q < kRsPer / 4; ++q) {
      const float4 c = *reinterpret_cast<const float4*>(sO + rs_skew(m0 + 64 + 4 * q));
      cen[4 * q] = c.x; cen[4 * q + 1] = c.y; cen[4 * q + 2] = c.z; cen[4 * q + 3] = c.w;
    }
#pragma unroll
    for (int m = 0; m < kRsPer; ++m) {
      const float a = (m & 1) ? acc[m >> 1].y : acc[m >> 1].x;
      res[m] = 1.41421356237309515f * __fmaf_rn(c_centre, cen[m], a);
    }
  }
  const long long nb = n0 + m0;
  float* op = orow + pad_out + nb;
  if (nb + kRsPer <= len_out && ((pad_out | out_stride) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < kRsPer / 4; ++q) *reinterpret_cast<float4*>(op + 4 * q) = make_float4(res[4 * q], res[4 * q + 1], res[4 * q + 2], res[4 * q + 3]);
  } else {
#pragma unroll
    for (int m = 0; m < kRsPer; ++m) {
      if (nb + m < len_out) op[m] = res[m];
      else if (nb + m < len_out + tail_out) op[m] = 0.f;
    }
  }
}

// One octave: mag[b][bin0 + k][t] = | sum_n g[k][n] * y[b][pad + t * hop - n_fft / 2 + n] |  for the n_k bins of the octave.
// Thread = (group of 4 bins, group of 4 frames): per 4 samples four 16-byte frame loads and sixteen 8-byte tap loads
// feed 64 FFMA2 (one bin per thread made the kernel L1-bound at 98 %: the 12 threads of a frame group repeated its
// loads).  The taps are staged in shared memory once per CTA; the frames are read where they lie -- the octave
// buffers are padded (pad = n_fft / 2 zeros in front, n_fft behind), so every frame of octaves >= 1 is interior; in
// octave 0 (the caller's waveform) the first and the last frames of an utterance take the masked path.
// A CTA covers kOctUtt utterances x 64 frames: (n_k / 4) x 16 threads per utterance.  grid (frame blocks of 64, B / kOctUtt).
constexpr int kOctFrames = 64, kOctUtt = 4, kOctBins = 4;
__global__ void __launch_bounds__(kOctUtt * 16 * 6) k_cqt_octave(const void* y, int y_i16, long long y_stride,
                                                                 const long long* row_off, int pad, const int32_t* lengths,
                                                                 int B, int shift, int hop, int n_fft,
                                                                 const float2* g, int n_k, int bin0, float* mag,
                                                                 long long mag_stride_b, int t_alloc, int32_t* utt_max) {
  extern __shared__ __align__(16) float smem[];
  float2* sG = reinterpret_cast<float2*>(smem);                  // [n_fft][n_kp], n_kp = n_k rounded up to 4 (zero taps)
  const int n_bg = (n_k + kOctBins - 1) / kOctBins, n_kp = n_bg * kOctBins;
  for (int i = threadIdx.x; i < n_fft * n_kp; i += blockDim.x) {
    const int n = i / n_kp, k = i - n * n_kp;
    sG[i] = k < n_k ? __ldg(g + (size_t)k * n_fft + n) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const int per_utt = n_bg * 16;
  const int u = threadIdx.x / per_utt, r = threadIdx.x - u * per_utt;
  const int b = blockIdx.y * kOctUtt + u;
  if (u >= kOctUtt || b >= B) return;
  const long long len0 = lengths[b];
  if (len0 <= 0) return;
  const int T = (int)min((long long)t_alloc, 1 + len0 / kHop);  // frames of the utterance (all octaves are trimmed to it)
  const int kg = r % n_bg, fg = r / n_bg;                        // bin group, frame group (< 16)
  const int tb = blockIdx.x * kOctFrames + 4 * fg, k0 = kOctBins * kg;
  if (tb >= T) return;
  long long len = len0;
  for (int s = 0; s < shift; ++s) len = (len + 1) >> 1;
  float2 acc[4][kOctBins];
  long long s0[4];  // first sample of each of the thread's frames, relative to the signal's first sample
  const long long row = (row_off ? __ldg(row_off + b) : (long long)b * y_stride) + pad;
  const bool vec_ok = !y_i16 && (row & 3) == 0 && (hop & 3) == 0 && (n_fft & 7) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) & 15) == 0;
  bool fast = vec_ok;
#pragma unroll
  for (int f = 0; f < 4; ++f) {
#pragma unroll
    for (int c = 0; c < kOctBins; ++c) acc[f][c] = make_float2(0.f, 0.f);
    s0[f] = (long long)(tb + f) * hop - n_fft / 2;
    if (tb + f >= T) s0[f] = pad > 0 ? 0 : (max(0ll, len - n_fft) & ~3ll);  // a frame that is not stored: read anything valid (aligned)
    if (pad == 0) fast = fast && s0[f] >= 0 && s0[f] + n_fft <= len;
  }
  const float2* gk = sG + k0;
  for (int n = 0; n < n_fft; n += 4) {
    float v[4][4];
    if (fast) {
      const float* yf = static_cast<const float*>(y) + row;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(yf + s0[f] + n));
        v[f][0] = q.x; v[f][1] = q.y; v[f][2] = q.z; v[f][3] = q.w;
      }
    } else {  // frames touching the zero padding of the caller's waveform, int16 input, unaligned rows
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const long long src = s0[f] + n;
        if (vec_ok && src >= 0 && src + 4 <= len) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(y) + row + src));
          v[f][0] = q.x; v[f][1] = q.y; v[f][2] = q.z; v[f][3] = q.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[f][e] = (src + e >= 0 && src + e < len) ? load_sample(y, y_i16, row + src + e) : 0.f;
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 gg[kOctBins];
#pragma unroll
      for (int c = 0; c < kOctBins; c += 2) {  // two bins per 16-byte load
        const float4 q = *reinterpret_cast<const float4*>(gk + (n + e) * n_kp + c);
        gg[c] = make_float2(q.x, q.y);
        gg[c + 1] = make_float2(q.z, q.w);
      }
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int c = 0; c < kOctBins; ++c) acc[f][c] = __ffma2_rn(make_float2(v[f][e], v[f][e]), gg[c], acc[f][c]);
    }
  }
  float vmax = 0.f;
  bool poison = false;
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    if (tb + f < T) {
#pragma unroll
      for (int c = 0; c < kOctBins; ++c) {
        if (k0 + c < n_k) {
          const float m = sqrtf(__fmaf_rn(acc[f][c].x, acc[f][c].x, acc[f][c].y * acc[f][c].y));
          mag[(long long)b * mag_stride_b + (long long)(bin0 + k0 + c) * t_alloc + tb + f] = m;
          if (m == m) vmax = fmaxf(vmax, m);
          else poison = true;
        }
      }
    }
  }
  // non-negative floats order like their bit patterns; a NaN poisons the utterance
  atomicMax(utt_max + b, poison ? 0x7fc00000 : __float_as_int(vmax));
}

// ---- the same octave on the tensor cores ------------------------------------------------------------------------
// D[frame][2 bin + {re, im}] = sum_n y[frame start + n] * G[n][2 bin + {re, im}] is a GEMM with M = frames, K = n_fft,
// N = 24: mma.sync.m16n8k8 TF32 with the 3-term split (a_hi b_hi + a_lo b_hi + a_hi b_lo: float32-accurate, the same
// scheme as the DCT of k_cepstra).  At ~12 issue-blocking cycles per HMMA the 9 MMAs of a k-step still beat the 96
// packed FMAs they replace by 3x.  A warp owns 16 consecutive frames of one utterance and reads them where they lie
// in the padded octave buffer: lane (g, t) loads 16-byte pieces of rows g and g + 8; the k order inside a 32-tap
// chunk is permuted so that those pieces ARE its A fragments (k-step j of the chunk: column t = tap 4 t + j, column
// t + 4 = tap 16 + 4 t + j) -- the plan stores the B fragments (taps, pre-split into hi / lo) in the matching order,
// one LDS.128 per (k-step, n-tile).  The C fragment holds (re, im) of one bin side by side: the magnitude needs no
// exchange.  Octaves whose buffers allow aligned 16-byte frame reads (hop % 4 == 0; not the caller's own waveform)
// take this kernel, the others k_cqt_octave.
constexpr int kMmaWarps = 8, kMmaNT = 3;   // n-tiles of 8: up to 12 bins per octave
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// GENERIC = false: padded float octave buffer of the library (aligned 16-byte reads, every frame interior).
// GENERIC = true: the caller's waveform (octave 0): float or int16, any row offset, frames that reach in front of /
// behind the signal read zeros -- scalar loads, masked only in the tiles that touch an end.
template <bool GENERIC>
__global__ void __launch_bounds__(kMmaWarps * 32, 3) k_cqt_octave_mma(const void* __restrict__ yv, int y_i16, long long y_stride,
                                                                   const long long* __restrict__ row_off, int pad,
                                                                   const int32_t* __restrict__ lengths, int B, int hop, int n_fft,
                                                                   const float4* __restrict__ gfrag, int n_k, int bin0,
                                                                   float* __restrict__ mag, long long mag_stride_b, int t_alloc,
                                                                   int32_t* utt_max, int mt_per_utt) {
  const float* y = static_cast<const float*>(yv);
  extern __shared__ __align__(16) float smem[];
  float4* sB = reinterpret_cast<float4*>(smem);  // [n_fft / 8][kMmaNT][32] = {b0 hi, b1 hi, b0 lo, b1 lo}
  const int n_frag = n_fft / 8 * kMmaNT * 32;
  for (int i = threadIdx.x; i < n_frag; i += blockDim.x) sB[i] = __ldg(gfrag + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
  const long long n_items = (long long)B * mt_per_utt;
  for (long long item = (long long)blockIdx.x * kMmaWarps + warp; item < n_items; item += (long long)gridDim.x * kMmaWarps) {
    const int b = (int)(item / mt_per_utt), t0 = 16 * (int)(item - (long long)b * mt_per_utt);
    const long long len0 = __ldg(lengths + b);
    if (len0 <= 0) continue;
    const int T = (int)min((long long)t_alloc, 1 + len0 / kHop);
    if (t0 >= T) continue;
    const int ta = t0 + g, tb = t0 + g + 8;
    const long long row0 = GENERIC ? (row_off ? __ldg(row_off + b) : (long long)b * y_stride) : (long long)b * y_stride + pad;
    // first sample of this lane's pieces of rows g and g + 8, relative to the signal's first sample (rows that are not
    // stored re-read the last frame)
    const long long sa = (long long)min(ta, T - 1) * hop - n_fft / 2 + 4 * tig;
    const long long sb = (long long)min(tb, T - 1) * hop - n_fft / 2 + 4 * tig;
    const float* pa = y + row0 + sa;
    const float* pb = y + row0 + sb;
    // GENERIC: does any frame of the tile reach outside [0, len0)?  (warp-uniform)
    const bool edge = GENERIC && ((long long)t0 * hop - n_fft / 2 < 0 || (long long)min(t0 + 15, T - 1) * hop + n_fft / 2 > len0);
    const bool vec4 = GENERIC && !edge && (row0 & 3) == 0 && (hop & 3) == 0 && (n_fft & 7) == 0 &&
                      (reinterpret_cast<uintptr_t>(yv) & 15) == 0;
    float acc[kMmaNT][4];
#pragma unroll
    for (int nt = 0; nt < kMmaNT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    const float4* bp = sB + lane;
    for (int c = 0; c < n_fft; c += 32, bp += 4 * kMmaNT * 32) {
      float xa0[4], xa1[4], xb0[4], xb1[4];
      if constexpr (GENERIC) {
        // interior tiles of rows that start on a 4-sample boundary: one 16-byte (float) or 8-byte (int16) load per piece
        auto piece = [&](long long s, float (&x)[4]) {
          if (vec4) {
            if (y_i16) {
              const short4 q = __ldg(reinterpret_cast<const short4*>(static_cast<const short*>(yv) + row0 + s));
              x[0] = (float)q.x * (1.0f / 32768.0f); x[1] = (float)q.y * (1.0f / 32768.0f);
              x[2] = (float)q.z * (1.0f / 32768.0f); x[3] = (float)q.w * (1.0f / 32768.0f);
            } else {
              const float4 q = __ldg(reinterpret_cast<const float4*>(y + row0 + s));
              x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              x[e] = (!edge || (s + e >= 0 && s + e < len0)) ? load_sample(yv, y_i16, row0 + s + e) : 0.f;
          }
        };
        piece(sa + c, xa0); piece(sa + c + 16, xa1); piece(sb + c, xb0); piece(sb + c + 16, xb1);
      } else {
        const float4 qa0 = __ldg(reinterpret_cast<const float4*>(pa + c)), qa1 = __ldg(reinterpret_cast<const float4*>(pa + c + 16));
        const float4 qb0 = __ldg(reinterpret_cast<const float4*>(pb + c)), qb1 = __ldg(reinterpret_cast<const float4*>(pb + c + 16));
        xa0[0] = qa0.x; xa0[1] = qa0.y; xa0[2] = qa0.z; xa0[3] = qa0.w; xa1[0] = qa1.x; xa1[1] = qa1.y; xa1[2] = qa1.z; xa1[3] = qa1.w;
        xb0[0] = qb0.x; xb0[1] = qb0.y; xb0[2] = qb0.z; xb0[3] = qb0.w; xb1[0] = qb1.x; xb1[1] = qb1.y; xb1[2] = qb1.z; xb1[3] = qb1.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float av[4] = {xa0[j], xb0[j], xa1[j], xb1[j]};  // (g, t), (g + 8, t), (g, t + 4), (g + 8, t + 4)
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          hi[i] = tf32_rna(av[i]);
          lo[i] = __float_as_uint(av[i] - __uint_as_float(hi[i]));
        }
        // the three terms of an accumulator are dependent MMAs: issue them n-tile by n-tile per term, so that two
        // independent MMAs sit between an MMA and the next one into the same accumulator
        float4 bf[kMmaNT];
#pragma unroll
        for (int nt = 0; nt < kMmaNT; ++nt) bf[nt] = bp[(j * kMmaNT + nt) * 32];
#pragma unroll
        for (int nt = 0; nt < kMmaNT; ++nt) mma_tf32(acc[nt], lo, __float_as_uint(bf[nt].x), __float_as_uint(bf[nt].y));
#pragma unroll
        for (int nt = 0; nt < kMmaNT; ++nt) mma_tf32(acc[nt], hi, __float_as_uint(bf[nt].z), __float_as_uint(bf[nt].w));
#pragma unroll
        for (int nt = 0; nt < kMmaNT; ++nt) mma_tf32(acc[nt], hi, __float_as_uint(bf[nt].x), __float_as_uint(bf[nt].y));
      }
    }
    float vmax = 0.f;
    bool poison = false;
    float* mrow = mag + (long long)b * mag_stride_b + (long long)bin0 * t_alloc;
#pragma unroll
    for (int nt = 0; nt < kMmaNT; ++nt) {
      const int bin = 4 * nt + tig;
      if (bin < n_k) {
        const float m0 = sqrtf(__fmaf_rn(acc[nt][0], acc[nt][0], acc[nt][1] * acc[nt][1]));
        const float m1 = sqrtf(__fmaf_rn(acc[nt][2], acc[nt][2], acc[nt][3] * acc[nt][3]));
        if (ta < T) {
          mrow[(long long)bin * t_alloc + ta] = m0;
          if (m0 == m0) vmax = fmaxf(vmax, m0);
          else poison = true;
        }
        if (tb < T) {
          mrow[(long long)bin * t_alloc + tb] = m1;
          if (m1 == m1) vmax = fmaxf(vmax, m1);
          else poison = true;
        }
      }
    }
    int enc = poison ? 0x7fc00000 : __float_as_int(vmax);   // non-negative floats order like their bit patterns
    enc = __reduce_max_sync(0xffffffffu, enc);
    if (lane == 0) atomicMax(utt_max + b, enc);
  }
}

// dB relative to the utterance maximum (floor -80) -> interpolation onto the uniform frequency grid -> log(x^2 + 1e-12)
// -> DCT-II ortho.  grid (frame blocks of 32, B), 256 threads.
constexpr int kEpiFrames = 32;
__global__ void __launch_bounds__(256) k_cqcc_epilogue(const float* mag, long long mag_stride_b, int t_alloc, int n_bins,
                                                       const int32_t* lengths, const int32_t* utt_max, const int32_t* interp_lo,
                                                       const float* interp_w, const float* dct, int n_ceps, float* out,
                                                       long long out_stride_b, int32_t* n_frames, int32_t* status) {
  extern __shared__ float smem[];
  const int w0 = max(kEpiFrames + 1, (n_ceps + 3) & ~3);   // the first region later holds the transposed DCT matrix
  float* sDb = smem;                                // [n_bins][33]
  float* sLp = smem + n_bins * w0;                  // [n_bins][33]
  const int b = blockIdx.y;
  const long long len0 = lengths[b];
  int T = len0 > 0 ? (int)(1 + len0 / kHop) : 0;
  int st = len0 > 0 ? 0 : 1;
  if (T > t_alloc) {
    st = 4;
    T = 0;
  }
  const int mx = utt_max[b];
  if (st == 0 && mx >= 0x7f800000) st = 5;          // NaN / Inf in the audio
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    n_frames[b] = len0 > 0 ? (int)(1 + len0 / kHop) : 0;
    status[b] = st;
  }
  if (st != 0) return;
  const int t0 = blockIdx.x * kEpiFrames;
  if (t0 >= T) return;
  const int nt = min(kEpiFrames, T - t0);
  // librosa.amplitude_to_db(S, ref=np.max): power_to_db(S^2, ref=max^2, amin=1e-10, top_db=80), float32
  const float amin2 = 1e-5f * 1e-5f;
  const float vmax = __int_as_float(mx);
  // 10 log10(x) = 3.0103 log2(x), ln(x) = 0.6931 log2(x) with MUFU.LG2 (2^-22 absolute error in log2: 7e-7 dB, below the
  // float32 rounding of the dB values themselves); the reference cell is exact: the same expression of the same number
  const float kDb = 3.01029995663981195f, kLn = 0.69314718055994531f;
  const float ref_db = kDb * __log2f(fmaxf(amin2, vmax * vmax));
  const float* mb = mag + (long long)b * mag_stride_b;
  for (int i = threadIdx.x; i < n_bins * kEpiFrames; i += 256) {
    const int k = i / kEpiFrames, f = i - k * kEpiFrames;
    float v = 0.f;
    if (f < nt) {
      const float m = mb[(long long)k * t_alloc + t0 + f];
      v = fmaxf(kDb * __log2f(fmaxf(amin2, m * m)) - ref_db, -80.0f);   // the maximum of log_spec is 0 (the reference)
    }
    sDb[k * (kEpiFrames + 1) + f] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_bins * kEpiFrames; i += 256) {
    const int k = i / kEpiFrames, f = i - k * kEpiFrames;
    const int lo = interp_lo[k];
    const float x0 = sDb[lo * (kEpiFrames + 1) + f], x1 = sDb[(lo + 1) * (kEpiFrames + 1) + f];
    const float x = __fmaf_rn(interp_w[k], x1 - x0, x0);
    sLp[k * (kEpiFrames + 1) + f] = kLn * __log2f(__fmaf_rn(x, x, 1e-12f));
  }
  __syncthreads();
  // DCT: a thread owns one frame and four coefficients: per bin one LDS of the frame's value and one broadcast LDS.128
  // of the four DCT entries (sDb is free again: it holds the transposed matrix [bin][n_c4], n_c4 = n_ceps rounded up to 4)
  const int n_c4 = (n_ceps + 3) & ~3;
  float* sDt = sDb;
  __syncthreads();
  for (int i = threadIdx.x; i < n_bins * n_c4; i += 256) {
    const int k = i / n_c4, c = i - k * n_c4;
    sDt[i] = c < n_ceps ? __ldg(dct + (size_t)c * n_bins + k) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (n_c4 / 4) * kEpiFrames; i += 256) {
    const int cg = i / kEpiFrames, f = i - cg * kEpiFrames;
    if (f >= nt) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* dt = reinterpret_cast<const float4*>(sDt) + cg;
    for (int k = 0; k < n_bins; ++k) {
      const float x = sLp[k * (kEpiFrames + 1) + f];
      const float4 d = dt[k * (n_c4 / 4)];
      acc.x = __fmaf_rn(d.x, x, acc.x); acc.y = __fmaf_rn(d.y, x, acc.y);
      acc.z = __fmaf_rn(d.z, x, acc.z); acc.w = __fmaf_rn(d.w, x, acc.w);
    }
    float* o = out + (long long)b * out_stride_b + (long long)(4 * cg) * t_alloc + t0 + f;
    const float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * cg + j < n_ceps) o[(long long)j * t_alloc] = r[j];
  }
}

// ---------------------------------------------------------------------------------------------- plan (host tables)
std::vector<double> float_window_hann(double n) {  // librosa.filters.__float_window('hann')
  const int n_min = (int)std::floor(n), n_max = (int)std::ceil(n);
  std::vector<double> w(n_max, 0.0);
  for (int i = 0; i < n_min; ++i) w[i] = 0.5 - 0.5 * std::cos(2.0 * kPi * i / n_min);  // periodic hann of n_min samples
  return w;
}

}  // namespace

struct aad_cqcc_plan {
  int device = 0, sample_rate = 0, bpo = 12, n_ceps = 19, n_bins = 0, n_oct = 0;
  int sm_count = 148, mma_ctas = 1;  // resident CTAs of k_cqt_octave_mma per SM
  int n_fft = 0;                    // per-octave transform size (the same for every octave: wavelet lengths repeat)
  int n_k[kMaxOct] = {0};           // bins of octave i (top first)
  int bin0[kMaxOct] = {0};
  float2* d_g[kMaxOct] = {nullptr}; // [n_k][n_fft] taps of octave i
  float4* d_gfrag[kMaxOct] = {nullptr};  // the same taps as pre-split TF32 B fragments (k_cqt_octave_mma), or null
  int32_t* d_interp_lo = nullptr;
  float* d_interp_w = nullptr;
  float* d_dct = nullptr;           // [n_ceps][n_bins]
  std::vector<float> h_freqs;
};

extern "C" {

int aad_cqcc_plan_destroy(aad_cqcc_plan* pl) {
  if (!pl) return AAD_OK;
  DeviceGuard guard(pl->device);
  for (auto& p : pl->d_g) cudaFree(p);
  for (auto& p : pl->d_gfrag) cudaFree(p);
  cudaFree(pl->d_interp_lo);
  cudaFree(pl->d_interp_w);
  cudaFree(pl->d_dct);
  delete pl;
  return AAD_OK;
}

int aad_cqcc_plan_create(int sample_rate, int bins_per_octave, int n_ceps, int device, aad_cqcc_plan** out) {
  if (!out || sample_rate < 1000 || bins_per_octave < 1 || bins_per_octave > 24 || n_ceps < 1) return AAD_ERR_INVALID_ARG;
  const double sr = sample_rate, fmin = kFminC1, fmax = sr / 2 - 100;
  if (fmax <= fmin) return AAD_ERR_INVALID_ARG;
  const int bpo = bins_per_octave;
  const int n_bins = (int)(std::floor(std::log2(fmax / fmin)) * bpo);  // ASV_dl_func.py:455
  if (n_bins < 2 || n_ceps > n_bins) return AAD_ERR_INVALID_ARG;
  const int n_oct = (n_bins + bpo - 1) / bpo, n_filters = std::min(bpo, n_bins);
  if (n_oct > kMaxOct || (kHop >> (n_oct - 1)) < 1 || (kHop % (1 << (n_oct - 1))) != 0) return AAD_ERR_UNSUPPORTED;
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return AAD_ERR_CUDA;
  aad_cqcc_plan* pl = new (std::nothrow) aad_cqcc_plan();
  if (!pl) return AAD_ERR_INVALID_ARG;
  pl->device = device; pl->sample_rate = sample_rate; pl->bpo = bpo; pl->n_ceps = n_ceps; pl->n_bins = n_bins; pl->n_oct = n_oct;
  std::vector<double> freqs(n_bins), lengths(n_bins);
  const double r = std::pow(2.0, 2.0 / bpo), alpha = (r - 1) / (r + 1), Q = 1.0 / alpha;
  double cutoff = 0;
  for (int k = 0; k < n_bins; ++k) {
    freqs[k] = fmin * std::pow(2.0, (double)k / bpo);
    lengths[k] = Q * sr / freqs[k];
    cutoff = std::max(cutoff, freqs[k] * (1 + 0.5 * 1.50018310546875 / Q));
  }
  if (cutoff > sr / 2) {
    delete pl;
    return AAD_ERR_UNSUPPORTED;
  }
  pl->h_freqs.assign(freqs.begin(), freqs.end());
  cudaError_t e = cudaSuccess;
  // per octave (top first): wavelets at my_sr -> padded basis -> FFT -> sparsify -> back to n_fft taps
  double my_sr = sr;
  for (int i = 0; i < n_oct && e == cudaSuccess; ++i) {
    const int hi = n_bins - n_filters * i, lo = std::max(0, hi - n_filters);  // freqs[sl]
    const int nk = hi - lo;
    pl->n_k[i] = nk;
    pl->bin0[i] = lo;
    std::vector<double> len_oct(nk);
    double max_len = 0;
    for (int k = 0; k < nk; ++k) {
      len_oct[k] = Q * my_sr / freqs[lo + k];
      max_len = std::max(max_len, len_oct[k]);
    }
    const int n_fft = 1 << (int)std::ceil(std::log2(max_len));
    if (i == 0) pl->n_fft = n_fft;
    if (n_fft != pl->n_fft || n_fft > 2048) {  // every octave sees the same lengths in its own samples
      aad_cqcc_plan_destroy(pl);
      return AAD_ERR_UNSUPPORTED;
    }
    const int K = n_fft / 2 + 1;
    std::vector<float2> g((size_t)nk * n_fft);
    for (int k = 0; k < nk; ++k) {
      const double ilen = len_oct[k], f = freqs[lo + k];
      const long long n_lo = (long long)std::floor(-ilen / 2), n_hi = (long long)std::floor(ilen / 2);  // arange(-l//2, l//2)
      const int L = (int)(n_hi - n_lo);
      std::vector<double> win = float_window_hann((double)L);
      std::vector<std::complex<double>> sig(L);
      double l1 = 0;
      for (int j = 0; j < L; ++j) {
        const double ang = 2.0 * kPi * f * (double)(n_lo + j) / my_sr;
        sig[j] = std::complex<double>(std::cos(ang), std::sin(ang)) * win[j];
        l1 += std::abs(sig[j]);
      }
      std::vector<std::complex<double>> basis(n_fft, 0.0);
      const int lpad = (n_fft - L) / 2;
      for (int j = 0; j < L; ++j) basis[lpad + j] = sig[j] / l1 * (ilen / n_fft);  // normalize(norm=1); *= lengths / n_fft
      // fft (positive half) and sparsify_rows(quantile = 0.01)
      std::vector<std::complex<double>> B(K);
      std::vector<double> mags(K);
      double norm = 0;
      for (int q = 0; q < K; ++q) {
        std::complex<double> acc = 0;
        for (int n = 0; n < n_fft; ++n) {
          const double ang = -2.0 * kPi * (double)((long long)q * n % n_fft) / n_fft;
          acc += basis[n] * std::complex<double>(std::cos(ang), std::sin(ang));
        }
        B[q] = acc;
        mags[q] = std::abs(acc);
        norm += mags[q];
      }
      std::vector<double> srt(mags);
      std::sort(srt.begin(), srt.end());
      double cum = 0, thr = srt.back();
      for (int q = 0; q < K; ++q) {
        cum += srt[q] / norm;
        if (!(cum < 0.01)) {
          thr = srt[q];
          break;
        }
      }
      // librosa stores the sparsified basis as complex64 and rescales it by sqrt(sr / my_sr); 1 / sqrt(length at sr) is
      // vqt's final scale
      const double scale = std::sqrt(sr / my_sr) / std::sqrt(lengths[lo + k]);
      for (int q = 0; q < K; ++q) {
        if (mags[q] < thr) B[q] = 0;
        else B[q] = std::complex<double>((double)(float)B[q].real(), (double)(float)B[q].imag());
      }
      for (int n = 0; n < n_fft; ++n) {
        std::complex<double> acc = 0;
        for (int q = 0; q < K; ++q) {
          if (B[q] == std::complex<double>(0, 0)) continue;
          const double ang = -2.0 * kPi * (double)((long long)q * n % n_fft) / n_fft;
          acc += B[q] * std::complex<double>(std::cos(ang), std::sin(ang));
        }
        acc *= scale;
        g[(size_t)k * n_fft + n] = make_float2((float)acc.real(), (float)acc.imag());
      }
    }
    e = cudaMalloc((void**)&pl->d_g[i], g.size() * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_g[i], g.data(), g.size() * sizeof(float2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nk <= 4 * kMmaNT && n_fft % 32 == 0) {
      // B fragments of k_cqt_octave_mma: [k-step][n-tile][lane = 4 gid + tig] = {b0 hi, b1 hi, b0 lo, b1 lo} with
      // b0 = G[tap(tig)][8 nt + gid], b1 = G[tap(tig + 4)][8 nt + gid]; k-step ks = 4 c + j covers, in column order,
      // taps 32 c + 4 tig + j (columns 0..3) and 32 c + 16 + 4 tig + j (columns 4..7); G[n][2 bin + {0, 1}] = {re, im}
      auto tf32 = [](float x) {
        uint32_t u;
        std::memcpy(&u, &x, 4);
        u = (u + 0x1000u) & 0xffffe000u;   // cvt.rna: nearest, ties away from zero (magnitude bits)
        float r;
        std::memcpy(&r, &u, 4);
        return r;
      };
      auto G = [&](int tap, int n) {
        const int bin = n >> 1;
        if (bin >= nk) return 0.f;
        const float2 v = g[(size_t)bin * n_fft + tap];
        return (n & 1) ? v.y : v.x;
      };
      std::vector<float4> frag((size_t)n_fft / 8 * kMmaNT * 32);
      for (int ks = 0; ks < n_fft / 8; ++ks)
        for (int nt = 0; nt < kMmaNT; ++nt)
          for (int lane = 0; lane < 32; ++lane) {
            const int gid = lane >> 2, tig = lane & 3, c = ks / 4, j = ks % 4;
            const float b0 = G(32 * c + 4 * tig + j, 8 * nt + gid), b1 = G(32 * c + 16 + 4 * tig + j, 8 * nt + gid);
            const float h0 = tf32(b0), h1 = tf32(b1);
            frag[((size_t)ks * kMmaNT + nt) * 32 + lane] = make_float4(h0, h1, b0 - h0, b1 - h1);
          }
      e = cudaMalloc((void**)&pl->d_gfrag[i], frag.size() * sizeof(float4));
      if (e == cudaSuccess) e = cudaMemcpy(pl->d_gfrag[i], frag.data(), frag.size() * sizeof(float4), cudaMemcpyHostToDevice);
    }
    my_sr /= 2.0;
  }
  // resampler taps: 0.5 sinc(n / 2) kaiser(255, 14), unit DC gain
  {
    std::vector<double> h(kTaps);
    auto bessel_i0 = [](double x) {
      double s = 1, t = 1;
      for (int k = 1; k < 60; ++k) {
        t *= (x / (2 * k)) * (x / (2 * k));
        s += t;
      }
      return s;
    };
    double sum = 0;
    for (int i = 0; i < kTaps; ++i) {
      const double n = i - (kTaps - 1) / 2.0, x = 0.5 * n;
      const double sinc = n == 0 ? 1.0 : std::sin(kPi * x) / (kPi * x);
      const double rr = 2.0 * i / (kTaps - 1) - 1.0;
      h[i] = 0.5 * sinc * bessel_i0(14.0 * std::sqrt(std::max(0.0, 1 - rr * rr))) / bessel_i0(14.0);
      sum += h[i];
    }
    // R[j] = half[128 - j] = h[2 (128 - j)] / sum for j = 1 .. 128, zero elsewhere
    std::vector<float> R(134, 0.f);
    for (int j = 1; j <= 128; ++j) R[j] = (float)(h[2 * (128 - j)] / sum);
    std::vector<float2> pa(66), pb(66);
    for (int t = 0; t < 66; ++t) {
      pa[t] = make_float2(R[2 * t], R[2 * t + 1]);
      pb[t] = make_float2(R[2 * t + 1], R[2 * t + 2]);
    }
    const float centre = (float)(h[(kTaps - 1) / 2] / sum);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_pa, pa.data(), 66 * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_pb, pb.data(), 66 * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_centre, &centre, sizeof(float));
  }
  // interpolation onto np.linspace(f[0], f[-1], n_bins) and the DCT-II (ortho) rows
  {
    std::vector<int32_t> ilo(n_bins);
    std::vector<float> iw(n_bins);
    const double step = (freqs[n_bins - 1] - freqs[0]) / (n_bins - 1);
    for (int j = 0; j < n_bins; ++j) {
      const double x = j == n_bins - 1 ? freqs[n_bins - 1] : freqs[0] + j * step;
      int hi = 1;
      while (hi < n_bins - 1 && freqs[hi] < x) ++hi;            // np.searchsorted(side='left') clipped to [1, n - 1]
      ilo[j] = hi - 1;
      iw[j] = (float)((x - freqs[hi - 1]) / (freqs[hi] - freqs[hi - 1]));
    }
    std::vector<float> dct((size_t)n_ceps * n_bins);
    for (int c = 0; c < n_ceps; ++c) {
      const double fk = c == 0 ? std::sqrt(1.0 / (4.0 * n_bins)) : std::sqrt(1.0 / (2.0 * n_bins));
      for (int k = 0; k < n_bins; ++k) dct[(size_t)c * n_bins + k] = (float)(2.0 * fk * std::cos(kPi * c * (2.0 * k + 1.0) / (2.0 * n_bins)));
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->d_interp_lo, n_bins * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_interp_lo, ilo.data(), n_bins * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->d_interp_w, n_bins * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_interp_w, iw.data(), n_bins * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->d_dct, dct.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_dct, dct.data(), dct.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) {
    const size_t smem = 2 * (size_t)pl->n_fft * ((bpo + 3) / 4 * 4) * 4;
    if (smem > 200 * 1024 || (bpo + 3) / 4 > 6) e = cudaErrorInvalidValue;
    else if (smem > 48 * 1024) e = cudaFuncSetAttribute((const void*)k_cqt_octave, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const size_t smem_mma = (size_t)pl->n_fft / 8 * kMmaNT * 32 * sizeof(float4);
    if (e == cudaSuccess && smem_mma > 48 * 1024)
      e = cudaFuncSetAttribute((const void*)k_cqt_octave_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma);
    if (e == cudaSuccess && smem_mma > 48 * 1024)
      e = cudaFuncSetAttribute((const void*)k_cqt_octave_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma);
    cudaDeviceGetAttribute(&pl->sm_count, cudaDevAttrMultiProcessorCount, device);
    int occ = 0;
    if (e == cudaSuccess && smem_mma <= 200 * 1024 &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cqt_octave_mma<true>, kMmaWarps * 32, smem_mma) == cudaSuccess)
      pl->mma_ctas = std::max(1, occ);   // both instantiations are compiled for three CTAs per SM (launch bounds)
  }
  if (e != cudaSuccess) {
    aad_cqcc_plan_destroy(pl);
    return AAD_ERR_CUDA;
  }
  *out = pl;
  return AAD_OK;
}

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }
struct CqccWs {
  size_t off_max, off_sig[kMaxOct], off_mag, total;
  long long stride[kMaxOct];
};
static CqccWs cqcc_ws(const aad_cqcc_plan* pl, int B, int64_t max_len, int t_alloc) {
  CqccWs w;
  size_t o = 0;
  w.off_max = o; o = up256(o + (size_t)B * 4);
  long long len = max_len;
  w.stride[0] = 0; w.off_sig[0] = 0;
  for (int i = 1; i < pl->n_oct; ++i) {
    len = (len + 1) >> 1;
    w.stride[i] = (pl->n_fft / 2 + len + pl->n_fft + 3) / 4 * 4;  // zeros in front (n_fft / 2) and behind (n_fft)
    w.off_sig[i] = o; o = up256(o + (size_t)B * w.stride[i] * 4);
  }
  w.off_mag = o; o = up256(o + (size_t)B * pl->n_bins * t_alloc * 4);
  w.total = o;
  return w;
}

int aad_cqcc_query(const aad_cqcc_plan* pl, int B, int64_t max_len, int32_t* t_max, int32_t* n_ceps, int32_t* n_bins,
                   size_t* workspace_bytes) {
  if (!pl || B < 0 || max_len < 0 || max_len > 0x7fffffffLL) return AAD_ERR_INVALID_ARG;
  const int T = max_len > 0 ? (int)(1 + max_len / kHop) : 0;
  if (t_max) *t_max = T;
  if (n_ceps) *n_ceps = pl->n_ceps;
  if (n_bins) *n_bins = pl->n_bins;
  if (workspace_bytes) *workspace_bytes = cqcc_ws(pl, std::max(B, 1), std::max<int64_t>(max_len, 1), std::max(T, 1)).total;
  return AAD_OK;
}

int aad_cqcc(const aad_cqcc_plan* pl, const void* wav, int wav_dtype, int64_t wav_stride, const int64_t* row_off,
             const int32_t* lengths, int B, int64_t max_len, float* out, int64_t out_stride_b, int32_t t_alloc,
             int32_t* n_frames, int32_t* status, float* cqt_mag_out, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!pl || !wav || !lengths || !out || !n_frames || !status || !workspace) return AAD_ERR_INVALID_ARG;
  if (B <= 0 || max_len <= 0 || (!row_off && max_len > wav_stride) || t_alloc <= 0) return AAD_ERR_INVALID_ARG;
  if (wav_dtype != AAD_F32 && wav_dtype != AAD_I16) return AAD_ERR_INVALID_ARG;
  const int t_max = (int)(1 + max_len / kHop);
  const int t_ws = std::max(t_alloc, 1);
  const CqccWs w = cqcc_ws(pl, B, max_len, t_ws);
  if (workspace_bytes < w.total) return AAD_ERR_WORKSPACE;
  if (out_stride_b == 0) out_stride_b = (int64_t)pl->n_ceps * t_alloc;
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  int32_t* d_max = (int32_t*)(ws + w.off_max);
  float* d_mag = cqt_mag_out ? cqt_mag_out : (float*)(ws + w.off_mag);
  const long long mag_stride_b = (long long)pl->n_bins * t_ws;
  (void)cudaGetLastError();
  if (cudaMemsetAsync(d_max, 0, (size_t)B * 4, stream) != cudaSuccess) return AAD_ERR_CUDA;
  const int i16 = wav_dtype == AAD_I16;
  const int frames = std::min(t_max, t_ws);
  long long len = max_len;
  for (int i = 0; i < pl->n_oct; ++i) {
    const void* y = i == 0 ? wav : (const void*)(ws + w.off_sig[i]);
    const long long ystride = i == 0 ? wav_stride : w.stride[i];
    const int n_bg = (pl->n_k[i] + kOctBins - 1) / kOctBins;
    const size_t smem = 2 * (size_t)pl->n_fft * n_bg * kOctBins * 4;
    const dim3 grid_oct((frames + kOctFrames - 1) / kOctFrames, (B + kOctUtt - 1) / kOctUtt);
    const int pad_i = i == 0 ? 0 : pl->n_fft / 2;
    const long long* roff = i == 0 ? reinterpret_cast<const long long*>(row_off) : nullptr;
    const int hop_i = kHop >> i;
    const size_t smem_mma = (size_t)pl->n_fft / 8 * kMmaNT * 32 * sizeof(float4);
    if (pl->d_gfrag[i] && (i == 0 || hop_i % 4 == 0) && smem_mma <= 200 * 1024) {
      // tensor-core form: the library's padded float octave buffers with aligned 16-byte frame reads, the caller's
      // waveform (octave 0) through the generic loads
      const int mt_per_utt = (frames + 15) / 16;
      const long long items = (long long)B * mt_per_utt;
      const int grid_mma = (int)std::min<long long>((long long)pl->sm_count * pl->mma_ctas, (items + kMmaWarps - 1) / kMmaWarps);
      if (i == 0)
        k_cqt_octave_mma<true><<<grid_mma, kMmaWarps * 32, smem_mma, stream>>>(y, i16, ystride, roff, 0, lengths, B, hop_i, pl->n_fft,
                                                                               pl->d_gfrag[i], pl->n_k[i], pl->bin0[i], d_mag,
                                                                               mag_stride_b, t_ws, d_max, mt_per_utt);
      else
        k_cqt_octave_mma<false><<<grid_mma, kMmaWarps * 32, smem_mma, stream>>>(y, 0, ystride, nullptr, pad_i, lengths, B, hop_i, pl->n_fft,
                                                                                pl->d_gfrag[i], pl->n_k[i], pl->bin0[i], d_mag,
                                                                                mag_stride_b, t_ws, d_max, mt_per_utt);
    } else
    k_cqt_octave<<<grid_oct, kOctUtt * 16 * n_bg, smem, stream>>>(y, i == 0 ? i16 : 0, ystride, roff, pad_i, lengths, B, i, hop_i, pl->n_fft,
                                                           pl->d_g[i], pl->n_k[i], pl->bin0[i], d_mag, mag_stride_b, t_ws, d_max);
    if (i + 1 < pl->n_oct) {
      const long long len_out = (len + 1) >> 1;
      const dim3 grid_rs((unsigned)((len_out + pl->n_fft + kRsOut - 1) / kRsOut), B);
      k_cqt_resample<<<grid_rs, 128, 0, stream>>>(y, i == 0 ? i16 : 0, ystride, roff, pad_i, (float*)(ws + w.off_sig[i + 1]),
                                                  w.stride[i + 1], pl->n_fft / 2, pl->n_fft, lengths, i);
      len = len_out;
    }
  }
  const dim3 grid_epi((frames + kEpiFrames - 1) / kEpiFrames, B);
  const size_t smem_epi = (size_t)pl->n_bins * (std::max(kEpiFrames + 1, (pl->n_ceps + 3) & ~3) + kEpiFrames + 1) * 4;
  k_cqcc_epilogue<<<grid_epi, 256, smem_epi, stream>>>(d_mag, mag_stride_b, t_ws, pl->n_bins, lengths, d_max, pl->d_interp_lo,
                                                       pl->d_interp_w, pl->d_dct, pl->n_ceps, out, out_stride_b, n_frames, status);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess && e != cudaErrorNotReady) return AAD_ERR_CUDA;
  return AAD_OK;
}

}  // extern "C"
