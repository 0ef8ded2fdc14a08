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

// y_out[b][n] = sqrt(2) * sum_k h[k] * y_in[b][2 n + k - 127],  n < ceil(len_in / 2); zero outside [0, len_in).
// grid (chunks of 256 outputs, B); the 766 inputs of a chunk are staged in shared memory.
__global__ void __launch_bounds__(256) k_cqt_resample(const void* in, int in_i16, long long in_stride, float* out,
                                                      long long out_stride, const int32_t* lengths, int shift,
                                                      const float* taps) {
  __shared__ float sIn[2 * 256 + kTaps + 1];
  __shared__ float sH[kTaps + 1];
  const int b = blockIdx.y;
  long long len0 = lengths[b];
  if (len0 < 0) len0 = 0;
  long long len_in = len0;
  for (int s = 0; s < shift; ++s) len_in = (len_in + 1) >> 1;  // length at the input octave
  const long long len_out = (len_in + 1) >> 1;
  const long long n0 = (long long)blockIdx.x * 256;
  if (n0 >= len_out) return;
  for (int i = threadIdx.x; i < kTaps; i += 256) sH[i] = taps[i];
  const long long base = 2 * n0 - (kTaps - 1) / 2;
  for (int i = threadIdx.x; i < 2 * 256 + kTaps - 1; i += 256) {
    const long long src = base + i;
    sIn[i] = (src >= 0 && src < len_in) ? load_sample(in, in_i16, (long long)b * in_stride + src) : 0.f;
  }
  __syncthreads();
  const long long n = n0 + threadIdx.x;
  if (n >= len_out) return;
  const float* x = sIn + 2 * threadIdx.x;
  float acc = sH[(kTaps - 1) / 2] * x[(kTaps - 1) / 2];
#pragma unroll 8
  for (int k = 0; k < kTaps; k += 2) acc = __fmaf_rn(sH[k], x[k], acc);  // odd offsets from the centre (centre = 127)
  out[(long long)b * out_stride + n] = 1.41421356237309515f * acc;
}

// One octave: mag[b][bin0 + k][t] = | sum_n g[k][n] * y[b][t * hop - n_fft / 2 + n] |  for the n_k bins of the octave.
// grid (frame blocks of 8, B); thread = (frame in block, bin): 8 x n_k <= 96 threads.  The 8 frames are staged in
// shared memory with an odd row stride (frames are hop apart, a multiple of 32 words in the top octaves).
constexpr int kFramesPerCta = 8;
__global__ void __launch_bounds__(128) k_cqt_octave(const void* y, int y_i16, long long y_stride, const int32_t* lengths,
                                                    int shift, int hop, int n_fft, const float2* g, int n_k, int bin0,
                                                    float* mag, long long mag_stride_b, int t_alloc, int32_t* utt_max) {
  extern __shared__ float smem[];
  float* sY = smem;                                             // [8][n_fft + 1]
  float2* sG = reinterpret_cast<float2*>(smem + ((kFramesPerCta * (n_fft + 1) + 1) & ~1));  // [n_fft][n_k]
  const int b = blockIdx.y;
  long long len0 = lengths[b];
  if (len0 <= 0) return;
  const int T = (int)min((long long)t_alloc, 1 + len0 / kHop);  // frames of the utterance (all octaves are trimmed to it)
  const int t0 = blockIdx.x * kFramesPerCta;
  if (t0 >= T) return;
  long long len = len0;
  for (int s = 0; s < shift; ++s) len = (len + 1) >> 1;
  for (int i = threadIdx.x; i < n_fft * n_k; i += blockDim.x) {
    const int n = i / n_k, k = i - n * n_k;
    sG[i] = __ldg(g + (size_t)k * n_fft + n);
  }
  for (int i = threadIdx.x; i < kFramesPerCta * n_fft; i += blockDim.x) {
    const int f = i / n_fft, n = i - f * n_fft;
    const long long src = (long long)(t0 + f) * hop - n_fft / 2 + n;
    sY[f * (n_fft + 1) + n] = (t0 + f < T && src >= 0 && src < len) ? load_sample(y, y_i16, (long long)b * y_stride + src) : 0.f;
  }
  __syncthreads();
  const int f = threadIdx.x / n_k, k = threadIdx.x - f * n_k;
  if (f >= kFramesPerCta || t0 + f >= T) return;
  const float* yr = sY + f * (n_fft + 1);
  float2 a0 = make_float2(0.f, 0.f), a1 = a0;
  for (int n = 0; n < n_fft; n += 2) {
    a0 = __ffma2_rn(make_float2(yr[n], yr[n]), sG[n * n_k + k], a0);
    a1 = __ffma2_rn(make_float2(yr[n + 1], yr[n + 1]), sG[(n + 1) * n_k + k], a1);
  }
  const float re = a0.x + a1.x, im = a0.y + a1.y;
  const float m = sqrtf(__fmaf_rn(re, re, im * im));
  mag[(long long)b * mag_stride_b + (long long)(bin0 + k) * t_alloc + t0 + f] = m;
  if (m == m) atomicMax(utt_max + b, __float_as_int(m));        // non-negative floats order like their bit patterns
  else atomicMax(utt_max + b, 0x7fc00000);                      // NaN poisons the utterance
}

// dB relative to the utterance maximum (floor -80) -> interpolation onto the uniform frequency grid -> log(x^2 + 1e-12)
// -> DCT-II ortho.  grid (frame blocks of 32, B), 256 threads.
constexpr int kEpiFrames = 32;
__global__ void __launch_bounds__(256) k_cqcc_epilogue(const float* mag, long long mag_stride_b, int t_alloc, int n_bins,
                                                       const int32_t* lengths, const int32_t* utt_max, const int32_t* interp_lo,
                                                       const float* interp_w, const float* dct, int n_ceps, float* out,
                                                       long long out_stride_b, int32_t* n_frames, int32_t* status) {
  extern __shared__ float smem[];
  float* sDb = smem;                                // [n_bins][33]
  float* sLp = smem + n_bins * (kEpiFrames + 1);    // [n_bins][33]
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
  const float ref_db = 10.0f * log10f(fmaxf(amin2, vmax * vmax));
  const float* mb = mag + (long long)b * mag_stride_b;
  for (int i = threadIdx.x; i < n_bins * kEpiFrames; i += 256) {
    const int k = i / kEpiFrames, f = i - k * kEpiFrames;
    float v = 0.f;
    if (f < nt) {
      const float m = mb[(long long)k * t_alloc + t0 + f];
      v = fmaxf(10.0f * log10f(fmaxf(amin2, m * m)) - ref_db, -80.0f);   // the maximum of log_spec is 0 (the reference)
    }
    sDb[k * (kEpiFrames + 1) + f] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_bins * kEpiFrames; i += 256) {
    const int k = i / kEpiFrames, f = i - k * kEpiFrames;
    const int lo = interp_lo[k];
    const float x0 = sDb[lo * (kEpiFrames + 1) + f], x1 = sDb[(lo + 1) * (kEpiFrames + 1) + f];
    const float x = __fmaf_rn(interp_w[k], x1 - x0, x0);
    sLp[k * (kEpiFrames + 1) + f] = logf(__fmaf_rn(x, x, 1e-12f));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_ceps * kEpiFrames; i += 256) {
    const int c = i / kEpiFrames, f = i - c * kEpiFrames;
    if (f >= nt) continue;
    const float* d = dct + (size_t)c * n_bins;
    float acc = 0.f;
    for (int k = 0; k < n_bins; ++k) acc = __fmaf_rn(__ldg(d + k), sLp[k * (kEpiFrames + 1) + f], acc);
    out[(long long)b * out_stride_b + (long long)c * t_alloc + t0 + f] = acc;
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
  int n_fft = 0;                    // per-octave transform size (the same for every octave: wavelet lengths repeat)
  int n_k[kMaxOct] = {0};           // bins of octave i (top first)
  int bin0[kMaxOct] = {0};
  float2* d_g[kMaxOct] = {nullptr}; // [n_k][n_fft] taps of octave i
  float* d_taps = nullptr;          // resampler
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
  cudaFree(pl->d_taps);
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
    std::vector<float> hf(kTaps);
    for (int i = 0; i < kTaps; ++i) hf[i] = (float)(h[i] / sum);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->d_taps, kTaps * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_taps, hf.data(), kTaps * sizeof(float), cudaMemcpyHostToDevice);
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
    const size_t smem = ((size_t)((kFramesPerCta * (pl->n_fft + 1) + 1) & ~1) + 2 * (size_t)pl->n_fft * bpo) * 4;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute((const void*)k_cqt_octave, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
    w.stride[i] = (len + 3) / 4 * 4;
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

int aad_cqcc(const aad_cqcc_plan* pl, const void* wav, int wav_dtype, int64_t wav_stride, const int32_t* lengths, int B,
             int64_t max_len, float* out, int64_t out_stride_b, int32_t t_alloc, int32_t* n_frames, int32_t* status,
             float* cqt_mag_out, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!pl || !wav || !lengths || !out || !n_frames || !status || !workspace) return AAD_ERR_INVALID_ARG;
  if (B <= 0 || max_len <= 0 || max_len > wav_stride || t_alloc <= 0) return AAD_ERR_INVALID_ARG;
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
  const dim3 grid_oct((frames + kFramesPerCta - 1) / kFramesPerCta, B);
  long long len = max_len;
  for (int i = 0; i < pl->n_oct; ++i) {
    const void* y = i == 0 ? wav : (const void*)(ws + w.off_sig[i]);
    const long long ystride = i == 0 ? wav_stride : w.stride[i];
    const size_t smem = ((size_t)((kFramesPerCta * (pl->n_fft + 1) + 1) & ~1) + 2 * (size_t)pl->n_fft * pl->n_k[i]) * 4;
    k_cqt_octave<<<grid_oct, 128, smem, stream>>>(y, i == 0 ? i16 : 0, ystride, lengths, i, kHop >> i, pl->n_fft, pl->d_g[i],
                                                  pl->n_k[i], pl->bin0[i], d_mag, mag_stride_b, t_ws, d_max);
    if (i + 1 < pl->n_oct) {
      const long long len_out = (len + 1) >> 1;
      const dim3 grid_rs((unsigned)((len_out + 255) / 256), B);
      k_cqt_resample<<<grid_rs, 256, 0, stream>>>(y, i == 0 ? i16 : 0, ystride, (float*)(ws + w.off_sig[i + 1]), w.stride[i + 1],
                                                  lengths, i, pl->d_taps);
      len = len_out;
    }
  }
  const dim3 grid_epi((frames + kEpiFrames - 1) / kEpiFrames, B);
  const size_t smem_epi = 2 * (size_t)pl->n_bins * (kEpiFrames + 1) * 4;
  k_cqcc_epilogue<<<grid_epi, 256, smem_epi, stream>>>(d_mag, mag_stride_b, t_ws, pl->n_bins, lengths, d_max, pl->d_interp_lo,
                                                       pl->d_interp_w, pl->d_dct, pl->n_ceps, out, out_stride_b, n_frames, status);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess && e != cudaErrorNotReady) return AAD_ERR_CUDA;
  return AAD_OK;
}

}  // extern "C"
