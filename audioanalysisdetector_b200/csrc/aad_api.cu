// C ABI + plan (host tables) + launch orchestration for libaad_b200.so.
// See include/aad.h for the contract of every entry point.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/aad.h"
#include "aad_stft_inst.h"

using namespace aad;

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
struct aad_plan {
  aad_params p;
  int device = 0;
  int sm_count = 148;
  int L = 0;          // n_fft / 64
  int K = 0;          // bins
  int tile = 32;      // frames per K1 tile (16 or 32)
  int warps = 0, ctas = 0;
  size_t k1_smem = 0;
  int n_w4 = 0, n_hdr = 0;
  bool dense = false;  // dense filter bank (gammatone / custom dense): k_stft_fb FBM = 1 (8 warps, two CTAs per SM)
  int smem_optin = 0;  // device limit of dynamic shared memory per CTA
  int n_ksteps = 0, n_tiles = 0, cep_nt = 1;  // K2: DCT as a GEMM (K steps of 8 filters, N tiles of 8 coefficients)
  int c_feat = 0;     // rows before deltas
  int c_out = 0;
  bool need_ws_E = false;     // filterbank energies go to workspace (cepstra / layout / mean follows)
  bool need_ws_feat = false;  // features go to workspace (time_mean follows)
  // host copies (introspection)
  std::vector<float> h_window, h_fb, h_dct;
  float taps[2][CEP_MAXW];
  // device tables
  float* d_window = nullptr;
  float* d_window_i16 = nullptr;  // window * i16_scale (int16 input)
  float2* d_tw1 = nullptr;
  float2* d_twp = nullptr;
  int2* d_filt_hdr = nullptr;
  float4* d_filt_w = nullptr;
  int4* d_warp_prog = nullptr;
  float4* d_dct_frag = nullptr;
  float* d_dct_colsum = nullptr;
  // optional per-kernel timing (bench roofline): events recorded around each launch
  bool profile = false;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // host-path resources (lazy)
  struct HostBuf {
    cudaStream_t stream = nullptr;
    void* d_wav = nullptr;
    size_t wav_bytes = 0;
    float* d_out = nullptr;
    size_t out_bytes = 0;
    int32_t *d_len = nullptr, *d_nf = nullptr, *d_st = nullptr;
    int cap_b = 0;
    void* d_ws = nullptr;
    size_t ws_bytes = 0;
  } hb[3];
  // pinned staging for the small per-utterance arrays of the host path: async copies to or from
  // pageable memory block the calling thread until they complete, which would serialise the chunks
  int32_t* h_stage = nullptr;  // [3][cap] lengths, n_frames, status
  int h_stage_cap = 0;
};

// Entry points that must run on the plan's device select it and restore the caller's current device on every exit
// path (single-process multi-GPU programs: torch's current device must not flip under the caller).
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

static const double kPiD = 3.141592653589793238462643383279502884;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static thread_local char g_err_detail[256] = "";
static int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err_detail, sizeof(g_err_detail), "%s: %s (%d)", what, cudaGetErrorString(e), (int)e);
  return AAD_ERR_CUDA;
}
#define CUDA_TRY(expr)                                  \
  do {                                                  \
    cudaError_t e__ = (expr);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #expr); \
  } while (0)
// launch check: cudaErrorNotReady can be left behind by event/stream queries and is not a failure
#define LAUNCH_CHECK(what)                                             \
  do {                                                                 \
    cudaError_t e__ = cudaGetLastError();                              \
    if (e__ != cudaSuccess && e__ != cudaErrorNotReady) return cuda_fail(e__, what); \
  } while (0)

// ---- Slaney mel scale (librosa.hz_to_mel / mel_to_hz, htk=False) ------------
static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  if (f >= min_log_hz) return min_log_mel + std::log(f / min_log_hz) / logstep;
  return f / f_sp;
}
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  if (m >= min_log_mel) return min_log_hz * std::exp(logstep * (m - min_log_mel));
  return f_sp * m;
}
static std::vector<double> linspace(double a, double b, int n) {
  std::vector<double> y(n);
  double step = n > 1 ? (b - a) / (n - 1) : 0.0;
  for (int i = 0; i < n; ++i) y[i] = i * step + a;
  if (n > 1) y[n - 1] = b;
  return y;
}

// dense float32 filterbank [n_filt][K], reference rounding order reproduced
static int build_filterbank(const aad_params& p, int K, std::vector<float>& fb) {
  const int nf = p.n_filt;
  const double sr = p.sample_rate;
  const double fmax = p.fmax > 0 ? p.fmax : sr / 2;
  const double fmin = p.fmin;
  const double pscale = p.power_scale != 0.f ? (double)p.power_scale : 1.0;
  fb.assign((size_t)nf * K, 0.f);
  if (p.fb_type == AAD_FB_MEL_SLANEY) {
    // librosa.filters.mel: float64 triangles -> float32 store -> in-place *= enorm (float64)
    std::vector<double> mels = linspace(hz_to_mel(fmin), hz_to_mel(fmax), nf + 2), mel_f(nf + 2);
    for (int i = 0; i < nf + 2; ++i) mel_f[i] = mel_to_hz(mels[i]);
    const double val = 1.0 / (p.n_fft * (1.0 / sr));
    for (int i = 0; i < nf; ++i) {
      const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
      const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
      for (int k = 0; k < K; ++k) {
        const double f = k * val;
        const double lower = -(mel_f[i] - f) / fd0, upper = (mel_f[i + 2] - f) / fd1;
        float w = (float)std::max(0.0, std::min(lower, upper));
        w = (float)((double)w * enorm);
        if (pscale != 1.0) w = (float)((double)w * pscale);
        fb[(size_t)i * K + k] = w;
      }
    }
  } else if (p.fb_type == AAD_FB_LINEAR_INTBIN) {
    // spafe 0.1.x / python_speech_features style: triangles on integer FFT bins
    std::vector<double> pts = linspace(fmin, fmax, nf + 2), bins(nf + 2);
    for (int i = 0; i < nf + 2; ++i) bins[i] = std::floor((p.n_fft + 1) * pts[i] / sr);
    for (int j = 0; j < nf; ++j) {
      const double b0 = bins[j], b1 = bins[j + 1], b2 = bins[j + 2];
      for (int i = (int)b0; i < (int)b1 && i < K; ++i)
        fb[(size_t)j * K + i] = (float)(std::fabs((i - (int)b0) / (b1 - b0)) * pscale);
      for (int i = (int)b1; i < (int)b2 && i < K; ++i)
        fb[(size_t)j * K + i] = (float)(std::fabs(((int)b2 - i) / (b2 - b1)) * pscale);
    }
  } else if (p.fb_type == AAD_FB_LINEAR_CONT) {
    // spafe 0.3.x linear_filter_banks (scale="constant"): edges low + k |high - low| / (nfilts + 1), bin
    // frequencies np.linspace(low, high, K), inclusive masks on both slopes (the falling slope wins at a centre)
    const double delta = std::fabs(fmax - fmin) / (nf + 1);
    std::vector<double> edges(nf + 2), freqs = linspace(fmin, fmax, K);
    for (int i = 0; i < nf + 2; ++i) edges[i] = fmin + delta * i;
    for (int j = 0; j < nf; ++j) {
      const double lo = edges[j], ce = edges[j + 1], hi = edges[j + 2];
      for (int k = 0; k < K; ++k) {
        const double f = freqs[k];
        double w = 0.0;
        if (f >= lo && f <= ce) w = (f - lo) / (ce - lo);
        if (f >= ce && f <= hi) w = (hi - f) / (hi - ce);
        fb[(size_t)j * K + k] = (float)(w * pscale);
      }
    }
  } else if (p.fb_type == AAD_FB_GAMMATONE) {
    // spafe 0.3.x gammatone_filter_banks(nfilts, nfft, fs, low_freq, high_freq, scale="constant", order=4): Slaney's
    // ERB filter cascade evaluated on the unit circle at the FFT bins.  Centre frequencies on the ERB scale
    // (generate_center_frequencies, ascending after the final reversal), bandwidths B = 1.019 * 2 pi * ERB with
    // ERB = ((fc / EarQ)^order + minBW^order)^(1 / order), four zeros A_i and a double pole pair; the gain factor and
    // T^4 cancel against the final "every filter has maximum 1" normalisation, the scale is constant 1.
    const double EarQ = 9.26449, minBW = 24.7, order = 4.0, PI = 3.141592653589793238462643383279502884;
    const double T = 1.0 / sr, c = EarQ * minBW;
    const double smax = std::sqrt(3.0 + std::pow(2.0, 1.5)), smin = std::sqrt(3.0 - std::pow(2.0, 1.5));
    for (int i = 0; i < nf; ++i) {
      // center_freqs = (max + c) exp((m / nfilts) log((min + c) / (max + c))) - c, m = 1 .. nfilts, then reversed
      const int m = nf - i;
      const double fc = (fmax + c) * std::exp(((double)m / nf) * std::log((fmin + c) / (fmax + c))) - c;
      const double erb = std::pow(std::pow(fc / EarQ, order) + std::pow(minBW, order), 1.0 / order);
      const double Bw = 1.019 * 2.0 * PI * erb;
      const double wT = 2.0 * fc * PI * T, Kx = std::exp(Bw * T);
      const double co = std::cos(wT), si = std::sin(wT);
      const double A[4] = {(co + smax * si) / Kx, (co - smax * si) / Kx, (co + smin * si) / Kx, (co - smin * si) / Kx};
      const std::complex<double> pole = std::polar(1.0 / Kx, wT);
      std::vector<double> row(K);
      double mx = 0.0;
      for (int k = 0; k < K; ++k) {
        const std::complex<double> u = std::polar(1.0, 2.0 * PI * k / p.n_fft);
        double num = 1.0;
        for (int q = 0; q < 4; ++q) num *= std::abs(u - A[q]);
        const double den = std::abs((u - pole) * (u - std::conj(pole)));
        row[k] = num * std::pow(den, -4.0);
        mx = std::max(mx, row[k]);
      }
      for (int k = 0; k < K; ++k) fb[(size_t)i * K + k] = (float)(row[k] / mx * pscale);
    }
  } else if (p.fb_type == AAD_FB_CUSTOM || p.fb_type == AAD_FB_CUSTOM_DENSE) {
    if (!p.custom_fb) return AAD_ERR_INVALID_ARG;
    for (size_t i = 0; i < (size_t)nf * K; ++i) fb[i] = (float)((double)p.custom_fb[i] * pscale);
  } else {
    return AAD_ERR_INVALID_ARG;
  }
  return AAD_OK;
}

// dense -> banded two-tap form: per bin (rising weight of filter seg, falling weight of filter seg-1)
static int band_filterbank(const std::vector<float>& fb, int nf, int K, std::vector<float2>& fbw,
                           std::vector<int32_t>& seg_bounds) {
  std::vector<int> seg(K);
  int cur = 0;
  for (int k = 0; k < K; ++k) {
    int jmin = -1, jmax = -1;
    for (int j = 0; j < nf; ++j)
      if (fb[(size_t)j * K + k] != 0.f) {
        if (jmin < 0) jmin = j;
        jmax = j;
      }
    int s = cur;
    if (jmin >= 0) {
      if (jmax - jmin > 1) return AAD_ERR_FILTERBANK;
      for (int j = jmin + 1; j < jmax; ++j)
        if (fb[(size_t)j * K + k] != 0.f) return AAD_ERR_FILTERBANK;
      if (jmax == jmin) {
        if (cur <= jmin) s = jmin;
        else if (cur == jmin + 1) s = cur;
        else return AAD_ERR_FILTERBANK;
      } else {
        s = jmin + 1;
        if (cur > s) return AAD_ERR_FILTERBANK;
      }
    }
    seg[k] = cur = s;
  }
  fbw.assign(K, make_float2(0.f, 0.f));
  for (int k = 0; k < K; ++k) {
    int s = seg[k];
    float wr = s < nf ? fb[(size_t)s * K + k] : 0.f;
    float wf = s >= 1 ? fb[(size_t)(s - 1) * K + k] : 0.f;
    fbw[k] = make_float2(wr, wf);
  }
  seg_bounds.assign(nf + 2, K);
  int k = 0;
  for (int s = 0; s <= nf + 1; ++s) {
    while (k < K && seg[k] < s) ++k;
    seg_bounds[s] = (s == nf + 1) ? K : k;
  }
  return AAD_OK;
}


static void savgol_taps(int width, float t1[CEP_MAXW], float t2[CEP_MAXW]) {
  // least-squares polynomial fit of degree d, d-th derivative at the window centre
  // (scipy.signal.savgol_coeffs(width, polyorder=d, deriv=d), d = 1, 2)
  const int h = width / 2;
  double s2 = 0, s4 = 0;
  for (int x = -h; x <= h; ++x) {
    s2 += (double)x * x;
    s4 += (double)x * x * x * x;
  }
  const double n = width;
  for (int i = 0; i < CEP_MAXW; ++i) t1[i] = t2[i] = 0.f;
  for (int i = 0; i < width; ++i) {
    const double x = i - h;
    t1[i] = (float)(x / s2);
    t2[i] = (float)(2.0 * (n * x * x - s2) / (n * s4 - s2 * s2));
  }
}

template <typename T>
static cudaError_t upload(T** dptr, const std::vector<T>& h) {
  cudaError_t e = cudaMalloc((void**)dptr, std::max<size_t>(h.size(), 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  if (h.empty()) return cudaSuccess;
  return cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

// Launch with programmatic dependent launch (PDL): the kernel may be scheduled while its predecessor in the stream
// is still running and executes its own prologue (tensor-memory allocation, table fills) meanwhile; it reaches the
// predecessor's results only behind `griddepcontrol.wait`.  Hides the launch gaps and K0 behind K1's set-up, which is
// what a step on a small batch (BASELINE configs[0]: 64 clips) consists of.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// dB reference / floor pass: rows shorter than 256 frames take the flat kernel (a CTA per 8192 elements of the utterance's
// block), longer rows the row kernel (a warp per row and 1024-frame chunk)
static int launch_db_finalize(const FinArgs& fa, int B, int t_max, cudaStream_t stream, bool pdl) {
  if (fa.stride_f < 256 && (long long)fa.n_filt * fa.stride_f < (1ll << 30)) {
    const int n_chunks = (int)(((long long)fa.n_filt * fa.stride_f + FIN_FLAT - 1) / FIN_FLAT);
    const long long nblk = (long long)B * n_chunks;
    if (nblk > 0x7fffffffLL) return AAD_ERR_UNSUPPORTED;
    launch_pdl(k_db_finalize_flat, dim3((unsigned)nblk), dim3(256), 0, stream, pdl, fa, n_chunks);
    return AAD_OK;
  }
  const int n_row_blocks = (fa.n_filt + FIN_ROWS - 1) / FIN_ROWS;
  const int n_chunks = (std::max(t_max, 1) + FIN_CHUNK - 1) / FIN_CHUNK;
  const long long nblk = (long long)B * n_row_blocks * n_chunks;
  if (nblk > 0x7fffffffLL) return AAD_ERR_UNSUPPORTED;
  launch_pdl(k_db_finalize, dim3((unsigned)nblk), dim3(256), 0, stream, pdl, fa, n_row_blocks, n_chunks);
  return AAD_OK;
}

// ---- kernel dispatch table --------------------------------------------------
static stft_kernel_t pick_stft(int L, int tile, int mode, bool pre, bool pair = false, bool dense = false) {
  if (tile != stft_tile(L, dense)) return nullptr;
  if (dense) return L == 8 && !pair ? pick_stft_L8_dense(mode, pre) : nullptr;
  switch (L) {  // instantiated per transform size in aad_stft_inst.cu
    case 4: return pick_stft_L4(mode, pre, pair);
    case 8: return pick_stft_L8(mode, pre, pair);
    case 16: return pick_stft_L16(mode, pre, pair);
    case 32: return pick_stft_L32(mode, pre, pair);
  }
  return nullptr;
}
template <int L, int TILE, bool DENSE = false>
static void stft_cfg_LT(int* warps, int* ctas, size_t* fixed, int* fbu) {
  using C = StftCfg<L, TILE, DENSE>;
  *warps = C::WARPS; *ctas = C::CTAS; *fixed = C::FIXED_BYTES; *fbu = DENSE ? kDenseFbu : C::FBU;
}
static void stft_cfg(int L, bool dense, int* warps, int* ctas, size_t* fixed, int* fbu) {
  if (dense) {  // n_fft 512 only (checked by the caller)
    stft_cfg_LT<8, stft_tile(8, true), true>(warps, ctas, fixed, fbu);
    return;
  }
  switch (L) {
    case 4: stft_cfg_LT<4, 32>(warps, ctas, fixed, fbu); break;
    case 8: stft_cfg_LT<8, stft_tile(8)>(warps, ctas, fixed, fbu); break;
    case 16: stft_cfg_LT<16, 32>(warps, ctas, fixed, fbu); break;
    default: stft_cfg_LT<32, 32>(warps, ctas, fixed, fbu);
  }
}

typedef void (*cep_kernel_t)(const CepArgs);
template <int TS>
static cep_kernel_t pick_cep_ts(int nt) {  // coefficient n-tiles per pass
  switch (nt) {
    case 1: return k_cepstra<1, TS>;
    case 2: return k_cepstra<2, TS>;
    case 3: return k_cepstra<3, TS>;
    case 4: return k_cepstra<4, TS>;
    case 5: return k_cepstra<5, TS>;
    case 6: return k_cepstra<6, TS>;
    case 7: return k_cepstra<7, TS>;
    default: return k_cepstra<8, TS>;
  }
}
// frames per k_cepstra tile: 64 when no utterance of the call is longer (the reference's 2-second chunks: 63 frames)
static int cep_tile(int t_max) { return t_max <= 64 ? 64 : CEP_TS; }
static cep_kernel_t pick_cep(int nt, int ts) { return ts == 64 ? pick_cep_ts<64>(nt) : pick_cep_ts<CEP_TS>(nt); }

static size_t cep_smem_bytes(const aad_plan* pl, int ts) {
  const int kf = pl->p.n_ceps > 0 ? 8 * pl->n_ksteps : pl->p.n_filt;
  size_t f = 4 + (size_t)kf * (ts + 8) + 8 * ts;
  if (pl->p.n_ceps > 0) {
    f += (size_t)pl->n_ksteps * pl->n_tiles * 32 * 4 + (size_t)pl->n_tiles * 8;
    if (pl->n_tiles > CEP_MAXNT) f += (size_t)pl->p.n_ceps * (ts + 4);  // no aliasing with several passes
  }
  return f * 4;
}

// ---------------------------------------------------------------------------
extern "C" {

int aad_version(void) { return AAD_VERSION; }

const char* aad_last_error_detail(void) { return g_err_detail; }

const char* aad_strerror(int err) {
  switch (err) {
    case AAD_OK: return "ok";
    case AAD_ERR_INVALID_ARG: return "invalid argument";
    case AAD_ERR_UNSUPPORTED: return "unsupported configuration";
    case AAD_ERR_CUDA: return "CUDA runtime error";
    case AAD_ERR_WORKSPACE: return "workspace too small";
    case AAD_ERR_KIND: return "plan kind does not match entry point";
    case AAD_ERR_PAIR: return "plans cannot be paired (different STFT, or the second is not a plain log filter bank)";
    case AAD_ERR_FILTERBANK: return "filterbank is not banded (at most two adjacent filters per bin)";
    case AAD_ERR_FORMAT: return "malformed or unsupported audio stream";
  }
  return "unknown error";
}

int aad_params_default(aad_params* p, int kind, int sample_rate) {
  if (!p || sample_rate <= 0) return AAD_ERR_INVALID_ARG;
  std::memset(p, 0, sizeof(*p));
  p->struct_size = (int32_t)sizeof(aad_params);
  p->kind = kind;
  p->sample_rate = sample_rate;
  p->amin = 1e-10f;
  p->top_db = 80.f;
  p->delta_width = 9;
  p->fmin = 0.f;
  p->fmax = 0.f;
  p->power_scale = 1.f;
  if (kind == AAD_KIND_LOGMEL || kind == AAD_KIND_MFCC) {
    p->n_fft = 2048;
    p->win_length = 2048;
    p->hop_length = 512;
    p->window = AAD_WIN_HANN_PERIODIC;
    p->center = 1;
    p->fb_type = AAD_FB_MEL_SLANEY;
    p->log_type = AAD_LOG_DB10;
    p->layout = AAD_LAYOUT_CT;
    if (kind == AAD_KIND_LOGMEL) {
      p->n_filt = 64;
      p->ref_type = AAD_REF_UTT_MAX;
      p->n_ceps = 0;
    } else {
      p->n_filt = 128;
      p->ref_type = AAD_REF_ONE;
      p->n_ceps = 13;
    }
  } else if (kind == AAD_KIND_LFCC) {
    p->n_fft = 512;
    p->win_length = (int)(0.025 * sample_rate);
    p->hop_length = (int)(0.01 * sample_rate);
    p->window = AAD_WIN_HAMMING_SYMMETRIC;
    p->center = 0;
    p->quantize_i16 = 1;
    p->pre_emph = 0.97f;
    p->n_filt = 24;
    p->fb_type = AAD_FB_LINEAR_CONT;  /* spafe ~= 0.3.3 (requirements.txt:5); see oracle/spafe_ref.py */
    p->power_scale = 1.0f / 512.0f;
    p->log_type = AAD_LOG_LN;
    p->ref_type = AAD_REF_ONE;
    p->top_db = -1.f;
    p->n_ceps = 13;
    p->layout = AAD_LAYOUT_TC;
  } else if (kind == AAD_KIND_GTCC) {
    // spafe 0.3.x gfcc as the reference calls it: gfcc(sig=y, fs=sr, num_ceps=n_ceps, nfilts=n_filters) with the
    // float waveform of librosa.load (no int16 step), n_filters = 40, n_ceps = 13
    p->n_fft = 512;
    p->win_length = (int)(0.025 * sample_rate);
    p->hop_length = (int)(0.01 * sample_rate);
    p->window = AAD_WIN_HAMMING_SYMMETRIC;
    p->center = 0;
    p->quantize_i16 = 0;
    p->pre_emph = 0.97f;
    p->n_filt = 40;
    p->fb_type = AAD_FB_GAMMATONE;
    p->power_scale = 1.0f / 512.0f;
    p->spectrum = AAD_SPEC_POWER;
    p->log_type = AAD_LOG_CBRT;
    p->ref_type = AAD_REF_ONE;
    p->top_db = -1.f;
    p->n_ceps = 13;
    p->layout = AAD_LAYOUT_TC;
  } else {
    return AAD_ERR_INVALID_ARG;
  }
  return AAD_OK;
}

int aad_plan_destroy(aad_plan* pl) {
  if (!pl) return AAD_OK;
  DeviceGuard guard(pl->device);
  cudaFree(pl->d_window);
  cudaFree(pl->d_window_i16);
  cudaFree(pl->d_tw1);
  cudaFree(pl->d_twp);
  cudaFree(pl->d_filt_hdr);
  cudaFree(pl->d_filt_w);
  cudaFree(pl->d_warp_prog);
  cudaFree(pl->d_dct_frag);
  cudaFree(pl->d_dct_colsum);
  for (auto& e : pl->ev)
    if (e) cudaEventDestroy(e);
  if (pl->h_stage) cudaFreeHost(pl->h_stage);
  for (auto& h : pl->hb) {
    cudaFree(h.d_wav);
    cudaFree(h.d_out);
    cudaFree(h.d_len);
    cudaFree(h.d_nf);
    cudaFree(h.d_st);
    cudaFree(h.d_ws);
    if (h.stream) cudaStreamDestroy(h.stream);
  }
  delete pl;
  return AAD_OK;
}

int aad_plan_create(const aad_params* pp, int device, aad_plan** out) {
  if (!pp || !out) return AAD_ERR_INVALID_ARG;
  if (pp->struct_size != (int32_t)sizeof(aad_params)) return AAD_ERR_INVALID_ARG;
  const aad_params& p = *pp;
  if (p.n_fft != 256 && p.n_fft != 512 && p.n_fft != 1024 && p.n_fft != 2048) return AAD_ERR_UNSUPPORTED;
  // win_length > n_fft is spafe's behaviour above 20.48 kHz (25 ms frames, np.fft.fft(frames, 512) keeps the first
  // n_fft windowed samples): allowed without centring only
  if (p.win_length <= 0 || p.hop_length <= 0 || (p.win_length > p.n_fft && p.center)) return AAD_ERR_INVALID_ARG;
  if (p.n_filt <= 0 || p.n_filt > 512 || p.sample_rate <= 0) return AAD_ERR_INVALID_ARG;
  if (p.n_ceps < 0 || p.n_ceps > p.n_filt) return AAD_ERR_INVALID_ARG;
  if (p.n_delta < 0 || p.n_delta > 2) return AAD_ERR_INVALID_ARG;
  if (p.n_delta > 0 && (p.delta_width < 3 || p.delta_width > CEP_MAXW || p.delta_width % 2 == 0))
    return AAD_ERR_INVALID_ARG;
  if (p.window != AAD_WIN_HANN_PERIODIC && p.window != AAD_WIN_HAMMING_SYMMETRIC) return AAD_ERR_INVALID_ARG;
  const bool dense_fb = p.fb_type == AAD_FB_GAMMATONE || p.fb_type == AAD_FB_CUSTOM_DENSE;
  if (p.log_type != AAD_LOG_DB10 && p.log_type != AAD_LOG_LN && !(p.log_type == AAD_LOG_CBRT && dense_fb)) return AAD_ERR_INVALID_ARG;
  if (p.spectrum != AAD_SPEC_POWER && !(p.spectrum == AAD_SPEC_MAGNITUDE && dense_fb)) return AAD_ERR_INVALID_ARG;
  if (dense_fb && (p.n_fft != 512 || p.n_filt > 64)) return AAD_ERR_UNSUPPORTED;  // one kernel family carries the dense form
  if (p.layout != AAD_LAYOUT_CT && p.layout != AAD_LAYOUT_TC) return AAD_ERR_INVALID_ARG;
  if (p.fmax > 0 && p.fmax > p.sample_rate / 2.0f + 1e-3f) return AAD_ERR_INVALID_ARG;
  if (p.znorm && p.time_mean) return AAD_ERR_INVALID_ARG;

  DeviceGuard guard(device);
  CUDA_TRY(guard.err);
  aad_plan* pl = new (std::nothrow) aad_plan();
  if (!pl) return AAD_ERR_INVALID_ARG;
  pl->p = p;
  pl->p.custom_fb = nullptr;
  pl->device = device;
  cudaDeviceGetAttribute(&pl->sm_count, cudaDevAttrMultiProcessorCount, device);
  pl->L = p.n_fft / 64;
  pl->K = p.n_fft / 2 + 1;
  pl->tile = stft_tile(pl->L, dense_fb);  // frames per K1 tile (n_fft 2048: one 16-warp CTA per SM)
  size_t k1_fixed = 0;
  int fbu = 1;
  stft_cfg(pl->L, dense_fb, &pl->warps, &pl->ctas, &k1_fixed, &fbu);
  pl->c_feat = p.n_ceps > 0 ? p.n_ceps : p.n_filt;
  pl->c_out = pl->c_feat * (1 + p.n_delta);
  pl->n_ksteps = (p.n_filt + 7) / 8;
  pl->n_tiles = (p.n_ceps + 7) / 8;
  pl->cep_nt = std::max(1, std::min(pl->n_tiles, CEP_MAXNT));
  const bool direct = (p.n_ceps == 0 && p.n_delta == 0 && p.layout == AAD_LAYOUT_CT && !p.time_mean);
  pl->need_ws_E = !direct;
  pl->need_ws_feat = p.time_mean != 0;

  const int N = p.n_fft, M = N / 2, L = pl->L, K = pl->K;
  // window (float64 -> float32), placed in the n_fft buffer
  std::vector<double> w(p.win_length);
  for (int n = 0; n < p.win_length; ++n) {
    if (p.window == AAD_WIN_HANN_PERIODIC) w[n] = 0.5 - 0.5 * std::cos(2.0 * kPiD * n / p.win_length);
    else w[n] = p.win_length > 1 ? 0.54 - 0.46 * std::cos(2.0 * kPiD * n / (p.win_length - 1)) : 1.0;
  }
  const int win_off = p.center ? (N - p.win_length) / 2 : 0;
  pl->h_window.assign(N, 0.f);
  std::vector<float> win_half(N, 0.f);
  for (int n = 0; n < std::min(p.win_length, N); ++n) {  // a window longer than n_fft is cut at n_fft (see above)
    pl->h_window[win_off + n] = (float)w[n];
    win_half[win_off + n] = 0.5f * (float)w[n];
  }
  std::vector<float2> tw1((size_t)32 * L), twp(M / 2);
  for (int ka = 0; ka < 32; ++ka)
    for (int b = 0; b < L; ++b) {
      double ang = -2.0 * kPiD * (double)((long long)b * ka % M) / M;
      tw1[(size_t)ka * L + b] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  for (int k = 0; k < M / 2; ++k) {
    double ang = -2.0 * kPiD * k / N;
    twp[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
  }
  int rc = build_filterbank(p, K, pl->h_fb);
  std::vector<float2> fbw;
  std::vector<int32_t> seg;
  pl->dense = dense_fb;
  if (rc == AAD_OK && !dense_fb) rc = band_filterbank(pl->h_fb, p.n_filt, K, fbw, seg);
  if (rc != AAD_OK) {
    delete pl;
    return rc;
  }
  // Filterbank program, two-tap banded form (see StftArgs).  Segment s = bins [seg[s], seg[s+1]) with
  // all-zero ends trimmed, first bin rounded down to a multiple of 4, in rounds of 4 bins.  Warps get
  // contiguous filter ranges balanced by cost; each warp's entry list is its segments wf0 .. wf1.
  std::vector<int2> fhdr;
  std::vector<float4> fw4;
  std::vector<int4> wprog(pl->warps, make_int4(0, 0, 0, 0));
  if (dense_fb) {
    // Dense form: entry e carries the full rows of filters 2e (x taps) and 2e + 1 (y taps) over all bins, in rounds of
    // 4 bins from bin 0; warps get contiguous entry ranges; wprog = {first filter, first header, entries, filters}.
    const int n_ent = (p.n_filt + 1) / 2, rounds = (K + 3) / 4;
    if (rounds > 0x7fff) {
      delete pl;
      return AAD_ERR_UNSUPPORTED;
    }
    // The phase is bound by shared-memory wavefronts: 4 per warp and round for the power group, 2 per entry and round
    // for the weights.  Four of the eight warps therefore take all the entries (5 each for 40 filters: 56 wavefronts
    // per round instead of 80); the other four go straight to the barrier.
    const int fb_warps = std::min(pl->warps, 4);
    for (int wi = 0; wi < fb_warps; ++wi) {
      const int e0 = (int)((long long)n_ent * wi / fb_warps), e1 = (int)((long long)n_ent * (wi + 1) / fb_warps);
      if (e1 <= e0) continue;
      std::vector<int> ents;
      for (int e = e0; e < e1; ++e) ents.push_back(e);
      while (ents.size() % fbu) ents.push_back(-1);
      wprog[wi] = make_int4(2 * e0, (int)fhdr.size(), (int)ents.size(), std::min(p.n_filt, 2 * e1) - 2 * e0);
      for (size_t b0 = 0; b0 < ents.size(); b0 += fbu) {
        const size_t w_off = fw4.size();
        fw4.resize(w_off + (size_t)rounds * fbu * 2, make_float4(0.f, 0.f, 0.f, 0.f));
        for (int u = 0; u < fbu; ++u) {
          const int e = ents[b0 + u];
          fhdr.push_back(make_int2(0 | (rounds << 16), (int)(w_off * sizeof(float4))));
          if (e < 0) continue;
          for (int g = 0; g < rounds; ++g) {
            float w[8];
            for (int i = 0; i < 4; ++i) {
              const int k = g * 4 + i;
              w[2 * i] = k < K ? pl->h_fb[(size_t)(2 * e) * K + k] : 0.f;
              w[2 * i + 1] = k < K && 2 * e + 1 < p.n_filt ? pl->h_fb[(size_t)(2 * e + 1) * K + k] : 0.f;
            }
            fw4[w_off + ((size_t)g * fbu + u) * 2] = make_float4(w[0], w[1], w[2], w[3]);
            fw4[w_off + ((size_t)g * fbu + u) * 2 + 1] = make_float4(w[4], w[5], w[6], w[7]);
          }
        }
      }
    }
  } else {
    const int nseg = p.n_filt + 1;
    std::vector<int> sk0(nseg), sk1(nseg), srounds(nseg);
    std::vector<double> cost(nseg);
    double tot = 0;
    for (int sgi = 0; sgi < nseg; ++sgi) {
      int k0 = seg[sgi], k1 = seg[sgi + 1];
      auto zero = [&](int k) { return fbw[k].x == 0.f && fbw[k].y == 0.f; };
      while (k0 < k1 && zero(k0)) ++k0;
      while (k1 > k0 && zero(k1 - 1)) --k1;
      sk0[sgi] = k0;
      sk1[sgi] = k1;
      srounds[sgi] = (k1 - (k0 & ~3) + 3) / 4;
      cost[sgi] = 8.0 * srounds[sgi] + 24.0;  // ~instructions: tap rounds + one emit
      tot += cost[sgi];
    }
    std::vector<int> wfilt(pl->warps + 1, p.n_filt);
    wfilt[0] = 0;
    double cum = 0;
    int j = 0;
    for (int wi = 1; wi < pl->warps; ++wi) {
      const double target = tot * wi / pl->warps;
      while (j < p.n_filt && cum + cost[j] * 0.5 < target) cum += cost[j++];
      wfilt[wi] = j;
    }
    const int row_limit = (K + 3) & ~3;  // bins + zeroed padding of a power row
    for (int wi = 0; wi < pl->warps; ++wi) {
      const int f0 = wfilt[wi], f1 = wfilt[wi + 1];
      if (f1 <= f0) continue;
      std::vector<int> ents;
      for (int sgi = f0; sgi <= f1; ++sgi) ents.push_back(sgi);
      while (ents.size() % fbu) ents.push_back(-1);  // empty entries
      wprog[wi] = make_int4(f0, (int)fhdr.size(), (int)ents.size(), f1 - f0);
      for (size_t e0 = 0; e0 < ents.size(); e0 += fbu) {  // one bundle
        int rounds = 0;
        for (int u = 0; u < fbu; ++u)
          rounds = std::max(rounds, ents[e0 + u] >= 0 ? srounds[ents[e0 + u]] : 0);
        const size_t w_off = fw4.size();
        fw4.resize(w_off + (size_t)rounds * fbu * 2, make_float4(0.f, 0.f, 0.f, 0.f));
        for (int u = 0; u < fbu; ++u) {
          const int sgi = ents[e0 + u];
          const int k0 = sgi >= 0 ? sk0[sgi] : 0, k1 = sgi >= 0 ? sk1[sgi] : 0;
          const int k0a = std::min(k0 & ~3, row_limit - 4 * rounds);
          if (k0a < 0 || 4 * k0a > 0xffff || rounds > 0x7fff) {
            delete pl;
            return AAD_ERR_UNSUPPORTED;
          }
          fhdr.push_back(make_int2((4 * k0a) | (rounds << 16), (int)(w_off * sizeof(float4))));
          for (int g = 0; g < rounds; ++g) {
            float w[8];
            for (int i = 0; i < 4; ++i) {
              const int k = k0a + g * 4 + i;
              const bool in = k >= k0 && k < k1;
              // the warp emits filters f0 .. f1 - 1: the falling taps of its first segment and the rising taps of its
              // last one belong to the neighbours' filters and are zeroed (k_stft_fb relies on it: entries that emit
              // nothing evaluate to the log floor)
              w[2 * i] = in && sgi != f1 ? fbw[k].x : 0.f;
              w[2 * i + 1] = in && sgi != f0 ? fbw[k].y : 0.f;
            }
            fw4[w_off + ((size_t)g * fbu + u) * 2] = make_float4(w[0], w[1], w[2], w[3]);
            fw4[w_off + ((size_t)g * fbu + u) * 2 + 1] = make_float4(w[4], w[5], w[6], w[7]);
          }
        }
      }
    }
  }
  pl->n_hdr = (int)fhdr.size();
  pl->n_w4 = (int)fw4.size();
  pl->k1_smem = k1_fixed + (size_t)((2 * fhdr.size() + 3) & ~3) * 4 + fw4.size() * sizeof(float4);
  if (AAD_TMA_STAGE && pl->L == 32)   // dev switch: staged samples of one tile (+ 16 floats of header)
    pl->k1_smem += (size_t)(16 + ((pl->tile * p.hop_length + p.n_fft + 8 + 3) & ~3)) * 4;
  // DCT-II ortho (scipy.fftpack.dct type 2 norm='ortho'), first n_ceps rows.  Device layout for K2: the
  // transposed table D^T[filter][coef] as mma.m16n8k8 B fragments, split into tf32 hi + lo parts:
  // [k-step][n-tile][lane = 4 g + t] = {b0 hi, b1 hi, b0 lo, b1 lo}, b0 = D^T[8 ks + t][8 nt + g],
  // b1 = D^T[8 ks + t + 4][8 nt + g]; plus the column sums (per-frame mean re-addition).
  std::vector<float4> dct_frag;
  std::vector<float> dct_colsum;
  if (p.n_ceps > 0) {
    const int Mf = p.n_filt;
    pl->h_dct.assign((size_t)p.n_ceps * Mf, 0.f);
    dct_colsum.assign((size_t)pl->n_tiles * 8, 0.f);
    for (int k = 0; k < p.n_ceps; ++k) {
      const double fk = k == 0 ? std::sqrt(1.0 / (4.0 * Mf)) : std::sqrt(1.0 / (2.0 * Mf));
      double colsum = 0.0;
      for (int m = 0; m < Mf; ++m) {
        float v = (float)(2.0 * fk * std::cos(kPiD * k * (2.0 * m + 1.0) / (2.0 * Mf)));
        pl->h_dct[(size_t)k * Mf + m] = v;
        colsum += (double)v;
      }
      dct_colsum[k] = (float)colsum;  // exact re-addition of the per-frame mean
    }
    auto dt = [&](int filt, int coef) { return filt < Mf && coef < p.n_ceps ? pl->h_dct[(size_t)coef * Mf + filt] : 0.f; };
    auto tf32_hi = [](float x) {  // cvt.rna.tf32.f32: round to nearest (ties away) on the low 13 mantissa bits
      uint32_t u;
      std::memcpy(&u, &x, 4);
      u = (u + 0x1000u) & 0xffffe000u;
      float r;
      std::memcpy(&r, &u, 4);
      return r;
    };
    dct_frag.resize((size_t)pl->n_ksteps * pl->n_tiles * 32);
    for (int ks = 0; ks < pl->n_ksteps; ++ks)
      for (int nt = 0; nt < pl->n_tiles; ++nt)
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, t = lane & 3;
          const float b0 = dt(8 * ks + t, 8 * nt + g), b1 = dt(8 * ks + t + 4, 8 * nt + g);
          const float h0 = tf32_hi(b0), h1 = tf32_hi(b1);
          dct_frag[((size_t)ks * pl->n_tiles + nt) * 32 + lane] = make_float4(h0, h1, b0 - h0, b1 - h1);
        }
  }
  savgol_taps(p.n_delta > 0 ? p.delta_width : 9, pl->taps[0], pl->taps[1]);

  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = upload(&pl->d_window, win_half);
  {
    // int16 input: the sample scale is folded into the window (exact for the power-of-two default)
    float sc = p.i16_scale != 0.f ? p.i16_scale : (p.kind == AAD_KIND_LFCC ? 1.0f : 1.0f / 32768.0f);
    pl->p.i16_scale = sc;
    std::vector<float> win_i16(win_half);
    for (auto& x : win_i16) x *= sc;
    if (e == cudaSuccess) e = upload(&pl->d_window_i16, win_i16);
  }
  if (e == cudaSuccess) e = upload(&pl->d_tw1, tw1);
  if (e == cudaSuccess) e = upload(&pl->d_twp, twp);
  if (e == cudaSuccess) e = upload(&pl->d_filt_hdr, fhdr);
  if (e == cudaSuccess) e = upload(&pl->d_filt_w, fw4);
  if (e == cudaSuccess) e = upload(&pl->d_warp_prog, wprog);
  if (e == cudaSuccess) e = upload(&pl->d_dct_frag, dct_frag);
  if (e == cudaSuccess) e = upload(&pl->d_dct_colsum, dct_colsum);
  // opt in to the large dynamic shared memory of every kernel variant this plan can launch
  // (the attribute is per function, not per plan: opt in to the device maximum)
  int optin = 0;
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  pl->smem_optin = optin;
  if (pl->k1_smem > (size_t)optin) {
    aad_plan_destroy(pl);
    return AAD_ERR_UNSUPPORTED;
  }
  for (int mode = 0; mode < 3 && e == cudaSuccess; ++mode)
    for (int pre = 0; pre < 2 && e == cudaSuccess; ++pre) {
      const void* fn = (const void*)pick_stft(L, pl->tile, mode, pre != 0, false, pl->dense);
      if (fn) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
      const void* fp = pl->dense ? nullptr : (const void*)pick_stft(L, pl->tile, mode, pre != 0, true);
      if (fp && e == cudaSuccess) e = cudaFuncSetAttribute(fp, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    }
  if (e == cudaSuccess && pl->need_ws_E) {
    if (cep_smem_bytes(pl, CEP_TS) > (size_t)optin) {
      aad_plan_destroy(pl);
      return AAD_ERR_UNSUPPORTED;
    }
    e = cudaFuncSetAttribute((const void*)pick_cep(pl->cep_nt, CEP_TS), cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute((const void*)pick_cep(pl->cep_nt, 64), cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
  }
  if (e != cudaSuccess) {
    aad_plan_destroy(pl);
    return AAD_ERR_CUDA;
  }
  *out = pl;
  return AAD_OK;
}

static int frames_for(const aad_params& p, int64_t len) {
  if (len <= 0) return 0;
  if (p.center) return (int)(1 + len / p.hop_length);
  return len >= p.win_length ? (int)((len - p.win_length) / p.hop_length + 1) : 0;
}

struct WsLayout {
  size_t off_frame_off, off_len, off_nf, off_max, off_zn, off_tile, off_E, off_feat, total;
  int t_ws;
  int max_tiles;
};
static WsLayout ws_layout(const aad_plan* pl, int B, int t_max) {
  WsLayout w;
  size_t o = 0;
  w.off_frame_off = o; o = align_up(o + (size_t)(B + 1) * 4, 256);
  w.off_len = o;       o = align_up(o + (size_t)B * 4, 256);
  w.off_nf = o;        o = align_up(o + (size_t)B * 4, 256);
  w.off_max = o;       o = align_up(o + (size_t)B * 4, 256);
  w.off_zn = o;        if (pl->p.znorm) o = align_up(o + (size_t)B * 2 * sizeof(double), 256);
  w.max_tiles = (int)(((long long)B * t_max + pl->tile - 1) / pl->tile);
  w.off_tile = o;      o = align_up(o + (size_t)(w.max_tiles + 1) * 48, 256);
  w.t_ws = (t_max + 31) / 32 * 32;
  w.off_E = o;
  if (pl->need_ws_E) o = align_up(o + (size_t)B * pl->p.n_filt * w.t_ws * 4, 256);
  w.off_feat = o;
  if (pl->need_ws_feat) o = align_up(o + (size_t)B * pl->c_out * w.t_ws * 4, 256);
  w.total = o;
  return w;
}

int aad_query(const aad_plan* pl, int B, int64_t max_len, int32_t* t_max, int32_t* c_out,
              size_t* workspace_bytes) {
  if (!pl || B < 0 || max_len < 0) return AAD_ERR_INVALID_ARG;
  if (max_len > 0x7fffffffLL) return AAD_ERR_UNSUPPORTED;
  int T = frames_for(pl->p, max_len);
  if ((double)B * (double)std::max(T, 1) > 2.0e9) return AAD_ERR_UNSUPPORTED;
  if (t_max) *t_max = T;
  if (c_out) *c_out = pl->c_out;
  if (workspace_bytes) *workspace_bytes = ws_layout(pl, std::max(B, 1), std::max(T, 1)).total;
  return AAD_OK;
}

int aad_plan_launches(const aad_plan* pl) {
  if (!pl) return AAD_ERR_INVALID_ARG;
  // prepare, stft_fb, cepstra | finalize (only when there is a reference or a floor to apply), [time_mean],
  // [znorm stats + apply]
  const aad_params& p = pl->p;
  const bool fin = pl->need_ws_E || (p.log_type == AAD_LOG_DB10 && (p.ref_type == AAD_REF_UTT_MAX || p.top_db >= 0.f));
  return 2 + (fin ? 1 : 0) + (pl->need_ws_feat ? 1 : 0) + (p.znorm ? 2 : 0);
}

// `pl2` (optional): a second plan over the SAME STFT whose filter bank runs in the same k_stft_fb launch (one
// STFT, two features); it must be a plain log filter-bank plan (no DCT, deltas, time mean or z-norm; CT layout).
struct PairArgs {
  const aad_plan* pl2 = nullptr;
  float* out2 = nullptr;
  int64_t out2_stride_b = 0;
  void* ws2 = nullptr;
  size_t ws2_bytes = 0;
};

static size_t prog_smem_bytes(const aad_plan* pl) { return (size_t)((2 * pl->n_hdr + 3) & ~3) * 4 + (size_t)pl->n_w4 * sizeof(float4); }

static int extract_impl(const aad_plan* pl, const void* wav, int wav_dtype, int64_t wav_stride,
                        const int64_t* row_off, const int32_t* lengths, int B, int64_t max_len, float* out,
                        int64_t out_stride_b, int32_t t_alloc, int32_t* n_frames, int32_t* status,
                        void* workspace, size_t workspace_bytes, void* stream_, const PairArgs& pair = PairArgs()) {
  if (!pl || !wav || !lengths || !out || !n_frames || !status || !workspace) return AAD_ERR_INVALID_ARG;
  if (B <= 0 || max_len <= 0 || (!row_off && max_len > wav_stride) || t_alloc <= 0) return AAD_ERR_INVALID_ARG;
  if (wav_dtype != AAD_F32 && wav_dtype != AAD_I16) return AAD_ERR_INVALID_ARG;
  const aad_params& p = pl->p;
  cudaStream_t stream = (cudaStream_t)stream_;
  int t_max = 0;
  size_t need = 0;
  int rc = aad_query(pl, B, max_len, &t_max, nullptr, &need);
  if (rc != AAD_OK) return rc;
  if (workspace_bytes < need) return AAD_ERR_WORKSPACE;
  const int t_cap = std::min<int>(t_alloc, std::max(t_max, 1));
  const WsLayout w = ws_layout(pl, B, std::max(t_max, 1));
  char* ws = (char*)workspace;
  int32_t* d_frame_off = (int32_t*)(ws + w.off_frame_off);
  int32_t* d_len = (int32_t*)(ws + w.off_len);
  int32_t* d_nf = (int32_t*)(ws + w.off_nf);
  int32_t* d_max = (int32_t*)(ws + w.off_max);
  float* d_E = (float*)(ws + w.off_E);
  float* d_feat = (float*)(ws + w.off_feat);

  if (p.time_mean) {
    if (out_stride_b == 0) out_stride_b = pl->c_out;
  } else if (out_stride_b == 0) {
    out_stride_b = (int64_t)pl->c_out * t_alloc;
  }

  // K0
  PrepArgs pa;
  pa.lengths = lengths; pa.B = B; pa.max_len = max_len;
  pa.hop = p.hop_length; pa.win_len = p.win_length; pa.center = p.center;
  pa.n_delta = p.n_delta; pa.delta_width = p.delta_width;
  pa.t_alloc = p.time_mean ? w.t_ws : t_cap;
  pa.n_frames = n_frames; pa.status = status; pa.len_c = d_len; pa.nf_eff = d_nf;
  pa.frame_off = d_frame_off; pa.utt_max = d_max;
  // paired call: the second plan shares the STFT; its only state is its own per-utterance maximum
  const aad_plan* pl2 = pair.pl2;
  int32_t* d_max2 = nullptr;
  int64_t out2_stride_b = pair.out2_stride_b;
  if (pl2) {
    const aad_params& q = pl2->p;
    if (!pair.out2 || !pair.ws2) return AAD_ERR_INVALID_ARG;
    if (q.n_fft != p.n_fft || q.hop_length != p.hop_length || q.win_length != p.win_length || q.window != p.window ||
        q.center != p.center || q.pre_emph != p.pre_emph || q.quantize_i16 != p.quantize_i16 ||
        q.sample_rate != p.sample_rate || q.i16_scale != p.i16_scale || pl2->warps != pl->warps || pl2->L != pl->L)
      return AAD_ERR_PAIR;
    if (pl2->need_ws_E || pl2->need_ws_feat || q.znorm || q.layout != AAD_LAYOUT_CT || pl2->dense) return AAD_ERR_PAIR;
    size_t need2 = 0;
    if ((rc = aad_query(pl2, B, max_len, nullptr, nullptr, &need2)) != AAD_OK) return rc;
    if (pair.ws2_bytes < need2) return AAD_ERR_WORKSPACE;
    d_max2 = (int32_t*)((char*)pair.ws2 + ws_layout(pl2, B, std::max(t_max, 1)).off_max);
    if (out2_stride_b == 0) out2_stride_b = (int64_t)pl2->c_out * t_alloc;
    if (pl->k1_smem + prog_smem_bytes(pl2) > (size_t)pl->smem_optin) return AAD_ERR_UNSUPPORTED;
  }
  pa.utt_max2 = d_max2;
  const int mode = wav_dtype == AAD_I16 ? IN_I16 : (p.quantize_i16 ? IN_F32_Q16 : IN_F32);
  const int k1_tile = pl->tile;
  pa.tile_rec = (int4*)(ws + w.off_tile); pa.tile = k1_tile; pa.max_tiles = w.max_tiles;
  pa.row_off = reinterpret_cast<const long long*>(row_off); pa.wav_stride = wav_stride;
  pa.zn_stats = p.znorm ? (double*)(ws + w.off_zn) : nullptr;
  const bool prof = pl->profile;
  if (prof) cudaEventRecord(pl->ev[0], stream);
  (void)cudaGetLastError();  // clear stale non-sticky state left by earlier calls in this thread
  k_prepare<<<(B + kPrepBlock - 1) / kPrepBlock, kPrepBlock, 0, stream>>>(pa);
  LAUNCH_CHECK("k_prepare launch");
  if (prof) cudaEventRecord(pl->ev[1], stream);

  // K1
  StftArgs sa;
  sa.wav = wav; sa.wav_stride = wav_stride; sa.row_off = reinterpret_cast<const long long*>(row_off); sa.len_c = d_len; sa.frame_off = d_frame_off; sa.B = B;
  sa.hop = p.hop_length; sa.s_off = p.center ? p.n_fft / 2 : 0;
  sa.win_off = p.center ? (p.n_fft - p.win_length) / 2 : 0; sa.win_len = std::min(p.win_length, p.n_fft);
  sa.pre_emph = p.pre_emph;
  sa.window = wav_dtype == AAD_I16 ? pl->d_window_i16 : pl->d_window; sa.tw1 = pl->d_tw1; sa.twp = pl->d_twp;
  sa.filt_hdr = pl->d_filt_hdr; sa.filt_w = pl->d_filt_w; sa.n_hdr = pl->n_hdr; sa.n_w4 = pl->n_w4;
  sa.warp_prog = pl->d_warp_prog; sa.tile_rec = pa.tile_rec; sa.n_filt = p.n_filt;
  sa.log_type = p.log_type; sa.amin = p.amin; sa.eps = 2.220446049250313e-16f; sa.spec_mag = p.spectrum == AAD_SPEC_MAGNITUDE;

  if (pl->need_ws_E) {
    sa.E = d_E; sa.e_stride_b = (long long)p.n_filt * w.t_ws; sa.e_stride_f = w.t_ws;
  } else {
    sa.E = out; sa.e_stride_b = out_stride_b; sa.e_stride_f = t_alloc;
  }
  sa.utt_max = p.log_type == AAD_LOG_DB10 ? d_max : nullptr;
  sa.status = status;
  sa.filt_hdr2 = nullptr; sa.filt_w2 = nullptr; sa.n_hdr2 = 0; sa.n_w42 = 0; sa.warp_prog2 = nullptr;
  sa.log_type2 = 0; sa.amin2 = 0.f; sa.E2 = nullptr; sa.e2_stride_b = 0; sa.e2_stride_f = 0; sa.utt_max2 = nullptr;
  if (pl2) {
    const aad_params& q = pl2->p;
    sa.filt_hdr2 = pl2->d_filt_hdr; sa.filt_w2 = pl2->d_filt_w; sa.n_hdr2 = pl2->n_hdr; sa.n_w42 = pl2->n_w4;
    sa.warp_prog2 = pl2->d_warp_prog; sa.log_type2 = q.log_type; sa.amin2 = q.amin;
    sa.E2 = pair.out2; sa.e2_stride_b = out2_stride_b; sa.e2_stride_f = t_alloc;
    sa.utt_max2 = q.log_type == AAD_LOG_DB10 ? d_max2 : nullptr;
  }
  if (pl->dense && pl2) return AAD_ERR_PAIR;
  stft_kernel_t kern = pick_stft(pl->L, pl->tile, mode, p.pre_emph != 0.f, pl2 != nullptr, pl->dense);
  if (!kern) return pl2 ? AAD_ERR_PAIR : AAD_ERR_UNSUPPORTED;
  const long long max_tiles = w.max_tiles;
  // persistent CTAs take tiles round-robin: with few tiles per CTA shrink the grid so that every CTA gets the same
  // number (802 tiles on 592 slots would leave 382 CTAs idle during the second round; 401 CTAs x 2 tiles share the
  // SMs evenly)
  const long long slots = (long long)pl->sm_count * pl->ctas, nt1 = std::max<long long>(max_tiles, 1);
  const long long per_cta = (nt1 + slots - 1) / slots;
  const int grid1 = (int)std::min<long long>(slots, (nt1 + per_cta - 1) / per_cta);
  launch_pdl(kern, dim3(grid1), dim3(pl->warps * 32), pl->k1_smem + (pl2 ? prog_smem_bytes(pl2) : 0), stream, !prof, sa);
  LAUNCH_CHECK("k_stft_fb launch");
  if (prof) cudaEventRecord(pl->ev[2], stream);

  // K2
  if (pl->need_ws_E) {
    CepArgs ca;
    ca.E = d_E; ca.e_stride_b = sa.e_stride_b; ca.e_stride_f = sa.e_stride_f;
    ca.nf_eff = d_nf; ca.utt_max = d_max; ca.n_filt = p.n_filt;
    ca.log_type = p.log_type; ca.ref_type = p.ref_type; ca.top_db = p.top_db;
    ca.n_ceps = p.n_ceps; ca.n_ksteps = pl->n_ksteps; ca.n_tiles = pl->n_tiles;
    ca.dct_frag = pl->d_dct_frag; ca.dct_colsum = pl->d_dct_colsum;
    ca.n_delta = p.n_delta; ca.width = p.n_delta > 0 ? p.delta_width : 1;
    std::memcpy(ca.taps, pl->taps, sizeof(ca.taps));
    if (p.time_mean) {
      ca.out = d_feat; ca.out_stride_b = (long long)pl->c_out * w.t_ws;
      ca.out_stride_c = w.t_ws; ca.out_stride_t = 1;
    } else if (p.layout == AAD_LAYOUT_CT) {
      ca.out = out; ca.out_stride_b = out_stride_b; ca.out_stride_c = t_alloc; ca.out_stride_t = 1;
    } else {
      ca.out = out; ca.out_stride_b = out_stride_b; ca.out_stride_c = 1; ca.out_stride_t = pl->c_out;
    }
    const int cep_ts = cep_tile(t_max);
    ca.tile_out = cep_ts - (p.n_delta > 0 ? 2 * (p.delta_width / 2) : 0);
    const int gx = t_max <= cep_ts ? 1 : (t_max + ca.tile_out - 1) / ca.tile_out;
    if ((long long)B * gx > 0x7fffffffLL) return AAD_ERR_UNSUPPORTED;
    ca.tiles_per_utt = gx;
    launch_pdl(pick_cep(pl->cep_nt, cep_ts), dim3((unsigned)((long long)B * gx)), dim3(2 * cep_ts), cep_smem_bytes(pl, cep_ts), stream, !prof, ca);
    LAUNCH_CHECK("k_cepstra launch");
    if (prof) cudaEventRecord(pl->ev[3], stream);
    if (p.time_mean) {
      dim3 gm(B, (pl->c_out + 3) / 4);
      k_time_mean<<<gm, 128, 0, stream>>>(d_feat, ca.out_stride_b, w.t_ws, d_nf, pl->c_out, out, out_stride_b);
    }
  } else {
    FinArgs fa;
    fa.out = out; fa.stride_b = out_stride_b; fa.stride_f = t_alloc; fa.nf_eff = d_nf; fa.utt_max = d_max; fa.utt_max_f = nullptr;
    fa.n_filt = p.n_filt; fa.ref_type = p.ref_type; fa.top_db = p.top_db;
    if (p.log_type == AAD_LOG_DB10 && (p.ref_type == AAD_REF_UTT_MAX || p.top_db >= 0.f)) {  // else: identity
      if ((rc = launch_db_finalize(fa, B, t_max, stream, !prof)) != AAD_OK) return rc;
    }
    if (prof) cudaEventRecord(pl->ev[3], stream);
  }
  if (pl2) {
    const aad_params& q = pl2->p;
    if (q.log_type == AAD_LOG_DB10 && (q.ref_type == AAD_REF_UTT_MAX || q.top_db >= 0.f)) {
      FinArgs fb;
      fb.out = pair.out2; fb.stride_b = out2_stride_b; fb.stride_f = t_alloc; fb.nf_eff = d_nf; fb.utt_max = d_max2;
      fb.utt_max_f = nullptr; fb.n_filt = q.n_filt; fb.ref_type = q.ref_type; fb.top_db = q.top_db;
      if ((rc = launch_db_finalize(fb, B, t_max, stream, false)) != AAD_OK) return rc;
    }
  }
  if (p.znorm) {
    ZnArgs za;
    za.out = out; za.stride_b = out_stride_b; za.nf_eff = d_nf; za.C = pl->c_out;
    if (p.layout == AAD_LAYOUT_CT) { za.stride_c = t_alloc; za.stride_t = 1; }
    else { za.stride_c = 1; za.stride_t = pl->c_out; }
    za.n_chunks = (int)(((long long)pl->c_out * std::max(t_max, 1) + ZN_CHUNK - 1) / ZN_CHUNK);
    za.stats = pa.zn_stats;
    const long long nblk = (long long)B * za.n_chunks;
    if (nblk > 0x7fffffffLL) return AAD_ERR_UNSUPPORTED;
    k_znorm<false><<<(unsigned)nblk, 256, 0, stream>>>(za);
    k_znorm<true><<<(unsigned)nblk, 256, 0, stream>>>(za);
  }
  if (prof) cudaEventRecord(pl->ev[4], stream);
  LAUNCH_CHECK("epilogue launch");
  return AAD_OK;
}

int aad_extract(const aad_plan* pl, const void* wav, int wav_dtype, int64_t wav_stride,
                const int32_t* lengths, int B, int64_t max_len, float* out, int64_t out_stride_b,
                int32_t t_alloc, int32_t* n_frames, int32_t* status, void* workspace,
                size_t workspace_bytes, void* stream) {
  return extract_impl(pl, wav, wav_dtype, wav_stride, nullptr, lengths, B, max_len, out, out_stride_b, t_alloc,
                      n_frames, status, workspace, workspace_bytes, stream);
}

int aad_extract_pair(const aad_plan* pl, const aad_plan* pl2, const void* wav, int wav_dtype, int64_t wav_stride,
                     const int64_t* row_off, const int32_t* lengths, int B, int64_t max_len, float* out,
                     int64_t out_stride_b, float* out2, int64_t out2_stride_b, int32_t t_alloc, int32_t* n_frames,
                     int32_t* status, void* workspace, size_t workspace_bytes, void* workspace2,
                     size_t workspace2_bytes, void* stream) {
  if (!pl2) return AAD_ERR_INVALID_ARG;
  PairArgs pair;
  pair.pl2 = pl2; pair.out2 = out2; pair.out2_stride_b = out2_stride_b; pair.ws2 = workspace2; pair.ws2_bytes = workspace2_bytes;
  return extract_impl(pl, wav, wav_dtype, wav_stride, row_off, lengths, B, max_len, out, out_stride_b, t_alloc,
                      n_frames, status, workspace, workspace_bytes, stream, pair);
}

int aad_extract_indexed(const aad_plan* pl, const void* wav, int wav_dtype, const int64_t* row_off,
                        const int32_t* lengths, int B, int64_t max_len, float* out, int64_t out_stride_b,
                        int32_t t_alloc, int32_t* n_frames, int32_t* status, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (!row_off) return AAD_ERR_INVALID_ARG;
  return extract_impl(pl, wav, wav_dtype, 0, row_off, lengths, B, max_len, out, out_stride_b, t_alloc,
                      n_frames, status, workspace, workspace_bytes, stream);
}

#define AAD_KIND_ALIAS(NAME, KIND)                                                                        \
  int NAME(const aad_plan* pl, const void* wav, int wav_dtype, int64_t wav_stride, const int32_t* lengths, \
           int B, int64_t max_len, float* out, int64_t out_stride_b, int32_t t_alloc, int32_t* n_frames,  \
           int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {                      \
    if (!pl) return AAD_ERR_INVALID_ARG;                                                                  \
    if (pl->p.kind != KIND) return AAD_ERR_KIND;                                                          \
    return aad_extract(pl, wav, wav_dtype, wav_stride, lengths, B, max_len, out, out_stride_b, t_alloc,   \
                       n_frames, status, workspace, workspace_bytes, stream);                             \
  }
AAD_KIND_ALIAS(aad_logmel, AAD_KIND_LOGMEL)
AAD_KIND_ALIAS(aad_mfcc, AAD_KIND_MFCC)
AAD_KIND_ALIAS(aad_lfcc, AAD_KIND_LFCC)
AAD_KIND_ALIAS(aad_gtcc, AAD_KIND_GTCC)

int aad_delta(const float* x, const int32_t* n_frames, int B, int C, int32_t t_stride, int width,
              int order, float* out, void* stream) {
  if (!x || !n_frames || !out || B <= 0 || C <= 0 || t_stride <= 0) return AAD_ERR_INVALID_ARG;
  if (width < 3 || width > CEP_MAXW || width % 2 == 0 || order < 1 || order > 2) return AAD_ERR_INVALID_ARG;
  if (C > 65535) return AAD_ERR_UNSUPPORTED;
  DeltaArgs da;
  float t1[CEP_MAXW], t2[CEP_MAXW];
  savgol_taps(width, t1, t2);
  std::memcpy(da.taps, order == 1 ? t1 : t2, sizeof(da.taps));
  da.x = x; da.out = out; da.n_frames = n_frames; da.C = C; da.t_stride = t_stride; da.width = width;
  dim3 grid(B, C, std::max(1, std::min(64, (t_stride + 255) / 256)));
  (void)cudaGetLastError();
  k_delta<<<grid, 256, 0, (cudaStream_t)stream>>>(da);
  LAUNCH_CHECK("k_delta launch");
  return AAD_OK;
}

int aad_db_reference(float* x, int64_t stride_b, int32_t stride_f, const int32_t* n_frames, const float* utt_max,
                     int B, int n_filt, int32_t t_max, int ref_type, float top_db, void* stream) {
  if (!x || !n_frames || !utt_max || B <= 0 || n_filt <= 0 || t_max <= 0 || stride_f < t_max) return AAD_ERR_INVALID_ARG;
  if (ref_type != AAD_REF_ONE && ref_type != AAD_REF_UTT_MAX) return AAD_ERR_INVALID_ARG;
  if (stride_b == 0) stride_b = (int64_t)n_filt * stride_f;
  FinArgs fa;
  fa.out = x; fa.stride_b = stride_b; fa.stride_f = stride_f; fa.nf_eff = n_frames; fa.utt_max = nullptr;
  fa.utt_max_f = utt_max; fa.n_filt = n_filt; fa.ref_type = ref_type; fa.top_db = top_db;
  (void)cudaGetLastError();
  if (int rc2 = launch_db_finalize(fa, B, t_max, (cudaStream_t)stream, false)) return rc2;
  LAUNCH_CHECK("k_db_finalize launch");
  return AAD_OK;
}

int aad_scaler_accumulate(const float* x, int64_t n_rows, int32_t W, int64_t row_stride, double* stats,
                          void* stream) {
  if (!x || !stats || n_rows < 0 || W <= 0 || row_stride < W) return AAD_ERR_INVALID_ARG;
  if (n_rows == 0) return AAD_OK;
  const long long nblk = (n_rows + SC_ROWS - 1) / SC_ROWS;
  if (nblk > 0x7fffffffLL) return AAD_ERR_UNSUPPORTED;
  (void)cudaGetLastError();
  k_col_stats<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(x, n_rows, W, row_stride, stats, nullptr, nullptr, 1);
  LAUNCH_CHECK("k_col_stats launch");
  return AAD_OK;
}

int aad_scaler_accumulate_ragged(const float* x, int B, int32_t rows_per_utt, int32_t W, int64_t row_stride,
                                 const int32_t* n_frames, const int32_t* status, double* stats, void* stream) {
  if (!x || !stats || !n_frames || B <= 0 || rows_per_utt <= 0 || W <= 0 || row_stride < W) return AAD_ERR_INVALID_ARG;
  const long long n_rows = (long long)B * rows_per_utt;
  const long long nblk = (n_rows + SC_ROWS - 1) / SC_ROWS;
  if (nblk > 0x7fffffffLL) return AAD_ERR_UNSUPPORTED;
  (void)cudaGetLastError();
  k_col_stats<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(x, n_rows, W, row_stride, stats, n_frames, status, rows_per_utt);
  LAUNCH_CHECK("k_col_stats launch");
  return AAD_OK;
}

int aad_scaler_apply(float* x, int64_t n_rows, int32_t W, int64_t row_stride, const float* mean,
                     const float* inv_scale, void* stream) {
  if (!x || !mean || !inv_scale || n_rows < 0 || W <= 0 || row_stride < W) return AAD_ERR_INVALID_ARG;
  if (n_rows == 0) return AAD_OK;
  const long long nblk = (n_rows + SC_ROWS - 1) / SC_ROWS;
  if (nblk > 0x7fffffffLL) return AAD_ERR_UNSUPPORTED;
  (void)cudaGetLastError();
  k_col_apply<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(x, n_rows, W, row_stride, mean, inv_scale);
  LAUNCH_CHECK("k_col_apply launch");
  return AAD_OK;
}

int64_t aad_plan_table(const aad_plan* pl, int which, float* host_out, int64_t capacity) {
  if (!pl) return AAD_ERR_INVALID_ARG;
  const float* src = nullptr;
  int64_t n = 0;
  float tapbuf[2 * CEP_MAXW];
  switch (which) {
    case AAD_TABLE_WINDOW: src = pl->h_window.data(); n = (int64_t)pl->h_window.size(); break;
    case AAD_TABLE_FILTERBANK: src = pl->h_fb.data(); n = (int64_t)pl->h_fb.size(); break;
    case AAD_TABLE_DCT: src = pl->h_dct.data(); n = (int64_t)pl->h_dct.size(); break;
    case AAD_TABLE_DELTA_TAPS: {
      const int wdt = pl->p.n_delta > 0 ? pl->p.delta_width : 9;
      for (int i = 0; i < wdt; ++i) {
        tapbuf[i] = pl->taps[0][i];
        tapbuf[wdt + i] = pl->taps[1][i];
      }
      src = tapbuf;
      n = 2 * wdt;
      break;
    }
    default: return AAD_ERR_INVALID_ARG;
  }
  if (host_out) {
    if (capacity < n) return AAD_ERR_INVALID_ARG;
    std::memcpy(host_out, src, (size_t)n * 4);
  }
  return n;
}

int aad_plan_set_profiling(aad_plan* pl, int enable) {
  if (!pl) return AAD_ERR_INVALID_ARG;
  DeviceGuard guard(pl->device);
  CUDA_TRY(guard.err);
  if (enable)
    for (auto& e : pl->ev)
      if (!e) CUDA_TRY(cudaEventCreate(&e));
  pl->profile = enable != 0;
  return AAD_OK;
}

int aad_plan_kernel_times(const aad_plan* pl, float* ms_out) {
  if (!pl || !ms_out || !pl->ev[0]) return AAD_ERR_INVALID_ARG;
  for (int i = 0; i < 4; ++i) {
    ms_out[i] = 0.f;
    if (cudaEventElapsedTime(&ms_out[i], pl->ev[i], pl->ev[i + 1]) != cudaSuccess) return AAD_ERR_CUDA;
  }
  return AAD_OK;
}

int aad_fp32_peak(int device, int iters, double* tflops_out) {
  if (!tflops_out || iters <= 0) return AAD_ERR_INVALID_ARG;
  DeviceGuard guard(device);
  CUDA_TRY(guard.err);
  int sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  float* sink = nullptr;
  CUDA_TRY(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sms * 8, block = 256;
  k_fma_peak<<<grid, block>>>(sink, 16);  // warm-up
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_fma_peak<<<grid, block>>>(sink, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8 * 16 * (double)iters * grid * block;
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops_out = best;
  return cudaGetLastError() == cudaSuccess ? AAD_OK : AAD_ERR_CUDA;
}


// ---- host-buffer path: chunked H2D -> kernels -> D2H on three internal streams ----
static int ensure(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return AAD_OK;
  cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  if (cudaMalloc(p, need) != cudaSuccess) return AAD_ERR_CUDA;
  *cap = need;
  return AAD_OK;
}

// internal buffers of the host path: three {stream, device wav / out / lengths / n_frames / status / workspace}
// sets sized for one chunk, and the pinned staging of the per-utterance arrays.  Grown, never shrunk.
static int host_reserve_impl(aad_plan* pl, int wav_dtype, int B, int64_t max_len, int32_t t_alloc, int chunk_utts,
                             int* chunk_out) {
  const size_t esz = wav_dtype == AAD_I16 ? 2 : 4;
  const int64_t row_out = pl->p.time_mean ? pl->c_out : (int64_t)pl->c_out * t_alloc;
  if (chunk_utts <= 0) {
    // ~32 MB of samples per chunk keeps both copy engines and the SMs busy
    chunk_utts = (int)std::max<int64_t>(1, (32ll << 20) / (int64_t)(max_len * esz));
  }
  chunk_utts = std::min(chunk_utts, B);
  *chunk_out = chunk_utts;
  int t_max = 0;
  size_t ws_need = 0;
  int rc = aad_query(pl, chunk_utts, max_len, &t_max, nullptr, &ws_need);
  if (rc != AAD_OK) return rc;
  for (auto& h : pl->hb) {
    if (!h.stream) CUDA_TRY(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking));
    if ((rc = ensure(&h.d_wav, &h.wav_bytes, (size_t)chunk_utts * max_len * esz)) != AAD_OK) return rc;
    if ((rc = ensure((void**)&h.d_out, &h.out_bytes, (size_t)chunk_utts * row_out * 4)) != AAD_OK) return rc;
    if ((rc = ensure(&h.d_ws, &h.ws_bytes, ws_need)) != AAD_OK) return rc;
    if (h.cap_b < chunk_utts) {
      cudaFree(h.d_len); cudaFree(h.d_nf); cudaFree(h.d_st);
      h.d_len = h.d_nf = h.d_st = nullptr;
      h.cap_b = 0;
      CUDA_TRY(cudaMalloc((void**)&h.d_len, (size_t)chunk_utts * 4));
      CUDA_TRY(cudaMalloc((void**)&h.d_nf, (size_t)chunk_utts * 4));
      CUDA_TRY(cudaMalloc((void**)&h.d_st, (size_t)chunk_utts * 4));
      h.cap_b = chunk_utts;
    }
  }
  if (pl->h_stage_cap < B) {
    if (pl->h_stage) cudaFreeHost(pl->h_stage);
    pl->h_stage = nullptr;
    pl->h_stage_cap = 0;
    CUDA_TRY(cudaHostAlloc((void**)&pl->h_stage, (size_t)3 * B * sizeof(int32_t), cudaHostAllocDefault));
    pl->h_stage_cap = B;
  }
  return AAD_OK;
}

int aad_host_reserve(aad_plan* pl, int wav_dtype, int B, int64_t max_len, int32_t t_alloc, int chunk_utts) {
  if (!pl || B <= 0 || max_len <= 0 || t_alloc <= 0) return AAD_ERR_INVALID_ARG;
  if (wav_dtype != AAD_F32 && wav_dtype != AAD_I16) return AAD_ERR_INVALID_ARG;
  DeviceGuard guard(pl->device);
  CUDA_TRY(guard.err);
  int chunk = 0;
  return host_reserve_impl(pl, wav_dtype, B, max_len, t_alloc, chunk_utts, &chunk);
}

int aad_host_alloc(void** ptr, size_t bytes, int write_combined) {
  if (!ptr || bytes == 0) return AAD_ERR_INVALID_ARG;
  *ptr = nullptr;
  CUDA_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
  return AAD_OK;
}

int aad_host_free(void* ptr) {
  if (ptr) CUDA_TRY(cudaFreeHost(ptr));
  return AAD_OK;
}

int aad_extract_host(aad_plan* pl, const void* wav_host, int wav_dtype, int64_t wav_stride,
                     const int32_t* lengths_host, int B, int64_t max_len, float* out_host,
                     int64_t out_stride_b, int32_t t_alloc, int32_t* n_frames_host,
                     int32_t* status_host, int chunk_utts) {
  if (!pl || !wav_host || !lengths_host || !out_host || !n_frames_host || !status_host)
    return AAD_ERR_INVALID_ARG;
  if (B <= 0 || max_len <= 0 || max_len > wav_stride || t_alloc <= 0) return AAD_ERR_INVALID_ARG;
  if (wav_dtype != AAD_F32 && wav_dtype != AAD_I16) return AAD_ERR_INVALID_ARG;
  DeviceGuard guard(pl->device);
  CUDA_TRY(guard.err);
  const size_t esz = wav_dtype == AAD_I16 ? 2 : 4;
  const int64_t row_out = pl->p.time_mean ? pl->c_out : (int64_t)pl->c_out * t_alloc;
  if (out_stride_b == 0) out_stride_b = row_out;
  int rc = host_reserve_impl(pl, wav_dtype, B, max_len, t_alloc, chunk_utts, &chunk_utts);  // no-op once warm
  if (rc != AAD_OK) return rc;
  int32_t* st_len = pl->h_stage;
  int32_t* st_nf = pl->h_stage + pl->h_stage_cap;
  int32_t* st_st = pl->h_stage + 2 * (size_t)pl->h_stage_cap;
  std::memcpy(st_len, lengths_host, (size_t)B * sizeof(int32_t));
  // on any failure: wait for the copies already in flight (they target the caller's buffers and the pinned stage)
  auto drain = [&](int code) {
    for (auto& h : pl->hb)
      if (h.stream) cudaStreamSynchronize(h.stream);
    return code;
  };
#define HOST_TRY(expr)                                              \
  do {                                                              \
    cudaError_t e__ = (expr);                                       \
    if (e__ != cudaSuccess) return drain(cuda_fail(e__, #expr));    \
  } while (0)
  int ci = 0;
  for (int b0 = 0; b0 < B; b0 += chunk_utts, ++ci) {
    aad_plan::HostBuf& h = pl->hb[ci % 3];
    const int nb = std::min(chunk_utts, B - b0);
    const char* src = (const char*)wav_host + (size_t)b0 * wav_stride * esz;
    if (wav_stride == max_len)
      HOST_TRY(cudaMemcpyAsync(h.d_wav, src, (size_t)nb * max_len * esz, cudaMemcpyHostToDevice, h.stream));
    else
      HOST_TRY(cudaMemcpy2DAsync(h.d_wav, (size_t)max_len * esz, src, (size_t)wav_stride * esz,
                                 (size_t)max_len * esz, nb, cudaMemcpyHostToDevice, h.stream));
    HOST_TRY(cudaMemcpyAsync(h.d_len, st_len + b0, (size_t)nb * 4, cudaMemcpyHostToDevice, h.stream));
    // rows with non-zero status are left untouched by the kernels: start from zeros
    HOST_TRY(cudaMemsetAsync(h.d_out, 0, (size_t)nb * row_out * 4, h.stream));
    rc = aad_extract(pl, h.d_wav, wav_dtype, max_len, h.d_len, nb, max_len, h.d_out, row_out, t_alloc,
                     h.d_nf, h.d_st, h.d_ws, h.ws_bytes, h.stream);
    if (rc != AAD_OK) return drain(rc);
    if (out_stride_b == row_out)
      HOST_TRY(cudaMemcpyAsync(out_host + (size_t)b0 * out_stride_b, h.d_out, (size_t)nb * row_out * 4,
                               cudaMemcpyDeviceToHost, h.stream));
    else
      HOST_TRY(cudaMemcpy2DAsync(out_host + (size_t)b0 * out_stride_b, (size_t)out_stride_b * 4, h.d_out,
                                 (size_t)row_out * 4, (size_t)row_out * 4, nb, cudaMemcpyDeviceToHost, h.stream));
    HOST_TRY(cudaMemcpyAsync(st_nf + b0, h.d_nf, (size_t)nb * 4, cudaMemcpyDeviceToHost, h.stream));
    HOST_TRY(cudaMemcpyAsync(st_st + b0, h.d_st, (size_t)nb * 4, cudaMemcpyDeviceToHost, h.stream));
  }
#undef HOST_TRY
  for (auto& h : pl->hb) CUDA_TRY(cudaStreamSynchronize(h.stream));
  std::memcpy(n_frames_host, st_nf, (size_t)B * sizeof(int32_t));
  std::memcpy(status_host, st_st, (size_t)B * sizeof(int32_t));
  return AAD_OK;
}

}  // extern "C"
