// CNN-BiLSTM consumer of the features: inference of the reference's
// `AudioDeepfakeDetector` (cnn_bilstm_hybrid.py:20-68, eval mode) as three hand-written sm_100a kernels that
// read the front-end's output where it lies in HBM ((B, F, 63) float32, CT layout), replacing
//   x.permute -> Conv1d(63 -> 64, k 3, pad 1) -> BatchNorm1d -> ReLU -> MaxPool1d(2)   (:24-30, :56-57)
//   -> BiLSTM(64 -> 2 x 32) (:33-39, :60) -> attention / LayerNorm(1) weighting -> max over time (:61-66)
//   -> Linear(64, 64) -> ReLU -> Linear(64, 1) -> Sigmoid                                (:46-52, :67)
// The model is tiny (0.62 MFLOP per 2-second chunk, 160 KB of weights) and the batch is large, so every
// kernel runs with lane = sample: activations are kept feature-major ([feature][sample]) so that a warp
// reads 32 samples with one coalesced load, and the weights sit in shared memory where all lanes read the
// same address (broadcast LDS.128) and feed packed FP32 FMAs (FFMA2).
//   k_det_transpose : feats [B][F][63] -> X0 [F][63][Bp]                        (HBM-bound, 8 bytes / element)
//   k_det_conv      : conv + folded BatchNorm + ReLU at the 2 (F/2) positions the pool keeps -> H0 [p][64][Bp]
//   k_det_lstm      : pool (max of two positions) -> both LSTM directions, 6 steps each, h in registers,
//                     c in local memory -> LayerNorm(1) weighting, max over time -> MLP -> sigmoid
// LayerNorm over a dimension of size 1 returns its bias for every finite input ((s - s) * rsqrt(0 + eps) * w
// + b), so the reference's attention projection and softmax cannot influence the output ("the
// LayerNorm(1) degeneracy"); the kernel applies `lstm_out * ln_bias` exactly as the reference's arithmetic
// does and does not compute the projection (NaN propagation through it is not reproduced).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/aad.h"

namespace aadd {

constexpr int T_IN = 63;    // conv input channels = time frames of a 2-second chunk
constexpr int C1 = 64;      // conv output channels = LSTM input size
constexpr int HID = 32;     // LSTM hidden size per direction
constexpr int DENSE = 64;   // classifier width
constexpr int KCONV = 3 * T_IN;

// ---------------------------------------------------------------- transpose to feature-major
__global__ void __launch_bounds__(256) k_det_transpose(const float* __restrict__ feats, long long stride_b, int stride_f,
                                                       int B, int Bp, float* __restrict__ X0) {
  __shared__ float tile[32][33];
  const int q = blockIdx.z, b0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int b = b0 + r, ci = c0 + tx;
    tile[r][tx] = (b < B && ci < T_IN) ? __ldg(feats + (long long)b * stride_b + (long long)q * stride_f + ci) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int ci = c0 + r, b = b0 + tx;
    if (ci < T_IN && b < Bp) X0[((long long)q * T_IN + ci) * Bp + b] = tile[tx][r];
  }
}

// ---------------------------------------------------------------- conv + BN + ReLU, lane = sample
// smem: Wt[KCONV][64] (k = tap * 63 + ci), scale[64], shift[64] (BatchNorm folded around the conv bias)
// A CTA = 32 samples x 4 warps; warp w = (position-pair group w >> 1, channel half w & 1): it computes TWO adjacent
// positions for 32 of the 64 channels, so that every broadcast LDS.128 of weights feeds four FFMA2 instead of two
// (both compute kernels are co-bound by the shared-memory pipe and the FMA pipe) and the batch still spreads over
// many warps per SM.  Pairs (2i, 2i + 1) are exactly what MaxPool1d(2) combines later.
constexpr int CONV_WARPS = 4;
__global__ void __launch_bounds__(32 * CONV_WARPS) k_det_conv(const float* __restrict__ X0, int F, int Bp, int n_pos,
                                                             const float* __restrict__ Wt, const float* __restrict__ scale_shift,
                                                             float* __restrict__ H0) {
  extern __shared__ __align__(16) float sm[];
  float* sW = sm;
  float* sSS = sm + KCONV * C1;
  for (int i = threadIdx.x; i < KCONV * C1 / 4; i += blockDim.x)
    reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(Wt) + i);
  for (int i = threadIdx.x; i < 2 * C1; i += blockDim.x) sSS[i] = __ldg(scale_shift + i);
  __syncthreads();
  const int b = blockIdx.x * 32 + (threadIdx.x & 31);  // Bp is a multiple of 32
  const int warp = threadIdx.x >> 5, half = warp & 1, c0 = half * (C1 / 2);
  for (int p = 2 * (warp >> 1); p < n_pos; p += CONV_WARPS) {   // n_pos is even: positions p, p + 1
    float2 acc0[C1 / 4], acc1[C1 / 4];
#pragma unroll
    for (int i = 0; i < C1 / 4; ++i) acc0[i] = acc1[i] = make_float2(0.f, 0.f);
    for (int tap = 0; tap < 3; ++tap) {
      const int q0 = p + tap - 1, q1 = q0 + 1;           // input rows of the two positions for this tap
      const bool in0 = q0 >= 0 && q0 < F, in1 = q1 < F;  // zero padding (q1 >= 0 always)
      const float* xp0 = X0 + (long long)(in0 ? q0 : 0) * T_IN * Bp + b;
      const float* xp1 = X0 + (long long)(in1 ? q1 : 0) * T_IN * Bp + b;
      const float4* wp = reinterpret_cast<const float4*>(sW + tap * T_IN * C1 + c0);
#pragma unroll 1
      for (int cb = 0; cb < T_IN; cb += 9) {             // the inputs come from L2: 18 loads in flight per lane
        float xa[9], xb[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          xa[i] = in0 ? __ldg(xp0 + (long long)(cb + i) * Bp) : 0.f;
          xb[i] = in1 ? __ldg(xp1 + (long long)(cb + i) * Bp) : 0.f;
        }
#pragma unroll
        for (int ci = 0; ci < 9; ++ci) {
          const float2 x0 = make_float2(xa[ci], xa[ci]), x1 = make_float2(xb[ci], xb[ci]);
#pragma unroll
          for (int i = 0; i < C1 / 8; ++i) {
            const float4 w = wp[(cb + ci) * (C1 / 4) + i];
            acc0[2 * i] = __ffma2_rn(x0, make_float2(w.x, w.y), acc0[2 * i]);
            acc0[2 * i + 1] = __ffma2_rn(x0, make_float2(w.z, w.w), acc0[2 * i + 1]);
            acc1[2 * i] = __ffma2_rn(x1, make_float2(w.x, w.y), acc1[2 * i]);
            acc1[2 * i + 1] = __ffma2_rn(x1, make_float2(w.z, w.w), acc1[2 * i + 1]);
          }
        }
      }
    }
    float* hp0 = H0 + ((long long)p * C1 + c0) * Bp + b;
    float* hp1 = hp0 + (long long)C1 * Bp;
#pragma unroll
    for (int i = 0; i < C1 / 4; ++i) {
      const float s0 = sSS[c0 + 2 * i], s1 = sSS[c0 + 2 * i + 1], t0 = sSS[C1 + c0 + 2 * i], t1 = sSS[C1 + c0 + 2 * i + 1];
      hp0[(long long)(2 * i) * Bp] = fmaxf(fmaf(acc0[i].x, s0, t0), 0.f);
      hp0[(long long)(2 * i + 1) * Bp] = fmaxf(fmaf(acc0[i].y, s1, t1), 0.f);
      hp1[(long long)(2 * i) * Bp] = fmaxf(fmaf(acc1[i].x, s0, t0), 0.f);
      hp1[(long long)(2 * i + 1) * Bp] = fmaxf(fmaf(acc1[i].y, s1, t1), 0.f);
    }
  }
}

// ---------------------------------------------------------------- BiLSTM + weighting + max + MLP, lane = sample
// smem: Wq[2 dirs][96 inputs][32 units] float4 {i, f, g, o} rows of [W_ih | W_hh], bq[2][32] float4 (b_ih + b_hh),
//       W1t[64 in][64 out], b1[64], w2[64], misc[2] = {b2, ln_bias}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// A tile of 32 samples is run by TWO warps, one per LSTM direction (the batch is small: this doubles the warps per
// SM and halves the critical path); they meet once, at a named barrier, to hand the backward half of the pooled
// vector to the forward warp, which runs the classifier.  sPool: [pairs per CTA][32 units][32 lanes].
constexpr int LSTM_THREADS = 512;
__global__ void __launch_bounds__(LSTM_THREADS) k_det_lstm(const float* __restrict__ H0, int P, int B, int Bp,
                                                          const float* __restrict__ blob, int blob_floats,
                                                          float* __restrict__ scores) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < blob_floats / 4; i += blockDim.x)
    reinterpret_cast<float4*>(sm)[i] = __ldg(reinterpret_cast<const float4*>(blob) + i);
  __syncthreads();
  const float4* sWq = reinterpret_cast<const float4*>(sm);                 // [2][96][32]
  const float4* sBq = sWq + 2 * 96 * HID;                                  // [2][32]
  const float* sW1 = reinterpret_cast<const float*>(sBq + 2 * HID);        // [64][64] (in-major)
  const float* sB1 = sW1 + 2 * HID * DENSE;
  const float* sW2 = sB1 + DENSE;
  const float b2 = sW2[DENSE], ln_b = sW2[DENSE + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pair = warp >> 1, dir = warp & 1, pairs_per_cta = LSTM_THREADS / 64;
  float* sPool = sm + blob_floats + pair * HID * 32;

  // tile -> (CTA, warp pair) with the CTA index fastest: a small batch spreads over all SMs before it stacks up
  for (int tile = pair * gridDim.x + blockIdx.x; tile * 32 < B; tile += gridDim.x * pairs_per_cta) {
    const int b = tile * 32 + lane;  // < Bp
    float v[C1 + HID];               // [x_t | h_{t-1}], static indices only: registers
    float c[HID], hn[HID], pm[HID];  // dynamic index: local memory
#pragma unroll
    for (int j = 0; j < HID; ++j) v[C1 + j] = 0.f;
    for (int u = 0; u < HID; ++u) {
      c[u] = 0.f;
      pm[u] = -INFINITY;             // running max over time of lstm_out * ln_bias
    }
    const float4* Wd = sWq + dir * 96 * HID;
    const float4* Bd = sBq + dir * HID;
    for (int s = 0; s < P; ++s) {
      const int t = dir ? P - 1 - s : s;
      const float* h0 = H0 + (long long)(2 * t) * C1 * Bp + b;
#pragma unroll
      for (int j = 0; j < C1; ++j)   // MaxPool1d(2) of the conv output (already BN + ReLU)
        v[j] = fmaxf(__ldg(h0 + (long long)j * Bp), __ldg(h0 + (long long)(C1 + j) * Bp));
#pragma unroll 1
      for (int u = 0; u < HID; ++u) {
        const float4 bq = Bd[u];
        float2 g01 = make_float2(bq.x, bq.y), g23 = make_float2(bq.z, bq.w);
        float2 h01 = make_float2(0.f, 0.f), h23 = make_float2(0.f, 0.f);  // second chain (ILP)
#pragma unroll
        for (int j = 0; j < C1 + HID; j += 2) {
          const float4 w0 = Wd[j * HID + u], w1 = Wd[(j + 1) * HID + u];
          const float2 x0 = make_float2(v[j], v[j]), x1 = make_float2(v[j + 1], v[j + 1]);
          g01 = __ffma2_rn(x0, make_float2(w0.x, w0.y), g01);
          g23 = __ffma2_rn(x0, make_float2(w0.z, w0.w), g23);
          h01 = __ffma2_rn(x1, make_float2(w1.x, w1.y), h01);
          h23 = __ffma2_rn(x1, make_float2(w1.z, w1.w), h23);
        }
        const float gi = sigmoidf_(g01.x + h01.x), gf = sigmoidf_(g01.y + h01.y);
        const float gg = tanhf(g23.x + h23.x), go = sigmoidf_(g23.y + h23.y);
        const float cn = fmaf(gf, c[u], gi * gg);
        c[u] = cn;
        const float h = go * tanhf(cn);
        hn[u] = h;
        pm[u] = fmaxf(pm[u], h * ln_b);
      }
#pragma unroll
      for (int j = 0; j < HID; ++j) v[C1 + j] = hn[j];
    }
    // the backward warp hands its half of the pooled vector over; the forward warp runs the classifier
    if (dir == 1)
      for (int u = 0; u < HID; ++u) sPool[u * 32 + lane] = pm[u];
    asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
    if (dir == 0) {
      float x[2 * HID];
#pragma unroll
      for (int j = 0; j < HID; ++j) {
        x[j] = pm[j];
        x[HID + j] = sPool[j * 32 + lane];
      }
      float z = b2;
#pragma unroll 1
      for (int n = 0; n < DENSE; n += 4) {
        float4 a = *reinterpret_cast<const float4*>(sB1 + n);
#pragma unroll
        for (int j = 0; j < 2 * HID; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(sW1 + j * DENSE + n);
          a.x = fmaf(x[j], w.x, a.x); a.y = fmaf(x[j], w.y, a.y); a.z = fmaf(x[j], w.z, a.z); a.w = fmaf(x[j], w.w, a.w);
        }
        const float4 w2 = *reinterpret_cast<const float4*>(sW2 + n);
        z = fmaf(fmaxf(a.x, 0.f), w2.x, z); z = fmaf(fmaxf(a.y, 0.f), w2.y, z);
        z = fmaf(fmaxf(a.z, 0.f), w2.z, z); z = fmaf(fmaxf(a.w, 0.f), w2.w, z);
      }
      if (b < B) scores[b] = sigmoidf_(z);
    }
    asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");  // sPool is free for the pair's next tile
  }
}

}  // namespace aadd

// ================================================================== C ABI
struct aad_detector {
  int device = 0, F = 0, P = 0, sm_count = 0;
  float* d_wt = nullptr;     // conv weights [189][64]
  float* d_ss = nullptr;     // BN scale / shift [2][64]
  float* d_blob = nullptr;   // LSTM + classifier block, copied to shared memory by k_det_lstm
  int blob_floats = 0;
};

using namespace aadd;

static size_t det_conv_smem() { return (size_t)(KCONV * C1 + 2 * C1) * 4; }
static size_t det_lstm_smem(int blob_floats) { return ((size_t)blob_floats + (LSTM_THREADS / 64) * HID * 32) * 4; }

static void det_layout(const aad_detector* d, int B, int* Bp, size_t* off_x0, size_t* off_h0, size_t* total) {
  *Bp = (B + 31) / 32 * 32;
  size_t o = 0;
  *off_x0 = o; o += (size_t)d->F * T_IN * *Bp * 4; o = (o + 255) & ~(size_t)255;
  *off_h0 = o; o += (size_t)2 * d->P * C1 * *Bp * 4; o = (o + 255) & ~(size_t)255;
  *total = o;
}

namespace {
// select the object's device for the call and restore the caller's current device afterwards
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
}  // namespace

extern "C" {

int aad_detector_create(const aad_detector_weights* w, int device, aad_detector** out) {
  if (!w || !out || w->struct_size != (int32_t)sizeof(aad_detector_weights)) return AAD_ERR_INVALID_ARG;
  if (w->feature_dim < 2 || w->feature_dim > 4096) return AAD_ERR_INVALID_ARG;
  const float* need[] = {w->conv_w, w->conv_b, w->bn_w, w->bn_b, w->bn_mean, w->bn_var, w->w_ih, w->w_hh, w->b_ih, w->b_hh,
                         w->w_ih_r, w->w_hh_r, w->b_ih_r, w->b_hh_r, w->ln_b, w->fc1_w, w->fc1_b, w->fc2_w, w->fc2_b};
  for (const float* p : need)
    if (!p) return AAD_ERR_INVALID_ARG;
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return AAD_ERR_CUDA;
  aad_detector* d = new aad_detector;
  d->device = device;
  d->F = w->feature_dim;
  d->P = w->feature_dim / 2;
  cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, device);
  // conv weights [co][ci][tap] -> Wt[tap * 63 + ci][co]; BatchNorm (eval) folded: y = conv * s + ((bias - mean) * s + beta)
  std::vector<float> wt((size_t)KCONV * C1), ss(2 * C1);
  for (int co = 0; co < C1; ++co) {
    for (int ci = 0; ci < T_IN; ++ci)
      for (int tap = 0; tap < 3; ++tap) wt[(size_t)(tap * T_IN + ci) * C1 + co] = w->conv_w[((size_t)co * T_IN + ci) * 3 + tap];
    const double s = (double)w->bn_w[co] / std::sqrt((double)w->bn_var[co] + (double)w->bn_eps);
    ss[co] = (float)s;
    ss[C1 + co] = (float)(((double)w->conv_b[co] - (double)w->bn_mean[co]) * s + (double)w->bn_b[co]);
  }
  // LSTM block: Wq[dir][j][u] = {i, f, g, o} rows (PyTorch gate order) of [W_ih | W_hh], biases summed
  std::vector<float> blob;
  blob.resize((size_t)2 * 96 * HID * 4 + 2 * HID * 4 + (size_t)2 * HID * DENSE + DENSE + DENSE + 4, 0.f);
  size_t o = 0;
  for (int dir = 0; dir < 2; ++dir) {
    const float* wih = dir ? w->w_ih_r : w->w_ih;
    const float* whh = dir ? w->w_hh_r : w->w_hh;
    for (int j = 0; j < C1 + HID; ++j)
      for (int u = 0; u < HID; ++u)
        for (int g = 0; g < 4; ++g)
          blob[o++] = j < C1 ? wih[(size_t)(g * HID + u) * C1 + j] : whh[(size_t)(g * HID + u) * HID + (j - C1)];
  }
  for (int dir = 0; dir < 2; ++dir) {
    const float* bi = dir ? w->b_ih_r : w->b_ih;
    const float* bh = dir ? w->b_hh_r : w->b_hh;
    for (int u = 0; u < HID; ++u)
      for (int g = 0; g < 4; ++g) blob[o++] = bi[g * HID + u] + bh[g * HID + u];
  }
  for (int j = 0; j < 2 * HID; ++j)
    for (int n = 0; n < DENSE; ++n) blob[o++] = w->fc1_w[(size_t)n * 2 * HID + j];
  for (int n = 0; n < DENSE; ++n) blob[o++] = w->fc1_b[n];
  for (int n = 0; n < DENSE; ++n) blob[o++] = w->fc2_w[n];
  blob[o++] = w->fc2_b[0];
  blob[o++] = w->ln_b[0];
  o += 2;
  d->blob_floats = (int)blob.size();
  cudaError_t e = cudaMalloc(&d->d_wt, wt.size() * 4);
  if (e == cudaSuccess) e = cudaMalloc(&d->d_ss, ss.size() * 4);
  if (e == cudaSuccess) e = cudaMalloc(&d->d_blob, blob.size() * 4);
  if (e == cudaSuccess) e = cudaMemcpy(d->d_wt, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d->d_ss, ss.data(), ss.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d->d_blob, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)k_det_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_conv_smem());
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)k_det_lstm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_lstm_smem(d->blob_floats));
  if (e != cudaSuccess) {
    aad_detector_destroy(d);
    return AAD_ERR_CUDA;
  }
  *out = d;
  return AAD_OK;
}

int aad_detector_destroy(aad_detector* d) {
  if (!d) return AAD_OK;
  DeviceGuard guard(d->device);
  cudaFree(d->d_wt);
  cudaFree(d->d_ss);
  cudaFree(d->d_blob);
  delete d;
  return AAD_OK;
}

int aad_detector_query(const aad_detector* d, int B, size_t* workspace_bytes) {
  if (!d || B <= 0 || !workspace_bytes) return AAD_ERR_INVALID_ARG;
  int Bp;
  size_t a, b;
  det_layout(d, B, &Bp, &a, &b, workspace_bytes);
  return AAD_OK;
}

int aad_detector_forward(const aad_detector* d, const float* feats, int64_t stride_b, int32_t stride_f, int B,
                         float* scores, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!d || !feats || !scores || !workspace || B <= 0 || stride_f < T_IN) return AAD_ERR_INVALID_ARG;
  if (stride_b == 0) stride_b = (int64_t)d->F * stride_f;
  int Bp;
  size_t off_x0, off_h0, need;
  det_layout(d, B, &Bp, &off_x0, &off_h0, &need);
  if (workspace_bytes < need) return AAD_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  float* X0 = (float*)((char*)workspace + off_x0);
  float* H0 = (float*)((char*)workspace + off_h0);
  (void)cudaGetLastError();
  k_det_transpose<<<dim3(Bp / 32, 2, d->F), 256, 0, stream>>>(feats, stride_b, stride_f, B, Bp, X0);
  k_det_conv<<<Bp / 32, 32 * CONV_WARPS, det_conv_smem(), stream>>>(X0, d->F, Bp, 2 * d->P, d->d_wt, d->d_ss, H0);
  const int tiles = (B + 31) / 32, per_cta = LSTM_THREADS / 64;
  const int grid = std::max(1, std::min(d->sm_count, tiles));
  (void)per_cta;
  k_det_lstm<<<grid, LSTM_THREADS, det_lstm_smem(d->blob_floats), stream>>>(H0, d->P, B, Bp, d->d_blob, d->blob_floats, scores);
  return cudaGetLastError() == cudaSuccess ? AAD_OK : AAD_ERR_CUDA;
}

}  // extern "C"
