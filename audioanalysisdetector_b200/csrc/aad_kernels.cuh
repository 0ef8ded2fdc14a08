// sm_100a kernels of the spectral front-end.
//
//   k_prepare   : per-utterance frame counts, status, exclusive scan of frames, max init
//   k_stft_fb   : fused framing + window + real FFT + |X|^2 + banded filterbank + log
//                 (persistent; one warp-iteration = 32/L frames, FFT entirely in registers,
//                 packed FP32: FFMA2 / FADD2 / FMUL2 on float2 = one complex number; window and
//                 twiddle tables in tensor memory, read with tcgen05.ld)
//   k_cepstra   : dB reference/floor + DCT-II (3xTF32 mma.sync) + delta/delta-delta stencil + layout
//   k_db_finalize, k_znorm, k_time_mean, k_delta : small epilogues
//
// Replaces (reference call chain): librosa.stft / np.abs()**2 / filters.mel einsum /
// power_to_db / scipy dct inside librosa.feature.{melspectrogram,mfcc}
// (ASV_dl_func.py:416,533-534) and spafe pre_emphasis / framing / windowing / fft /
// linear filterbank / log / dct inside spafe.features.lfcc.lfcc (ASV_dl_func.py:435).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "aad_fft.cuh"

namespace aad {

enum InMode { IN_F32 = 0, IN_F32_Q16 = 1, IN_I16 = 2 };

// Development only: -DAAD_ABLATE=<mask> removes parts of k_stft_fb (results become wrong) so that the
// marginal cost of each part can be timed (tools/ablate_k1.sh).  0 in every shipped build.
#ifndef AAD_ABLATE
#define AAD_ABLATE 0
#endif
constexpr int ABL = AAD_ABLATE;
// mask bits: 4 transposes, 8 filterbank
// phase, 16 split exchange + twp table, 32 power stores, 64 butterflies, 128 global sample loads, 256 log
// in the filterbank emit
//
// The shared-memory / L1 data pipe is the most loaded unit of k_stft_fb, the FMA pipe has headroom, so
// table look-ups are kept off it: the lane x register tables (window, twiddles) live in tensor
// memory; on variants without TMEM room for the split twiddle, AAD_TWPGEN forms it as
// W_N^k = W_N^(j + L q) [per-lane register] * W_N^(32 s) [immediate] instead of loading it.
#ifndef AAD_TWPGEN
#define AAD_TWPGEN 1
#endif
#ifndef AAD_ROW_SKEW
#define AAD_ROW_SKEW 1
#endif
#ifndef AAD_TWFOLD
#define AAD_TWFOLD 1
#endif

// ---------------------------------------------------------------------------
// ordered-int encoding of floats (monotone), for atomicMax / redux on floats
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ int enc_ordered(float f) {
#ifdef __CUDA_ARCH__
  int i = __float_as_int(f);
#else
  int i;
  memcpy(&i, &f, 4);
#endif
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float dec_ordered(int i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}
#define AAD_ENC_NEG_INF ((int)0x807fffff)

// ---------------------------------------------------------------------------
// K0: prepare
// ---------------------------------------------------------------------------
struct PrepArgs {
  const int32_t* lengths;
  int B;
  long long max_len;
  int hop, win_len, center, n_delta, delta_width, t_alloc;
  int32_t* n_frames;   // out (user)
  int32_t* status;     // out (user)
  int32_t* len_c;      // ws: clamped lengths
  int32_t* nf_eff;     // ws: frames actually computed (0 when status != 0)
  int32_t* frame_off;  // ws: [B+1] exclusive scan of nf_eff
  int32_t* utt_max;    // ws: ordered-int encoded running max, init -inf
  int32_t* utt_max2;   // the same for the second filter bank of a paired call (or null)
  int4* tile_rec;      // ws: [max_tiles][3] what k_stft_fb needs to place the frames of a tile (see TileRec)
  int tile, max_tiles;
  const long long* row_off;  // [B] element offsets of the utterances (or null: b * wav_stride)
  long long wav_stride;
  double* zn_stats;    // ws: [B][2] z-norm accumulators (null when unused)
};

#ifndef AAD_STFT_ONLY  // the translation units that only instantiate k_stft_fb (aad_stft_inst.cu) skip the other kernels
// frames of utterance b and its status (the geometry of librosa.stft(center=True) / spafe's framing)
__device__ __forceinline__ int prep_geometry(const PrepArgs& a, int b, int& st, int& len_c) {
  long long len = __ldg(a.lengths + b);
  if (len > a.max_len) len = a.max_len;
  st = 0;
  int T = 0;
  if (len <= 0) {
    st = 1;
    len = 0;
  } else {
    T = a.center ? (int)(1 + len / a.hop) : (len >= a.win_len ? (int)((len - a.win_len) / a.hop + 1) : 0);
    if (T == 0) st = 2;
    else if (a.n_delta > 0 && T < a.delta_width) st = 3;
    else if (T > a.t_alloc) st = 4;
  }
  len_c = (int)len;
  return T;
}

// One CTA per kPrepBlock utterances.  The exclusive scan of the frame counts needs the total of all earlier
// utterances: every CTA recomputes it from the lengths (base / kPrepBlock reads per thread: a few microseconds even for
// 10^5 utterances) instead of waiting for its predecessors, so the grid has no inter-CTA dependency and the tile table
// of a large batch (51 000 tiles for 4096 x 4 s at hop 160) is written by many SMs.
constexpr int kPrepBlock = 256;
__global__ void __launch_bounds__(kPrepBlock) k_prepare(PrepArgs a) {
  __shared__ int warp_sums[kPrepBlock / 32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // programmatic dependent launch: k_stft_fb may start its set-up now; it waits for this grid before it reads anything
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int base = blockIdx.x * kPrepBlock;
  {  // frames in front of this CTA's utterances
    int part = 0;
    for (int i = tid; i < base; i += kPrepBlock) {
      int st, lc;
      const int T = prep_geometry(a, i, st, lc);
      part += st ? 0 : T;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) warp_sums[warp] = part;
    __syncthreads();
    if (tid == 0) {
      int c = 0;
#pragma unroll
      for (int w = 0; w < kPrepBlock / 32; ++w) c += warp_sums[w];
      carry_s = c;
    }
    __syncthreads();
  }
  const int b = base + tid;
  int nf = 0;
  if (b < a.B) {
    int st, lc;
    const int T = prep_geometry(a, b, st, lc);
    nf = st ? 0 : T;
    a.n_frames[b] = T;
    a.status[b] = st;
    a.len_c[b] = lc;
    a.nf_eff[b] = nf;
    a.utt_max[b] = AAD_ENC_NEG_INF;
    if (a.utt_max2) a.utt_max2[b] = AAD_ENC_NEG_INF;
    if (a.zn_stats) {
      a.zn_stats[2 * b] = 0.0;
      a.zn_stats[2 * b + 1] = 0.0;
    }
  }
  // block exclusive scan of nf
  int x = nf;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  const int carry = carry_s;
  __syncthreads();  // warp_sums is reused
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  int before = 0;
#pragma unroll
  for (int w = 0; w < kPrepBlock / 32; ++w) before += w < warp ? warp_sums[w] : 0;
  const int excl = carry + before + (x - nf);
  if (b < a.B) a.frame_off[b] = excl;
  if (b == a.B - 1) a.frame_off[a.B] = excl + nf;
  // tile records for k_stft_fb: tile tl (first frame tl*tile) starts inside utterance b iff
  // frame_off[b] <= tl*tile < frame_off[b+1].  The record carries everything the tile's frames need as long as they
  // lie in utterances b and b + 1 (frame offsets, clamped lengths, row offsets), so that k_stft_fb places a tile with
  // one round of independent loads instead of a chain of four dependent ones.  Each thread writes the few tiles of its
  // own utterance; utterances that own many tiles (long-form audio) are written by the whole warp.
  {
    int4 r0, r1;
    longlong2 r2;
    {
      int st1 = 1, len1 = 0, nf1 = 0, lc0 = 0, st0;
      if (b < a.B) prep_geometry(a, b, st0, lc0);
      if (b + 1 < a.B) {
        const int T1 = prep_geometry(a, b + 1, st1, len1);
        nf1 = st1 ? 0 : T1;
      }
      r0 = make_int4(b, excl, excl + nf, excl + nf + nf1);
      r1 = make_int4(lc0, len1, 0, 0);
      r2.x = b < a.B ? (a.row_off ? __ldg(a.row_off + b) : (long long)b * a.wav_stride) : 0;
      r2.y = b + 1 < a.B ? (a.row_off ? __ldg(a.row_off + b + 1) : (long long)(b + 1) * a.wav_stride) : 0;
    }
    auto put = [&](int tl, const int4& q0, const int4& q1, const longlong2& q2) {
      if (tl < a.max_tiles) {
        a.tile_rec[3 * tl] = q0;
        a.tile_rec[3 * tl + 1] = q1;
        *reinterpret_cast<longlong2*>(a.tile_rec + 3 * tl + 2) = q2;
      }
    };
    const int t_first = (excl + a.tile - 1) / a.tile, t_end = nf ? (excl + nf + a.tile - 1) / a.tile : t_first;
    const bool big = t_end - t_first > 32;
    if (!big)
      for (int tl = t_first; tl < t_end; ++tl) put(tl, r0, r1, r2);
    unsigned todo = __ballot_sync(0xffffffffu, big);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int f0 = __shfl_sync(0xffffffffu, t_first, src), f1 = __shfl_sync(0xffffffffu, t_end, src);
      int4 q0, q1;
      longlong2 q2;
      q0.x = __shfl_sync(0xffffffffu, r0.x, src); q0.y = __shfl_sync(0xffffffffu, r0.y, src);
      q0.z = __shfl_sync(0xffffffffu, r0.z, src); q0.w = __shfl_sync(0xffffffffu, r0.w, src);
      q1.x = __shfl_sync(0xffffffffu, r1.x, src); q1.y = __shfl_sync(0xffffffffu, r1.y, src);
      q1.z = 0; q1.w = 0;
      q2.x = __shfl_sync(0xffffffffu, r2.x, src); q2.y = __shfl_sync(0xffffffffu, r2.y, src);
      for (int tl = f0 + lane; tl < f1; tl += 32) put(tl, q0, q1, q2);
    }
  }
}

#endif  // AAD_STFT_ONLY

// ---------------------------------------------------------------------------
// Tensor memory (TMEM) as a per-lane table store.  The window and the pass-1 twiddles of k_stft_fb
// are lane x register tables read once per frame; from shared memory they cost 126 wavefronts per
// frame on the L1/shared data pipe, the kernel's most loaded unit.  TMEM (256 KB per SM, 128 lanes x
// 512 columns x 32 bit) is read through its own datapath (tcgen05.ld, SASS LDTM): with the 32x32b
// shape thread i of warp w reads lane 32*(w%4) + i, i.e. exactly "its own row".
// ---------------------------------------------------------------------------
#ifndef AAD_TWP_TMEM
#define AAD_TWP_TMEM 1
#endif
#ifndef AAD_T64
#define AAD_T64 1
#endif
// dev switch (profiles/r2_experiments.md 12): samples of a tile staged once in shared memory by a 1-D bulk copy (TMA)
// issued during the previous tile's filterbank phase; n_fft 2048, float32 input, no pre-emphasis, dense rows
#ifndef AAD_TMA_STAGE
#define AAD_TMA_STAGE 0
#endif
// columns: [0, 64) window pairs, [64, 128) pass-1 twiddles, [128, 160) split twiddles (only when at
// most two CTAs share the SM's 512 columns: allocations are powers of two)
template <int CTAS>
struct TmemCfg {
  static constexpr bool TWP = AAD_TWP_TMEM && CTAS <= 2;
  static constexpr int COLS = TWP ? 256 : 128;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float2 (&w)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      :: "r"(taddr), "r"(__float_as_uint(w[0].x)), "r"(__float_as_uint(w[0].y)), "r"(__float_as_uint(w[1].x)),
         "r"(__float_as_uint(w[1].y)), "r"(__float_as_uint(w[2].x)), "r"(__float_as_uint(w[2].y)),
         "r"(__float_as_uint(w[3].x)), "r"(__float_as_uint(w[3].y)), "r"(__float_as_uint(w[4].x)),
         "r"(__float_as_uint(w[4].y)), "r"(__float_as_uint(w[5].x)), "r"(__float_as_uint(w[5].y)),
         "r"(__float_as_uint(w[6].x)), "r"(__float_as_uint(w[6].y)), "r"(__float_as_uint(w[7].x)),
         "r"(__float_as_uint(w[7].y))
      : "memory");
}
// 8 float2 of this thread's TMEM row, starting at column (taddr & 0xffff): issue, then wait.  The
// registers are operands of the wait so that no use can be scheduled ahead of it; the wait covers every
// load issued before it, so a second chunk issued right after the wait stays in flight while the first
// one is consumed.
struct TmemChunk {
  uint32_t r[16];
  __device__ __forceinline__ void issue(uint32_t taddr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
  }
  __device__ __forceinline__ void wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
  }
  __device__ __forceinline__ float2 get(int i) const {
    return make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
  }
};

// ---------------------------------------------------------------------------
// K1: fused STFT + power + filterbank + log
// ---------------------------------------------------------------------------
// frames per k_stft_fb tile.  n_fft 512: 64 (two frames per lane in the filterbank phase: the walk over the program --
// headers, loop control, weight look-ups -- is paid once per 64 frames, and there is one barrier pair per 64)
#ifndef AAD_TILE_L8
#define AAD_TILE_L8 32
#endif
#ifndef AAD_TILE_DENSE
#define AAD_TILE_DENSE 64
#endif
__host__ __device__ constexpr int stft_tile(int L, bool dense = false) { return dense ? AAD_TILE_DENSE : (L == 8 ? AAD_TILE_L8 : 32); }
// dense filter banks (k_stft_fb FBM = 1): entries per bundle.  All entries of a bundle read the same bins, so the power
// group is loaded once per round for the whole bundle; with 8 warps a warp owns 2 or 3 of the 20 entries of spafe's 40
// filters.
constexpr int kDenseFbu = 5;

template <int L, int TILE_, bool DENSE = false>
struct StftCfg {
  static constexpr int Q = 32 / L;        // frames per warp-iteration
  static constexpr int M = 32 * L;        // complex FFT length (n_fft / 2)
  static constexpr int N = 2 * M;         // n_fft
  static constexpr int K = M + 1;         // bins
  static constexpr int TILE = TILE_;      // frames per tile
  // filterbank phase: entries processed together (ILP).  Measured (profiles/r1_experiments.md): bundles of 2
  // are best for n_fft 2048 (1 / 2 / 3 / 4: 1.48 / 1.21 / 1.25 / 1.26 ms), bundles of 3 for n_fft 512 with its
  // short segments (2 / 3 / 4: 1.143 / 1.120 / 1.116 ms on 80 mels; LFCC 1.239 / 1.235 / 1.288)
#ifdef AAD_FBU
  static constexpr int FBU = AAD_FBU;
#else
  static constexpr int FBU = L == 8 ? 3 : 2;
#endif
  // power-row stride in floats.  The filterbank phase reads P[frame][4g .. 4g+3] as one LDS.128 per
  // lane: conflict-free iff SP/4 is odd.  Q rows double as the warp's 32x33 transpose scratch
  // (Q*SP >= 1056), and the row holds K bins plus zeroed padding (PAD words).
  // For L = 8 / 16 (4 / 2 frames per warp-iteration) no single stride serves both access patterns: the
  // power stores of a warp-iteration (Q rows x L consecutive bins) need stride = L (mod 32), the
  // filterbank reads stride/4 odd.  There the stride is a multiple of 32 and row r is skewed by
  // skew(r) = L (r mod Q) + 4 ((r / Q) mod (8 / Q)) words: the Q rows of an iteration tile the 32 banks
  // and 8 consecutive rows start in 8 different 16-byte bank groups (measured before: 2-way conflicts on
  // every power store of the n_fft 512 shape, 9 % of its shared-memory wavefronts).
  // (dense filter banks with 64-frame tiles: the skewed rows would not leave room for the weights of two CTAs per SM)
  static constexpr bool SKEW = AAD_ROW_SKEW && (L == 8 || L == 16) && !(DENSE && TILE_ == 64);
  static constexpr int SP = L == 4 ? 132 : (L == 8 ? (SKEW ? 288 : (DENSE ? 268 : 292)) : (L == 16 ? (SKEW ? 544 : 548) : 1060));
  __host__ __device__ static constexpr int skew(int r) {
    return SKEW ? L * (r & (Q - 1)) + 4 * ((r / Q) & (8 / Q - 1)) : 0;
  }
  static constexpr int PAD = 3;
  static constexpr int ITERS = TILE / Q;  // warp-iterations per tile
#ifdef AAD_WARPS_DEV
  static constexpr int WARPS = AAD_WARPS_DEV;
#else
  // dense filter banks keep their 42 KB of weights in shared memory: two 8-warp CTAs per SM instead of four 4-warp ones
  static constexpr int WARPS = DENSE ? 8 : (ITERS >= 8 ? ITERS / 2 : 4);
#endif
  static constexpr int FPL = TILE / 32;    // frames per lane in the filterbank phase (lane handles frames lane + 32 f)
  static constexpr int CTAS = L == 32 ? 1 : (L == 16 || TILE == 64 || DENSE ? 2 : 4);
  // shared memory carve-up, in floats (the filterbank program follows at OFF_PROG)
  // the window / twiddle tables live in tensor memory (or are generated) and take no shared memory then
  static constexpr bool SMEM_TWP = !(TmemCfg<CTAS>::TWP || (AAD_TWPGEN && Q <= 4));
  static constexpr int OFF_P = 0;
  static constexpr int OFF_TWP = TILE * SP;
  static constexpr int OFF_META = OFF_TWP + (SMEM_TWP ? 2 * (M / 2) : 0);  // 2 x {b[TILE], t[TILE]} (double buffered)
  static constexpr int OFF_TMEM = OFF_META + 4 * TILE;        // TMEM base address written by tcgen05.alloc
  static constexpr int OFF_PROG = OFF_TMEM + 4;               // segment headers + tap weights follow
  static constexpr size_t FIXED_BYTES = size_t(OFF_PROG) * 4;
  static_assert(TILE == 32 || TILE == 64, "filterbank phase runs with lane = frame (mod 32)");
  static_assert(SKEW ? SP % 32 == 0 : SP % 8 == 4, "LDS.128 over lane = frame needs 8 rows in 8 different 16-byte bank groups");
  static_assert(Q * SP >= 32 * 33, "power rows must hold the transpose scratch");
  static_assert(SP >= K + PAD + (SKEW ? 28 : 0), "row must hold the bins, the zero padding and the skew");
  static_assert(OFF_PROG % 4 == 0 && OFF_TWP % 4 == 0, "tables must be 16-byte aligned");
};

struct StftArgs {
  const void* wav;
  long long wav_stride;     // elements
  const long long* row_off; // [B] element offset of every utterance inside wav (chunks of decoded files,
                            // may overlap), or null: utterance b starts at b * wav_stride
  const int32_t* len_c;     // [B] clamped lengths
  const int32_t* frame_off; // [B+1]
  int B;
  int hop, s_off;           // first sample of frame t is t*hop - s_off
  int win_off, win_len;     // window support inside the n_fft buffer
  float pre_emph;
  const float* window;      // [N]   0.5 * window (zero outside support)
  const float2* tw1;        // [32*L] exp(-2 pi i b kA / M) at [kA*L + b]
  const float2* twp;        // [M/2]  exp(-2 pi i k / N)
  // filterbank program (copied to smem), two-tap banded form: the bins split into n_filt + 1
  // segments; inside segment s bin k feeds filter s with its rising weight wr[k] and filter s - 1
  // with its falling weight wf[k], so  energy[j] = R[j] + F[j + 1]  with (R, F)[s] the two
  // weighted sums over segment s and every power value is read once.  Each warp owns a list of
  // entries = the segments of its filter range; per entry a header {first bin (multiple of 4) |
  // n_rounds << 16 with the bin as a BYTE offset into the power row, weight BYTE offset}; per round 4
  // bins as two float4 {wr0, wf0, wr1, wf1}, {wr2, wf2, wr3, wf3} (zero padded).  Entries are processed
  // FBU at a time (a bundle: independent accumulators and log chains): the list is padded to a multiple
  // of FBU with empty entries, the rounds are equalised inside each bundle and the bundle's weights are
  // interleaved [round][entry][2 x float4] at the offset of its first entry.
  const int2* filt_hdr;     // [n_hdr]
  const float4* filt_w;     // [n_w4]
  int n_hdr, n_w4;
  const int4* warp_prog;    // [WARPS] {first filter, first entry, n_entries, n_filters}
  const int4* tile_rec;     // [n_tiles][3] from k_prepare: {b0, off[b0], off[b0+1], off[b0+2]}, {len[b0], len[b0+1]}, {row[b0], row[b0+1]}
  int n_filt;
  int log_type;             // 0 dB, 1 ln, 2 cube root (dense filter banks: spafe gfcc)
  int spec_mag;             // dense variant only: 1 = the filter bank is applied to |X| instead of |X|^2
  float amin, eps;
  float* E;                 // log-energies out: E[b*stride_b + f*stride_f + t]
  long long e_stride_b;
  int e_stride_f;
  int32_t* utt_max;         // ordered-int encoded, or null
  int32_t* status;          // for NONFINITE flagging
  // Optional SECOND filter bank over the same power spectra: one STFT, two features (e.g. the 128-mel bank of
  // the MFCC plan and the 64-mel bank of the log-mel plan, the extractor map of ASV_deep_learning.ipynb:152-160).
  // Same program format, built for the same warp count and bundle size; read by the PAIR instantiations only.
  const int2* filt_hdr2;
  const float4* filt_w2;
  int n_hdr2, n_w42;
  const int4* warp_prog2;
  int log_type2;
  float amin2;
  float* E2;
  long long e2_stride_b;
  int e2_stride_f;
  int32_t* utt_max2;
};

template <int MODE>
__device__ __forceinline__ float cvt_sample(float x) {
  if constexpr (MODE == IN_F32_Q16) {
    // (y * 32767).astype(np.int16): float32 multiply, truncate toward zero, keep low 16 bits
    int q = __float2int_rz(x * 32767.0f);
    return (float)(short)q;
  } else {
    return x;
  }
}

// slow path: one sample with all masks (window support, utterance bounds)
template <int MODE, bool PRE>
__device__ __forceinline__ float load_masked(const void* row, int idx, int n, int len, int win_off,
                                             int win_len, float pre) {
  bool ok = (unsigned)(n - win_off) < (unsigned)win_len && (unsigned)idx < (unsigned)len;
  float x = 0.f, xp = 0.f;
  if (ok) {
    if constexpr (MODE == IN_I16) {
      x = (float)__ldg((const short*)row + idx);
      if (PRE && idx >= 1) xp = (float)__ldg((const short*)row + idx - 1);
    } else {
      x = cvt_sample<MODE>(__ldg((const float*)row + idx));
      if (PRE && idx >= 1) xp = cvt_sample<MODE>(__ldg((const float*)row + idx - 1));
    }
  }
  if constexpr (PRE) x = __fmaf_rn(-pre, xp, x);
  return x;
}

__device__ __forceinline__ float2 i16pair_to_float2(unsigned p) {
  // exact int16 -> float via the 2^23 magic number (ALU + FADD instead of I2F)
  float2 m = make_float2(__uint_as_float(0x4B000000u | ((p & 0xffffu) ^ 0x8000u)),
                         __uint_as_float(0x4B000000u | ((p >> 16) ^ 0x8000u)));
  return __fadd2_rn(m, make_float2(-8421376.0f, -8421376.0f));
}

#ifdef AAD_PHASE_TIMING
// dev only: warp-cycles spent in {FFT phase, barrier after it, filterbank phase, barrier after it}
__device__ unsigned long long g_phase_cycles[4];
#define AAD_PHASE_MARK(i)                                                         \
  do {                                                                            \
    long long now__ = clock64();                                                  \
    if (lane == 0) atomicAdd(&g_phase_cycles[i], (unsigned long long)(now__ - tmark)); \
    tmark = now__;                                                                \
  } while (0)
#else
#define AAD_PHASE_MARK(i)
#endif

// ---------------------------------------------------------------------------
// One warp-iteration of the STFT: Q = 32 / L frames (L lanes each) -> their power spectra in shared memory.
// ---------------------------------------------------------------------------
template <int L, int MODE, bool PRE, int TILE>
struct FrameFft {
  using C = StftCfg<L, TILE>;
  static constexpr int Q = C::Q, M = C::M, N = C::N, LOG2L = ilog2(L);
  // n_fft 2048: the inter-pass twiddle is applied AFTER the transpose (the table is symmetric in (b, kA)) and fused
  // with the first butterfly stage of pass 2: 5 packed instructions per butterfly instead of 6
  static constexpr bool TWFOLD = AAD_TWFOLD && L == 32;
  static constexpr bool TWPTM = TmemCfg<C::CTAS>::TWP;
  static constexpr bool TWPGEN = !TWPTM && AAD_TWPGEN && Q <= 4;
  static constexpr bool T64 = AAD_T64 && Q >= 2 && Q <= 4;   // 64-bit transposes in two halves

  int lane, j, g, partner;  // lane = g * L + j: frame-in-iteration g, lane j within the frame's group
  uint32_t tmem_row;        // TMEM address of this thread's row (lane quarter of the warp, column 0 of the tables)
  float2 twp_base[Q];       // W_N^(j + L q) (TWPGEN)
  const float2* sTwp;       // split twiddles in shared memory (only the variants with neither TMEM room nor TWPGEN)

  __device__ __forceinline__ void init(const StftArgs& a, int lane_, uint32_t tmem_row_, const float2* sTwp_) {
    lane = lane_;
    g = lane / L;
    j = lane % L;
    partner = g * L + ((L - j) & (L - 1));
    tmem_row = tmem_row_;
    sTwp = sTwp_;
#pragma unroll
    for (int q = 0; q < Q; ++q) twp_base[q] = TWPGEN ? __ldg(a.twp + j + L * q) : make_float2(0.f, 0.f);
  }

  // Fill one TMEM lane quarter with the per-lane rows (executed by one warp per quarter):
  //   columns [0, 64): window pairs {0.5 w[2(L A + j)], 0.5 w[2(L A + j) + 1]} ordered {w[A], w[A + 16]} (see the window step)
  //   columns [64, 128): pass-1 twiddles W_M^(j kA)
  //   columns [128, 160): split twiddles W_N^(j + L q + 32 s) (TWPTM only)
  __device__ __forceinline__ static void fill_tables(const StftArgs& a, uint32_t tmem_row, int j) {
    const float2* gwin2 = reinterpret_cast<const float2*>(a.window);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float2 w8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int A = 4 * c + (i >> 1) + 16 * (i & 1);
        w8[i] = __ldg(gwin2 + L * A + j);
      }
      tmem_st16(tmem_row + 16 * c, w8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // TWFOLD: {W^(j b), W^(j (b + 16))} side by side, b = 4c .. 4c + 3 (the pairs of pass 2's first stage)
        const int kA = TWFOLD ? 4 * c + (i >> 1) + 16 * (i & 1) : 8 * c + i;
        w8[i] = __ldg(a.tw1 + kA * L + j);
      }
      tmem_st16(tmem_row + 64 + 16 * c, w8);
    }
    if constexpr (TWPTM) {  // entry q*(L/2) + s = W_N^(j + L q + 32 s)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float2 w8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int e = 8 * c + i, q = e / (L / 2), sidx = e % (L / 2);
          w8[i] = __ldg(a.twp + j + L * q + 32 * sidx);
        }
        tmem_st16(tmem_row + 128 + 16 * c, w8);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }

  // load: the samples of this lane's frame into v (bit-reversed register order), with pre-emphasis / quantisation.
  // valid / t / len / row: this lane's frame (the L lanes of a frame agree): frame t of an utterance of len samples
  // whose first sample is at `row`.
  __device__ __forceinline__ void load(const StftArgs& a, bool valid, int t, int len, const char* row,
                                       float2 (&v)[32], const float* stage = nullptr, int stage_a0 = 0) const {
    const int s0 = t * a.hop - a.s_off;
    bool fast = valid && s0 >= (PRE ? 1 : 0) && (s0 + N) <= len;
    if constexpr (AAD_TMA_STAGE && MODE == IN_F32 && !PRE) {
      // interior frame of a staged tile: the samples are in shared memory (a0 is a multiple of 4, s0 is even)
      if (stage != nullptr && __all_sync(0xffffffffu, fast)) {
        const float2* p = reinterpret_cast<const float2*>(stage + (s0 - stage_a0)) + j;
        static_for<0, 32>([&](auto a_) {
          constexpr int A = decltype(a_)::value;
          v[bitrev(A, 5)] = p[L * A];
        });
        return;
      }
    }
    if constexpr (MODE == IN_I16) fast = fast && (((uintptr_t)(row + 2ll * s0)) & 3) == 0;
    else fast = fast && (((uintptr_t)(row + 4ll * s0)) & 7) == 0;

    if (__all_sync(0xffffffffu, fast)) {
      if constexpr (MODE == IN_I16) {
        const unsigned* p = reinterpret_cast<const unsigned*>(row + 2ll * s0) + j;
        unsigned raw[32];
        static_for<0, 32>([&](auto a_) {
          constexpr int A = decltype(a_)::value;
          raw[A] = __ldg(p + L * A);
        });
        if constexpr (PRE) {
          const short* ps = reinterpret_cast<const short*>(row + 2ll * s0) + 2 * j - 1;
          static_for<0, 32>([&](auto a_) {
            constexpr int A = decltype(a_)::value;
            float xp = (float)__ldg(ps + 2 * L * A);
            float2 x = i16pair_to_float2(raw[A]);
            float2 e = make_float2(__fmaf_rn(-a.pre_emph, xp, x.x), __fmaf_rn(-a.pre_emph, x.x, x.y));
            v[bitrev(A, 5)] = e;
          });
        } else {
          static_for<0, 32>([&](auto a_) {
            constexpr int A = decltype(a_)::value;
            v[bitrev(A, 5)] = i16pair_to_float2(raw[A]);
          });
        }
      } else {
        const float2* p = reinterpret_cast<const float2*>(row + 4ll * s0) + j;
        static_for<0, 32>([&](auto a_) {
          constexpr int A = decltype(a_)::value;
          if constexpr (ABL & 128) v[bitrev(A, 5)] = make_float2(a.amin * (lane + A), a.eps * (s0 + A));
          else v[bitrev(A, 5)] = __ldg(p + L * A);
        });
        if constexpr (PRE) {
          const float* ps = reinterpret_cast<const float*>(row + 4ll * s0) + 2 * j - 1;
          static_for<0, 32>([&](auto a_) {
            constexpr int A = decltype(a_)::value;
            float xp = cvt_sample<MODE>(__ldg(ps + 2 * L * A));
            float2 x = v[bitrev(A, 5)];
            x.x = cvt_sample<MODE>(x.x);
            x.y = cvt_sample<MODE>(x.y);
            float2 e = make_float2(__fmaf_rn(-a.pre_emph, xp, x.x), __fmaf_rn(-a.pre_emph, x.x, x.y));
            v[bitrev(A, 5)] = e;
          });
        } else if constexpr (MODE == IN_F32_Q16) {
          static_for<0, 32>([&](auto a_) {
            constexpr int A = decltype(a_)::value;
            const float2 x = v[bitrev(A, 5)];
            v[bitrev(A, 5)] = make_float2(cvt_sample<MODE>(x.x), cvt_sample<MODE>(x.y));
          });
        }
      }
    } else if (!PRE && __all_sync(0xffffffffu, !valid || (((uintptr_t)(row + (MODE == IN_I16 ? 2ll : 4ll) * s0)) & (MODE == IN_I16 ? 3 : 7)) == 0)) {
      // edge frames on aligned rows (centre padding, utterance tail, zero-extended window): pairs
      // fully inside [n_lo, n_hi) keep the vector load, the (at most two) straddling pairs and the
      // pairs outside are handled per sample.  Keeps an edge frame within ~10 % of an interior one,
      // which matters because the 16 warps of a tile meet at a barrier.
      const int n_lo = max(a.win_off, -s0);
      const int n_hi = valid ? min(a.win_off + a.win_len, len - s0) : 0;
      static_for<0, 32>([&](auto a_) {
        constexpr int A = decltype(a_)::value;
        const int n0 = 2 * (L * A + j);
        float2 x;
        if (n0 >= n_lo && n0 + 2 <= n_hi) {
          if constexpr (MODE == IN_I16) {
            x = i16pair_to_float2(__ldg(reinterpret_cast<const unsigned*>(row + 2ll * (s0 + n0))));
          } else {
            x = __ldg(reinterpret_cast<const float2*>(row + 4ll * (s0 + n0)));
            if constexpr (MODE == IN_F32_Q16) x = make_float2(cvt_sample<MODE>(x.x), cvt_sample<MODE>(x.y));
          }
        } else {
          x.x = (n0 >= n_lo && n0 < n_hi) ? load_masked<MODE, false>(row, s0 + n0, n0, len, a.win_off, a.win_len, 0.f) : 0.f;
          x.y = (n0 + 1 >= n_lo && n0 + 1 < n_hi) ? load_masked<MODE, false>(row, s0 + n0 + 1, n0 + 1, len, a.win_off, a.win_len, 0.f) : 0.f;
        }
        v[bitrev(A, 5)] = x;
      });
    } else {
      // pre-emphasis or unaligned rows on an edge frame: every sample with all masks
      static_for<0, 32>([&](auto a_) {
        constexpr int A = decltype(a_)::value;
        const int n0 = 2 * (L * A + j);
        float x0 = load_masked<MODE, PRE>(row, s0 + n0, n0, len, a.win_off, a.win_len, a.pre_emph);
        float x1 = load_masked<MODE, PRE>(row, s0 + n0 + 1, n0 + 1, len, a.win_off, a.win_len, a.pre_emph);
        v[bitrev(A, 5)] = make_float2(x0, x1);
      });
    }

  }

  // request the samples of warp-iteration `it` of the tile whose meta is (mB, mT); false: no valid frame in it
  __device__ __forceinline__ bool load_iter(const StftArgs& a, const int* mB, const int* mT, int it, float2 (&v)[32],
                                            const float* stage = nullptr, int stage_a0 = 0) const {
    const int fi = it * Q + g;
    const int b = mB[fi], t = mT[fi];
    const bool valid = b >= 0;
    if (!__any_sync(0xffffffffu, valid)) return false;
    const int len = valid ? __ldg(a.len_c + b) : 0;
    const char* row = static_cast<const char*>(a.wav) +
                      (valid ? (a.row_off ? __ldg(a.row_off + b) : (long long)b * a.wav_stride) * (MODE == IN_I16 ? 2 : 4) : 0);
    load(a, valid, t, len, row, v, stage, stage_a0);
    return true;
  }

  // transform: window, FFT, real-input split, power.  scr: warp-private scratch of >= 32 * 33 floats (may alias prow:
  // the transposes are over before the powers are written).  prow: power row of this lane's frame, M + 1 + NPAD
  // floats are written.
  template <int NPAD, bool MAGOPT = false>
  __device__ __forceinline__ void transform(const StftArgs& a, float2 (&v)[32], float* scr, float* prow) const {
    // |X|^2, or |X| when the (dense filter bank) plan asks for the magnitude spectrum
    const bool mag = MAGOPT && a.spec_mag != 0;
    auto spec = [&](float re, float im) {
      float pw = __fmaf_rn(re, re, im * im);
      if constexpr (MAGOPT) {
        if (mag) asm("sqrt.approx.ftz.f32 %0, %0;" : "+f"(pw));
      }
      return pw;
    };
    // window: 0.5*w (zero outside its support) from this lane's TMEM row, next chunk in flight
    // window folded into the first butterfly stage of pass 1: samples A and A + 16 meet in stage 1
    // (registers bitrev(A) = 2m and 2m + 1), so a' = xa*wa + xb*wb, b' = xa*wa - xb*wb is one FMUL2
    // and two FFMA2 instead of two FMUL2 and two FADD2.  Chunk c holds {w[A], w[A + 16]}, A = 4c .. 4c + 3.
    {
      TmemChunk wc[2];
      wc[0].issue(tmem_row);
      static_for<0, 4>([&](auto c_) {
        constexpr int CH = decltype(c_)::value;
        wc[CH & 1].wait();
        if constexpr (CH < 3) wc[(CH + 1) & 1].issue(tmem_row + 16 * (CH + 1));
        static_for<0, 4>([&](auto i_) {
          constexpr int I = decltype(i_)::value;
          constexpr int RA = bitrev(4 * CH + I, 5);            // even register; RA + 1 = bitrev(A + 16)
          const float2 t = pk_mul(v[RA], wc[CH & 1].get(2 * I));
          const float2 xb = v[RA + 1], wb = wc[CH & 1].get(2 * I + 1);
          v[RA] = pk_fma(xb, wb, t);
          v[RA + 1] = pk_fma(make_float2(-xb.x, -xb.y), wb, t);
        });
      });
    }

    // pass 1: 32-point DFT over a (stride L), then twiddle W_M^(b*kA)
    if constexpr (!(ABL & 64)) fft_dit<32, 0, 32, 2>(v);
    {
      if constexpr (!TWFOLD) {
      TmemChunk tc[2];
      tc[0].issue(tmem_row + 64);
      static_for<0, 4>([&](auto c_) {
        constexpr int CH = decltype(c_)::value;
        tc[CH & 1].wait();
        if constexpr (CH < 3) tc[(CH + 1) & 1].issue(tmem_row + 64 + 16 * (CH + 1));
        static_for<0, 8>([&](auto i_) {
          constexpr int KA = 8 * CH + decltype(i_)::value;
          if constexpr (KA > 0) v[KA] = cmul(v[KA], tc[CH & 1].get(KA % 8));
        });
      });
      }
    }

    // transpose through the warp's scratch (= its own power rows)
    if constexpr (T64) {
      // n_fft 512 / 1024: complex words, rows kA < 16 then kA >= 16 (a lane's rows j + L q fall into the half q / (Q / 2));
      // half the instructions of the re / im form below for the same 128 wavefronts.  16 rows x 33 float2 = the scratch.
      float2* scr2 = reinterpret_cast<float2*>(scr);
      float2* w2 = scr2 + lane;
      const float2* r2 = scr2 + j * 33 + g * L;
      static_for<0, 2>([&](auto h_) {
        constexpr int H = decltype(h_)::value;
        static_for<0, 16>([&](auto k_) {
          constexpr int KA = 16 * H + decltype(k_)::value;
          w2[(KA - 16 * H) * 33] = v[KA];
        });
        __syncwarp();
        static_for<0, Q / 2>([&](auto q_) {
          constexpr int QL = decltype(q_)::value, QQ = H * (Q / 2) + QL;
          static_for<0, L>([&](auto b_) {
            constexpr int BB = decltype(b_)::value;
            v[QQ * L + bitrev(BB, LOG2L)] = r2[QL * L * 33 + BB];
          });
        });
        __syncwarp();
      });
    } else {
    float* scr_w = scr + lane;
    const float* scr_r = scr + j * 33 + g * L;
    if constexpr (!(ABL & 4)) {
    static_for<0, 32>([&](auto k_) {
      constexpr int KA = decltype(k_)::value;
      scr_w[KA * 33] = v[KA].x;
    });
    __syncwarp();
    static_for<0, Q>([&](auto q_) {
      constexpr int QQ = decltype(q_)::value;
      static_for<0, L>([&](auto b_) {
        constexpr int BB = decltype(b_)::value;
        v[QQ * L + bitrev(BB, LOG2L)].x = scr_r[QQ * L * 33 + BB];
      });
    });
    __syncwarp();
    static_for<0, 32>([&](auto k_) {
      constexpr int KA = decltype(k_)::value;
      scr_w[KA * 33] = v[KA].y;
    });
    __syncwarp();
    static_for<0, Q>([&](auto q_) {
      constexpr int QQ = decltype(q_)::value;
      static_for<0, L>([&](auto b_) {
        constexpr int BB = decltype(b_)::value;
        v[QQ * L + bitrev(BB, LOG2L)].y = scr_r[QQ * L * 33 + BB];
      });
    });
    __syncwarp();
    }
    }

    // pass 2: Q DFTs of length L over b  ->  v[q*L + kB] = Z[(j + L q) + 32 kB]
    if constexpr (TWFOLD) {
      // twiddle + first stage of pass 2: registers 2m, 2m + 1 hold b and b + 16;  a' = ta a + tb b,  b' = ta a - tb b
      TmemChunk tc[2];
      tc[0].issue(tmem_row + 64);
      static_for<0, 4>([&](auto c_) {
        constexpr int CH = decltype(c_)::value;
        tc[CH & 1].wait();
        if constexpr (CH < 3) tc[(CH + 1) & 1].issue(tmem_row + 64 + 16 * (CH + 1));
        static_for<0, 4>([&](auto i_) {
          constexpr int I = decltype(i_)::value;
          constexpr int BA = 4 * CH + I, RA = bitrev(BA, 5);
          const float2 tb = tc[CH & 1].get(2 * I + 1), b0 = v[RA + 1];
          float2 u = v[RA];
          if constexpr (BA > 0) u = cmul(u, tc[CH & 1].get(2 * I));
          const float2 t1 = __ffma2_rn(b0, make_float2(tb.x, tb.x), u);
          const float2 na = __ffma2_rn(make_float2(-b0.y, b0.x), make_float2(tb.y, tb.y), t1);
          v[RA + 1] = __ffma2_rn(u, make_float2(2.0f, 2.0f), make_float2(-na.x, -na.y));
          v[RA] = na;
        });
      });
      if constexpr (!(ABL & 64)) fft_dit<32, 0, 32, 2>(v);
    } else
    static_for<0, Q>([&](auto q_) {
      constexpr int QQ = decltype(q_)::value;
      if constexpr (!(ABL & 64)) fft_dit<L, QQ * L>(v);
    });

    // real-input split + power:  X[k] = E - T,  X[M-k] = conj(E + T),  T = i*w*O
    TmemChunk pc[2];
    if constexpr (TWPTM) {
      pc[0].issue(tmem_row + 128);
      pc[1].issue(tmem_row + 144);
      pc[0].wait();  // covers both
      pc[1].wait();
    }
    static_for<0, Q>([&](auto q_) {
      constexpr int QQ = decltype(q_)::value;
      static_for<0, L / 2>([&](auto s_) {
        constexpr int S = decltype(s_)::value;
        constexpr int GEN = (Q - 1 - QQ) * L + (L - 1 - S);
        constexpr int ALT = QQ == 0 ? ((L - S) % L) : (Q - QQ) * L + (L - 1 - S);
        const float2 snd = (j == 0) ? v[ALT] : v[GEN];
        float2 r;  // Z[M - k]
        if constexpr (ABL & 16) r = snd;
        else {
        r.x = __shfl_sync(0xffffffffu, snd.x, partner);
        r.y = __shfl_sync(0xffffffffu, snd.y, partner);
        }
        const float2 A = v[QQ * L + S];
        const float2 E = __fadd2_rn(A, make_float2(r.x, -r.y));   // A + conj(r)
        const float2 O = __fadd2_rn(A, make_float2(-r.x, r.y));   // A - conj(r)
        const int k = j + L * QQ + 32 * S;
        float2 wO;
        if constexpr (TWPTM) {
          constexpr int EI = QQ * (L / 2) + S;
          wO = cmul(O, pc[EI / 8].get(EI % 8));
        } else
        if constexpr (TWPGEN) wO = cmul(cmul_const<32 * S, N>(O), twp_base[QQ]);
        else wO = (ABL & 16) ? cmul(O, make_float2(a.amin, a.eps)) : cmul(O, sTwp[k]);
        const float2 x1 = __fadd2_rn(E, make_float2(wO.y, -wO.x));  // E - i*wO
        const float2 x2 = __fadd2_rn(E, make_float2(-wO.y, wO.x));  // E + i*wO
        if constexpr (ABL & 32) {
          if (x1.x * x2.y == 123.456f) prow[k] = x1.y;
        } else {
        prow[k] = spec(x1.x, x1.y);
        prow[M - k] = spec(x2.x, x2.y);
        }
      });
    });
    if (j == 0) {
      float2 A = v[L / 2];
      prow[M / 2] = mag ? 2.0f * spec(A.x, A.y) : 4.0f * spec(A.x, A.y);
      // padding read by the last tap groups of a segment that reaches the Nyquist bin
#pragma unroll
      for (int i = 1; i <= NPAD; ++i) prow[M + i] = 0.f;
    }
  }
};

// FBM: 0 = two-tap banded filter bank (mel / linear triangles);
//      1 = dense filter bank (gammatone): entries hold the full rows of TWO filters.  The 42 KB of weights live in shared
//          memory like the banded program (warp-uniform LDG.128 through L1 cost four wavefronts each and made the
//          first version L1-bound at 91 %), which is why this variant runs as two 8-warp CTAs per SM.
template <int L, int MODE, bool PRE, int TILE, bool PAIR = false, int FBM = 0>
__global__ void __launch_bounds__(StftCfg<L, TILE, FBM != 0>::WARPS * 32, StftCfg<L, TILE, FBM != 0>::CTAS)
k_stft_fb(const StftArgs a) {
  static_assert(!(PAIR && FBM), "paired plans are banded");
  using C = StftCfg<L, TILE, FBM != 0>;
  using FFT = FrameFft<L, MODE, PRE, TILE>;
  constexpr int Q = C::Q, M = C::M, N = C::N, SP = C::SP, FBU = FBM ? kDenseFbu : C::FBU;
  extern __shared__ __align__(16) float smem[];
  float* sP = smem + C::OFF_P;
  float2* sTwp = reinterpret_cast<float2*>(smem + C::OFF_TWP);
  int* sMeta = reinterpret_cast<int*>(smem + C::OFF_META);
  int2* sHdr = reinterpret_cast<int2*>(smem + C::OFF_PROG);
  float4* sW4 = reinterpret_cast<float4*>(smem + C::OFF_PROG + ((2 * a.n_hdr + 3) & ~3));

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane / L;
  // TMA staging (dev switch): [mbarrier (8 B) | a0, bytes, src lo, src hi per meta buffer | ... 16 floats] then the samples
  constexpr bool STAGE = AAD_TMA_STAGE && L == 32 && MODE == IN_F32 && !PRE && !PAIR && FBM == 0;
  float* sStageHdr = reinterpret_cast<float*>(sW4 + a.n_w4);
  int* sInfo = reinterpret_cast<int*>(sStageHdr) + 2;
  float* sS = sStageHdr + 16;
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(sStageHdr);
  const bool stage_on = STAGE && a.row_off == nullptr && (a.wav_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(a.wav) & 15) == 0;
  if constexpr (STAGE) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      sInfo[0] = -1;
      sInfo[4] = -1;
    }
  }

  if constexpr (C::SMEM_TWP)
    for (int i = tid; i < M / 2; i += nthr) sTwp[i] = a.twp[i];
  for (int i = tid; i < a.n_hdr; i += nthr) sHdr[i] = a.filt_hdr[i];
  for (int i = tid; i < a.n_w4; i += nthr) sW4[i] = a.filt_w[i];
  if constexpr (PAIR) {  // second program behind the first (pointers are re-derived where it runs: no live registers)
    int2* sHdr2 = reinterpret_cast<int2*>(sW4 + a.n_w4);
    float4* sW42 = reinterpret_cast<float4*>(reinterpret_cast<float*>(sHdr2) + ((2 * a.n_hdr2 + 3) & ~3));
    for (int i = tid; i < a.n_hdr2; i += nthr) sHdr2[i] = a.filt_hdr2[i];
    for (int i = tid; i < a.n_w42; i += nthr) sW42[i] = a.filt_w2[i];
  }
  // allocate the TMEM columns, fill this CTA's four lane quarters with the per-lane tables
  volatile uint32_t& s_tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(smem + C::OFF_TMEM)), "r"((uint32_t)TmemCfg<C::CTAS>::COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_row = s_tmem_base + (((uint32_t)(warp & 3) * 32u) << 16);
  if (warp < 4) FFT::fill_tables(a, tmem_row, lane % L);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  FFT fft;
  fft.init(a, lane, tmem_row, sTwp);
  const int4 wprog = a.warp_prog[warp];
  // everything above read plan tables only; k_prepare's results (frame_off, tile_b0, len_c, utt_max reset) from here on
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the epilogue kernel may queue up behind this grid
  const int total = a.frame_off[a.B];
  const int n_tiles = (total + C::TILE - 1) / C::TILE;

  // tile meta: global frame index -> (utterance, frame); computed one tile ahead by warp 0 so the
  // dependent index loads never sit on the critical path, and used to prefetch the next tile's
  // new samples into L2 (bulk prefetch: one instruction per frame)
  auto tile_meta = [&](int tile, int buf) {
    int4 r0 = make_int4(0, 0, 0, 0), r1 = r0;
    longlong2 r2 = make_longlong2(0, 0);
    if (tile < n_tiles) {
      r0 = __ldg(a.tile_rec + 3 * tile);
      r1 = __ldg(a.tile_rec + 3 * tile + 1);
      r2 = __ldg(reinterpret_cast<const longlong2*>(a.tile_rec + 3 * tile + 2));
    }
#pragma unroll
    for (int f = 0; f < C::FPL; ++f) {
      const int gf = tile * C::TILE + f * 32 + lane;
      int b = -1, t = 0;
      long long ro = 0;
      int len = 0;
      if (tile < n_tiles && gf < total) {
        if (gf < r0.z) {
          b = r0.x; t = gf - r0.y; len = r1.x; ro = r2.x;
        } else if (gf < r0.w) {
          b = r0.x + 1; t = gf - r0.z; len = r1.y; ro = r2.y;
        } else {  // utterances shorter than a tile: walk on (dependent loads, rare)
          b = r0.x + 1;
          int nxt = r0.w;
          while (gf >= nxt) nxt = __ldg(a.frame_off + (++b) + 1);
          t = gf - __ldg(a.frame_off + b);
          len = __ldg(a.len_c + b);
          ro = a.row_off ? __ldg(a.row_off + b) : (long long)b * a.wav_stride;
        }
      }
      sMeta[buf * 2 * C::TILE + f * 32 + lane] = b;
      sMeta[buf * 2 * C::TILE + C::TILE + f * 32 + lane] = t;
      if (b >= 0) {
        constexpr int ES = MODE == IN_I16 ? 2 : 4;
        long long s_lo = (long long)t * a.hop - a.s_off + (t == 0 ? 0 : N - a.hop);
        long long s_hi = (long long)t * a.hop - a.s_off + N;
        if (s_lo < 0) s_lo = 0;
        if (s_hi > len) s_hi = len;
        const char* row = static_cast<const char*>(a.wav) + ro * ES;
        uintptr_t p0 = ((uintptr_t)(row + s_lo * ES) + 15) & ~(uintptr_t)15;
        uintptr_t p1 = (uintptr_t)(row + s_hi * ES) & ~(uintptr_t)15;
        if (p1 > p0)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0), "r"((unsigned)(p1 - p0)) : "memory");
      }
    }
    if constexpr (STAGE) {
      // a tile whose frames all belong to one utterance is staged as ONE contiguous sample range
      const int b = sMeta[buf * 2 * C::TILE + lane], t = sMeta[buf * 2 * C::TILE + C::TILE + lane];
      const unsigned vm = __ballot_sync(0xffffffffu, b >= 0);
      const int b0 = __shfl_sync(0xffffffffu, b, 0), t0 = __shfl_sync(0xffffffffu, t, 0);
      const bool one = stage_on && vm != 0 && (vm & 1u) && __all_sync(0xffffffffu, b < 0 || b == b0);
      if (lane == 0) {
        int a0 = -1, bytes = 0;
        unsigned long long src = 0;
        if (one) {
          const int len = __ldg(a.len_c + b0), t1 = t0 + __popc(vm) - 1;
          long long lo = (long long)t0 * a.hop - a.s_off, hi = (long long)t1 * a.hop - a.s_off + N;
          lo = max(lo, 0ll) & ~3ll;
          hi = min(min(hi, (long long)len) + 3 & ~3ll, a.wav_stride);
          if (hi > lo) {
            a0 = (int)lo;
            bytes = (int)(hi - lo) * 4;
            src = reinterpret_cast<unsigned long long>(static_cast<const float*>(a.wav) + (long long)b0 * a.wav_stride + lo);
          }
        }
        sInfo[buf * 4] = a0;
        sInfo[buf * 4 + 1] = bytes;
        sInfo[buf * 4 + 2] = (int)(unsigned)(src & 0xffffffffull);
        sInfo[buf * 4 + 3] = (int)(unsigned)(src >> 32);
      }
      __syncwarp();
    }
  };
  // issue the bulk copy of the tile whose meta sits in `buf` (thread 0, after the barrier that ended all reads of sS)
  auto stage_issue = [&](int buf) {
    if constexpr (STAGE) {
      if (tid == 0 && sInfo[buf * 4] >= 0) {
        const unsigned bytes = (unsigned)sInfo[buf * 4 + 1];
        const unsigned long long src = ((unsigned long long)(unsigned)sInfo[buf * 4 + 3] << 32) | (unsigned)sInfo[buf * 4 + 2];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(sS)), "l"(src), "r"(bytes), "r"(mbar) : "memory");
      }
    }
  };
  if (warp == 0) tile_meta(blockIdx.x, 0);
  __syncthreads();
  stage_issue(0);
  unsigned stage_phase = 0;


#ifdef AAD_PHASE_TIMING
  long long tmark = clock64();
#endif
  int buf = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    const int* sMetaB = sMeta + buf * 2 * C::TILE;
    const int* sMetaT = sMetaB + C::TILE;
    if (warp == 0) tile_meta(tile + gridDim.x, buf ^ 1);  // consumed after the bottom barrier

    const float* stage_ptr = nullptr;
    int stage_a0 = 0;
    if constexpr (STAGE) {
      stage_a0 = sInfo[buf * 4];
      if (stage_a0 >= 0) {  // this tile's samples were requested during the previous tile's filterbank phase
        unsigned done = 0;
        for (int spins = 0; !done; ++spins) {
          asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                       : "=r"(done) : "r"(mbar), "r"(stage_phase) : "memory");
          if (spins > (1 << 22)) __trap();   // a copy that never lands must not hang the device
        }
        stage_phase ^= 1;
        stage_ptr = sS;
      }
    }
    // ---- FFT phase ---------------------------------------------------------------
    for (int it = warp; it < C::ITERS; it += C::WARPS) {
      float2 v[32];
      // transposes go through the warp's own power rows
      if (fft.load_iter(a, sMetaB, sMetaT, it, v, stage_ptr, stage_a0))
        fft.template transform<C::PAD, FBM != 0>(a, v, sP + (it * Q) * SP, sP + (it * Q + g) * SP + C::skew(it * Q + g));
    }
    AAD_PHASE_MARK(0);
    __syncthreads();
    stage_issue(buf ^ 1);   // the next tile's samples travel during this tile's filterbank phase

    // ---- filterbank + log phase: lane = frame, warps split the filters ----
    // Entries are walked FBU at a time (one bundle).  Per round and entry: 4 bins (powers: one
    // conflict-free LDS.128 per lane; weights: two LDS.128 broadcasts; four FFMA2 accumulate the
    // (rising, falling) sums with the power broadcast to both halves), then filter s-1 = R[s-1] + F[s]:
    // log, store, running max.  Headers carry byte offsets so that a bundle starts with two adds.
    auto fb_phase = [&](const int4 wprog, const int2* sHdr, const float4* sW4, float* Eout, long long e_stride_b,
                        int e_stride_f, int log_type, float amin, int32_t* utt_max) {
    if (wprog.w > 0 && !(ABL & 8)) {
      constexpr int FPL = C::FPL;
      int b[FPL];
      bool valid[FPL];
      const char* pbase[FPL];
      float* eptr[FPL];
      float vmax[FPL], chk[FPL], rprev[FPL];
#pragma unroll
      for (int f = 0; f < FPL; ++f) {
        b[f] = sMetaB[f * 32 + lane];
        const int t = sMetaT[f * 32 + lane];
#ifdef AAD_PHASE_TIMING
        if (b[f] == 0x7fffffff) return;  // forces the first post-barrier load to complete: the barrier wait ends here
#endif
        valid[f] = b[f] >= 0;
        pbase[f] = reinterpret_cast<const char*>(sP + (f * 32 + lane) * SP + C::skew(f * 32 + lane));
        // entry i of the list emits filter wf0 + i - 1 (banded) / filters wf0 + 2 i, wf0 + 2 i + 1 (dense)
        eptr[f] = Eout + (valid[f] ? (long long)b[f] * e_stride_b + (long long)(wprog.x - (FBM ? 0 : 1)) * e_stride_f + t : 0);
        vmax[f] = -INFINITY;
        chk[f] = 0.f;
        rprev[f] = 0.f;
      }
#ifdef AAD_PHASE_TIMING
      AAD_PHASE_MARK(1);
#endif
      const char* wbase = reinterpret_cast<const char*>(sW4);
      const long long estep = e_stride_f;
      const bool is_db = log_type == 0;
      const float lscale = is_db ? 3.01029995663981195f : 0.69314718055994531f;
      const float amin_n = fmaxf(amin, 1.17549435e-38f);  // keeps MUFU.LG2 off the denormal path
      const int2* hp = sHdr + wprog.y;
      int2 hd[FBU];
#pragma unroll
      for (int u = 0; u < FBU; ++u) hd[u] = hp[u];
      for (int i0 = 0; i0 < wprog.z; i0 += FBU) {
        // the next bundle's headers travel while this bundle is evaluated (the walk is a chain of dependent
        // shared-memory reads otherwise: header -> round count / offsets -> taps)
        hp += (i0 + FBU < wprog.z) ? FBU : 0;
        int2 hn[FBU];
#pragma unroll
        for (int u = 0; u < FBU; ++u) hn[u] = hp[u];
        int poff[FBU];
        float2 acc0[FPL][FBU], acc1[FPL][FBU];
#pragma unroll
        for (int u = 0; u < FBU; ++u) {
          poff[u] = hd[u].x & 0xffff;
#pragma unroll
          for (int f = 0; f < FPL; ++f) {
            acc0[f][u] = make_float2(0.f, 0.f);
            acc1[f][u] = make_float2(0.f, 0.f);
          }
        }
        const char* wp = wbase + hd[0].y;  // the bundle's weights are interleaved: [round][entry][2 x float4]
        for (int gq = hd[0].x >> 16; gq > 0; --gq) {  // same round count for the whole bundle
          float4 pv[FPL];
          if constexpr (FBM) {  // dense: every entry of the bundle starts at bin 0, one group serves them all
#pragma unroll
            for (int f = 0; f < FPL; ++f) pv[f] = *reinterpret_cast<const float4*>(pbase[f] + poff[0]);
            poff[0] += 16;
          }
#pragma unroll
          for (int u = 0; u < FBU; ++u) {
            const float4 wa = *reinterpret_cast<const float4*>(wp + 32 * u);
            const float4 wb = *reinterpret_cast<const float4*>(wp + 32 * u + 16);
#pragma unroll
            for (int f = 0; f < FPL; ++f) {
              float4 p;
              if constexpr (FBM) p = pv[f];
              else p = *reinterpret_cast<const float4*>(pbase[f] + poff[u]);
              acc0[f][u] = __ffma2_rn(make_float2(p.x, p.x), make_float2(wa.x, wa.y), acc0[f][u]);
              acc1[f][u] = __ffma2_rn(make_float2(p.y, p.y), make_float2(wa.z, wa.w), acc1[f][u]);
              acc0[f][u] = __ffma2_rn(make_float2(p.z, p.z), make_float2(wb.x, wb.y), acc0[f][u]);
              acc1[f][u] = __ffma2_rn(make_float2(p.w, p.w), make_float2(wb.z, wb.w), acc1[f][u]);
            }
            if constexpr (!FBM) poff[u] += 16;
          }
          wp += 32 * FBU;
        }
        // entries that emit nothing (the first one of the list, the padding behind the last one) have had the taps of
        // the neighbouring warps' filters zeroed by the plan: their value is log(floor), harmless for the running
        // maximum and the poison check, so that only the store depends on the entry index
        auto nonlin = [&](float en, float& bad) {
          float val;
          bad = en;  // dB: max(amin, NaN) hides a NaN energy, so the energy itself is the poison source
          // 10*log10(x) = 3.0103*log2(x), ln(x) = 0.6931*log2(x); MUFU.LG2 is accurate to 2 ulp,
          // i.e. <= 3e-5 dB / 4e-6 nepers here, far inside the 1e-3 parity tolerance
          if constexpr (ABL & 256) {
            val = en;
          } else if (is_db) {
            float l2;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(fmaxf(amin_n, en)));
            val = lscale * l2;
          } else if (FBM && log_type == 2) {
            val = cbrtf(en);  // spafe gfcc: np.power(features, 1 / 3)
          } else {
            val = lscale * __log2f(en == 0.f ? a.eps : en);
            bad = val;  // ln: also catches the log of a negative energy (custom filter banks)
          }
          return val;
        };
#pragma unroll
        for (int f = 0; f < FPL; ++f) {
#pragma unroll
          for (int u = 0; u < FBU; ++u) {
            const float2 rf = __fadd2_rn(acc0[f][u], acc1[f][u]);
            if constexpr (FBM) {  // the two filters of entry i0 + u
              float bad0, bad1;
              const float v0 = nonlin(rf.x, bad0), v1 = nonlin(rf.y, bad1);
              const int fi_ = 2 * (i0 + u);
              if (valid[f] && fi_ < wprog.w) eptr[f][(2 * u) * estep] = v0;
              if (valid[f] && fi_ + 1 < wprog.w) eptr[f][(2 * u + 1) * estep] = v1;
              vmax[f] = fmaxf(vmax[f], fmaxf(v0, v1));
              chk[f] = __fmaf_rn(bad0, 0.f, __fmaf_rn(bad1, 0.f, chk[f]));
            } else {              // (R, F) of entry i0 + u: filter wf0 + i0 + u - 1 = R[i0 + u - 1] + F[i0 + u]
              const float en = rprev[f] + rf.y;
              rprev[f] = rf.x;
              float bad;
              const float val = nonlin(en, bad);
              const int fi_ = i0 + u;  // emits filter wf0 + fi_ - 1 when 1 <= fi_ <= n_filters
              if (valid[f] && fi_ >= 1 && fi_ <= wprog.w) eptr[f][u * estep] = val;
              vmax[f] = fmaxf(vmax[f], val);
              chk[f] = __fmaf_rn(bad, 0.f, chk[f]);  // NaN/Inf poison
            }
          }
          eptr[f] += (FBM ? 2 : 1) * FBU * estep;
        }
#pragma unroll
        for (int u = 0; u < FBU; ++u) hd[u] = hn[u];
      }
#pragma unroll
      for (int f = 0; f < FPL; ++f) {
        if (valid[f]) {
          if (utt_max) {
            unsigned peers = __match_any_sync(__activemask(), b[f]);
            int enc = __reduce_max_sync(peers, enc_ordered(vmax[f]));
            if ((int)(__ffs(peers) - 1) == lane) atomicMax(utt_max + b[f], enc);
          }
          if (chk[f] != chk[f]) a.status[b[f]] = 5;
        }
      }
    }
    };
    fb_phase(wprog, sHdr, sW4, a.E, a.e_stride_b, a.e_stride_f, a.log_type, a.amin, a.utt_max);
    if constexpr (PAIR) {
      const int2* sHdr2 = reinterpret_cast<const int2*>(sW4 + a.n_w4);
      const float4* sW42 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sHdr2) + ((2 * a.n_hdr2 + 3) & ~3));
      fb_phase(__ldg(a.warp_prog2 + warp), sHdr2, sW42, a.E2, a.e2_stride_b, a.e2_stride_f, a.log_type2, a.amin2, a.utt_max2);
    }
    AAD_PHASE_MARK(2);
    __syncthreads();  // sP free for the next tile's FFTs; next tile's meta (written above) visible
#ifdef AAD_PHASE_TIMING
    if (sMeta[(buf ^ 1) * 2 * C::TILE] == 0x7fffffff) return;
    AAD_PHASE_MARK(3);
#endif
  }
  // every warp is past its last tcgen05.ld (the loop ends with a CTA barrier; CTAs without tiles
  // come straight from the set-up barrier)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"((uint32_t)s_tmem_base), "r"((uint32_t)TmemCfg<C::CTAS>::COLS) : "memory");
}

#ifndef AAD_STFT_ONLY
// ---------------------------------------------------------------------------
// K2: dB reference / floor + DCT-II + deltas + layout
// ---------------------------------------------------------------------------
constexpr int CEP_TS = 128;       // frames per tile held in smem
constexpr int CEP_SE = CEP_TS + 8;  // row stride of the energy tile: A fragments (filter t, frame g) hit 32 banks
constexpr int CEP_SC = CEP_TS + 4;  // row stride of the coefficient tile: C fragments (coef 2t, frame g) hit 32 banks
constexpr int CEP_MAXW = 9;
constexpr int CEP_THREADS = 256;  // 8 warps = 8 m-tiles of 16 frames
constexpr int CEP_MAXNT = 8;      // coefficient n-tiles (of 8) whose accumulators a warp holds at once

struct CepArgs {
  const float* E;         // [B][n_filt][e_stride_f]
  long long e_stride_b;
  int e_stride_f;
  const int32_t* nf_eff;  // [B]
  const int32_t* utt_max; // ordered-int encoded (dB only)
  int n_filt;
  int log_type, ref_type;
  float top_db;           // < 0: none
  int n_ceps;             // 0: identity
  int n_ksteps, n_tiles;  // DCT as GEMM: K = n_filt in steps of 8, N = n_ceps in tiles of 8
  const float4* dct_frag; // [n_ksteps][n_tiles][32] B fragments {b0 hi, b1 hi, b0 lo, b1 lo} (tf32 split)
  const float* dct_colsum; // [n_tiles * 8] column sums of the table (mean re-addition)
  int n_delta, width;
  float taps[2][CEP_MAXW];
  float* out;
  long long out_stride_b;
  int out_stride_c, out_stride_t;  // CT: (t_alloc, 1); TC: (1, c_out)
  int tile_out;           // output frames per tile when T > CEP_TS
  int tiles_per_utt;      // grid.x = B * tiles_per_utt
};

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One CTA = (utterance, tile of <= 128 frames).  Three phases:
//  load : E tile -> sE[m][t] with the dB reference / floor applied (float4 along t, warp w owns the
//         filterbank rows m = w (mod 8)), per-warp partial column sums for the per-frame mean
//  DCT  : C[frame][coef] = (E - mean)[frame][filter] . D^T[filter][coef] as a warp-level GEMM on the
//         tensor cores, mma.sync.m16n8k8 TF32 with the 3-term split (hi*hi + lo*hi + hi*lo, fp32
//         accumulate: ~2^-21 relative per product, inside the 1e-3 tolerance by three orders).  Warp w
//         owns the 16 frames [16w, 16w+16) and all coefficient tiles; the table comes pre-split and
//         pre-arranged as B fragments, the energies are centred (per-frame mean, re-added through the
//         column sums) and split in registers.  The result overwrites the energy tile.
//  out  : thread = frame column; static rows and the delta / delta-delta stencils as one FFMA2 per
//         tap ((d1, d2) accumulated together), coalesced stores along t (CT) or along c (TC)
// TS: frames per tile (128, or 64 when no utterance of the batch is longer: the reference's 2-second chunks have 63 frames,
// and half of a 128-frame tile's warps, shared memory and load lanes would idle); 2 TS threads = TS / 16 warps.
template <int NT, int TS = CEP_TS>
__global__ void __launch_bounds__(2 * TS, TS == 128 ? 2 : 4) k_cepstra(const CepArgs a) {
  constexpr int SE = TS + 8, SC = TS + 4, THREADS = 2 * TS;
  constexpr int LPR = TS / 4, RPW = 32 / LPR;   // load phase: lanes per row, rows per warp and step (8 rows per CTA step)
  extern __shared__ __align__(16) float smem[];
  asm volatile("griddepcontrol.wait;" ::: "memory");  // launched behind k_stft_fb (programmatic dependent launch)
  // grid.x = B * tiles_per_utt, utterance-major: consecutive CTAs read neighbouring tiles of one utterance
  const int b = blockIdx.x / a.tiles_per_utt, tile = blockIdx.x - b * a.tiles_per_utt;
  const int T = a.nf_eff[b];
  if (T == 0) return;
  int o0, o1;
  if (T <= TS) {
    if (tile > 0) return;
    o0 = 0;
    o1 = T;
  } else {
    o0 = tile * a.tile_out;
    if (o0 >= T) return;
    o1 = min(T, o0 + a.tile_out);
  }
  const int h = a.width >> 1;
  int lo = o0, hi = o1;
  if (a.n_delta > 0) {
    int c0 = min(max(o0, h), T - 1 - h), c1 = min(max(o1 - 1, h), T - 1 - h);
    lo = min(lo, c0 - h);
    hi = max(hi, c1 + h + 1);
  }
  const int nload = hi - lo;  // <= TS by construction
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = a.n_ceps > 0 ? a.n_ceps : a.n_filt;
  const int kf = a.n_ceps > 0 ? 8 * a.n_ksteps : a.n_filt;  // energy rows incl. zero padding of K
  const bool alias = a.n_tiles <= CEP_MAXNT;                 // coefficients overwrite the energy tile

  // 4 floats of slack in front: the fixed 9-tap stencil of a narrower delta window may read up to 3
  // columns before a row (with a zero tap)
  float* sE = smem + 4;                               // [kf][SE]
  float* sPart = sE + kf * SE;                    // [8][TS] per-warp partial column sums
  float4* sB = reinterpret_cast<float4*>(sPart + 8 * TS);  // [n_ksteps][n_tiles][32]
  float* sCs = reinterpret_cast<float*>(sB + a.n_ksteps * a.n_tiles * 32);  // [n_tiles * 8]
  float* sC = a.n_ceps == 0 ? sE : (alias ? sE : sCs + a.n_tiles * 8);      // [C][sc_stride]
  const int sc_stride = a.n_ceps == 0 ? SE : SC;

  // reference / floor (librosa.power_to_db):  ls = E - ref ; ls = max(ls, max(ls) - top_db)
  float ref = 0.f, floorv = -INFINITY;
  if (a.log_type == 0) {
    float m = dec_ordered(a.utt_max[b]);
    if (a.ref_type == 1) ref = m;
    if (a.top_db >= 0.f) floorv = (m - ref) - a.top_db;
  }
  // ---- load ---------------------------------------------------------------------------
  {
    const float* Eb = a.E + (long long)b * a.e_stride_b + lo;
    const bool vec = ((lo | a.e_stride_f) & 3) == 0 && (reinterpret_cast<uintptr_t>(a.E) & 15) == 0 &&
                     (a.e_stride_b & 3) == 0;
    const int t4 = 4 * (lane % LPR), rsub = lane / LPR;   // this lane's 4 columns and its row within the warp's step
    float4 ps = make_float4(0.f, 0.f, 0.f, 0.f);
    auto xf = [&](float v, int t) { return t < nload ? fmaxf(v - ref, floorv) : 0.f; };
    if (vec) {
      int m = warp * RPW + rsub;
      for (; m + 56 < a.n_filt; m += 64) {  // 8 independent 16-byte loads in flight per thread
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          v[u] = t4 < nload ? __ldg(reinterpret_cast<const float4*>(Eb + (long long)(m + 8 * u) * a.e_stride_f + t4))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float4 r = make_float4(xf(v[u].x, t4), xf(v[u].y, t4 + 1), xf(v[u].z, t4 + 2), xf(v[u].w, t4 + 3));
          *reinterpret_cast<float4*>(sE + (m + 8 * u) * SE + t4) = r;
          ps.x += r.x; ps.y += r.y; ps.z += r.z; ps.w += r.w;
        }
      }
      for (; m < a.n_filt; m += 8) {
        float4 v = t4 < nload ? __ldg(reinterpret_cast<const float4*>(Eb + (long long)m * a.e_stride_f + t4))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 r = make_float4(xf(v.x, t4), xf(v.y, t4 + 1), xf(v.z, t4 + 2), xf(v.w, t4 + 3));
        *reinterpret_cast<float4*>(sE + m * SE + t4) = r;
        ps.x += r.x; ps.y += r.y; ps.z += r.z; ps.w += r.w;
      }
    } else {
      for (int m = warp * RPW + rsub; m < a.n_filt; m += 8) {
        float r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int t = t4 + u;
          r[u] = t < nload ? fmaxf(__ldg(Eb + (long long)m * a.e_stride_f + t) - ref, floorv) : 0.f;
        }
        *reinterpret_cast<float4*>(sE + m * SE + t4) = make_float4(r[0], r[1], r[2], r[3]);
        ps.x += r[0]; ps.y += r[1]; ps.z += r[2]; ps.w += r[3];
      }
    }
    for (int m = a.n_filt + warp * RPW + rsub; m < kf; m += 8)  // zero rows that pad K to a multiple of 8
      *reinterpret_cast<float4*>(sE + m * SE + t4) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(sPart + (warp * RPW + rsub) * TS + t4) = ps;
    if (a.n_ceps > 0) {
      const int nb = a.n_ksteps * a.n_tiles * 32;
      for (int i = tid; i < nb; i += THREADS) sB[i] = __ldg(a.dct_frag + i);
      for (int i = tid; i < a.n_tiles * 8; i += THREADS) sCs[i] = __ldg(a.dct_colsum + i);
    }
  }
  __syncthreads();

  // ---- DCT on the tensor cores ------------------------------------------------------------
  if (a.n_ceps > 0) {
    const int g = lane >> 2, t = lane & 3;
    const int f0 = 16 * warp + g, f1 = f0 + 8;  // the two frame rows of this thread's fragments
    float mean0 = 0.f, mean1 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      mean0 += sPart[w * TS + f0];
      mean1 += sPart[w * TS + f1];
    }
    const float inv = 1.0f / (float)a.n_filt;
    mean0 *= inv;
    mean1 *= inv;
    const float* ea = sE + t * SE + f0;  // A fragment (row g / g+8 = frame, col t / t+4 = filter)
    for (int nt0 = 0; nt0 < a.n_tiles; nt0 += NT) {
      float acc[NT][4];
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
      const float4* bp = sB + nt0 * 32 + lane;
#pragma unroll 2
      for (int ks = 0; ks < a.n_ksteps; ++ks) {
        const float* e = ea + ks * 8 * SE;
        // centred energies (small magnitudes: the tf32 split and the fp32 accumulation lose nothing)
        const float av[4] = {e[0] - mean0, e[8] - mean1, e[4 * SE] - mean0, e[4 * SE + 8] - mean1};
        unsigned ahi[4], alo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(ahi[i]) : "f"(av[i]));
          alo[i] = __float_as_uint(av[i] - __uint_as_float(ahi[i]));
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          if (nt0 + n < a.n_tiles) {
            const float4 q = bp[(ks * a.n_tiles + n) * 32];
            mma_tf32(acc[n], ahi, __float_as_uint(q.x), __float_as_uint(q.y));
            mma_tf32(acc[n], alo, __float_as_uint(q.x), __float_as_uint(q.y));
            mma_tf32(acc[n], ahi, __float_as_uint(q.z), __float_as_uint(q.w));
          }
        }
      }
      if (alias) __syncthreads();  // single pass (n_tiles <= NT): every warp is done reading the energies
      // C fragment: rows g / g+8 = frames f0 / f1, cols 2t, 2t+1 = coefficients
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const int c = (nt0 + n) * 8 + 2 * t;
        if (nt0 + n < a.n_tiles) {
          const float cs0 = sCs[c], cs1 = sCs[c + 1];
          if (c < a.n_ceps) {
            sC[c * SC + f0] = __fmaf_rn(mean0, cs0, acc[n][0]);
            sC[c * SC + f1] = __fmaf_rn(mean1, cs0, acc[n][2]);
          }
          if (c + 1 < a.n_ceps) {
            sC[(c + 1) * SC + f0] = __fmaf_rn(mean0, cs1, acc[n][1]);
            sC[(c + 1) * SC + f1] = __fmaf_rn(mean1, cs1, acc[n][3]);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- static rows + delta stencils + layout -------------------------------------------
  const int col = tid & (TS - 1), half = tid / TS;
  const int t = lo + col;
  if (t >= o0 && t < o1) {
    float* ob = a.out + (long long)b * a.out_stride_b + (long long)t * a.out_stride_t + (long long)half * a.out_stride_c;
    const long long cstep = 2ll * a.out_stride_c, dstep = (long long)C * a.out_stride_c;
    const int te = min(max(t, h), T - 1 - h) - lo;  // stencil centre (edges replicate the interior fit)
    // taps centred in a fixed 9-wide window: immediate offsets.  Positions outside the delta width are neither read
    // nor accumulated (they lie in the slack in front of a row or in its pad columns, which nothing writes: a
    // zero tap times a NaN bit pattern left there would poison the result)
    float2 tp[CEP_MAXW];
#pragma unroll
    for (int i = 0; i < CEP_MAXW; ++i) {
      const int src = i - (CEP_MAXW / 2) + h;
      tp[i] = (src >= 0 && src < a.width) ? make_float2(a.taps[0][src], a.taps[1][src]) : make_float2(0.f, 0.f);
    }
    const int i_lo = CEP_MAXW / 2 - h, i_hi = CEP_MAXW / 2 + h;
    const float* row = sC + half * sc_stride;
    for (int k = half; k < C; k += 2, row += 2 * sc_stride, ob += cstep) {
      ob[0] = row[col];
      if (a.n_delta > 0) {
        const float* rc = row + te - CEP_MAXW / 2;
        float2 d12 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < CEP_MAXW; ++i)
          if (i >= i_lo && i <= i_hi) {
            const float x = rc[i];
            d12 = __ffma2_rn(make_float2(x, x), tp[i], d12);
          }
        ob[dstep] = d12.x;
        if (a.n_delta > 1) ob[2 * dstep] = d12.y;
      }
    }
  }
}

// log-mel without DCT/deltas in CT layout: in-place reference subtraction + floor
struct FinArgs {
  float* out;
  long long stride_b;
  int stride_f;
  const int32_t* nf_eff;
  const int32_t* utt_max;   // ordered-int encoded (written by k_stft_fb) ...
  const float* utt_max_f;   // ... or plain floats from the caller (aad_db_reference); one of the two
  int n_filt, ref_type;
  float top_db;
};
constexpr int FIN_ROWS = 8;      // filter rows per CTA (one per warp)
constexpr int FIN_CHUNK = 1024;  // frames per CTA
// grid.x = B * n_row_blocks * n_chunks (flattened, utterance-major)
__global__ void __launch_bounds__(256) k_db_finalize(const FinArgs a, int n_row_blocks, int n_chunks) {
  asm volatile("griddepcontrol.wait;" ::: "memory");  // a no-op unless launched as a programmatic dependent
  const int per_b = n_row_blocks * n_chunks;
  const int b = blockIdx.x / per_b, r = blockIdx.x - b * per_b;
  const int f = (r / n_chunks) * FIN_ROWS + (threadIdx.x >> 5);
  const int t0 = (r % n_chunks) * FIN_CHUNK;
  const int T = a.nf_eff[b];
  if (f >= a.n_filt || t0 >= T) return;
  const float m = a.utt_max_f ? a.utt_max_f[b] : dec_ordered(a.utt_max[b]);
  const float ref = a.ref_type == 1 ? m : 0.f;
  const float floorv = a.top_db >= 0.f ? (m - ref) - a.top_db : -INFINITY;
  float* row = a.out + (long long)b * a.stride_b + (long long)f * a.stride_f;
  const int t1 = min(T, t0 + FIN_CHUNK);
  int t = t0 + (threadIdx.x & 31);
  for (; t + 96 < t1; t += 128) {  // four independent coalesced accesses in flight per lane
    float v0 = row[t], v1 = row[t + 32], v2 = row[t + 64], v3 = row[t + 96];
    row[t] = fmaxf(v0 - ref, floorv);
    row[t + 32] = fmaxf(v1 - ref, floorv);
    row[t + 64] = fmaxf(v2 - ref, floorv);
    row[t + 96] = fmaxf(v3 - ref, floorv);
  }
  for (; t < t1; t += 32) row[t] = fmaxf(row[t] - ref, floorv);
}

// The same for SHORT rows (the reference's 2-second chunks: 63 frames): one warp per row leaves a CTA with 2 KB to move
// and the launch of 200 000 such CTAs costs three times the memory time.  Here a CTA sweeps FIN_FLAT consecutive elements
// of the utterance's (filter, frame) block -- rows lie back to back at stride_f -- and masks the columns behind T.
// grid.x = B * n_chunks, n_chunks = ceil(n_filt * stride_f / FIN_FLAT).
constexpr int FIN_FLAT = 8192;
__global__ void __launch_bounds__(256) k_db_finalize_flat(const FinArgs a, int n_chunks) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int b = blockIdx.x / n_chunks, ch = blockIdx.x - b * n_chunks;
  const int T = a.nf_eff[b];
  const int total = a.n_filt * a.stride_f, i0 = ch * FIN_FLAT;
  if (T == 0 || i0 >= total) return;
  const int i1 = min(total, i0 + FIN_FLAT);
  const float m = a.utt_max_f ? a.utt_max_f[b] : dec_ordered(a.utt_max[b]);
  const float ref = a.ref_type == 1 ? m : 0.f;
  const float floorv = a.top_db >= 0.f ? (m - ref) - a.top_db : -INFINITY;
  float* base = a.out + (long long)b * a.stride_b;
  int i = i0 + threadIdx.x;
  int t = i % a.stride_f;                       // column of element i; advanced incrementally below
  const int step_t = 256 % a.stride_f;          // 256 elements further on: the column moves by 256 mod stride_f
  for (; i + 768 < i1; i += 1024) {             // four independent coalesced accesses in flight per lane
    int tc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      tc[u] = t;
      t += step_t;
      if (t >= a.stride_f) t -= a.stride_f;
    }
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = tc[u] < T ? base[i + 256 * u] : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (tc[u] < T) base[i + 256 * u] = fmaxf(v[u] - ref, floorv);
  }
  for (; i < i1; i += 256) {
    if (t < T) base[i] = fmaxf(base[i] - ref, floorv);
    t += step_t;
    if (t >= a.stride_f) t -= a.stride_f;
  }
}

// z-normalisation of an utterance's whole feature matrix (compute_melspec, ASV_dataset.ipynb:1151):
// pass 1 accumulates sum and sum of squares in double (one atomicAdd pair per CTA), pass 2 applies
// (x - mean) * rsqrt(var).  grid.x = B * n_chunks, a chunk = ZN_CHUNK consecutive elements of the
// row-major (c, t) enumeration of the valid part.
constexpr int ZN_CHUNK = 8192;
struct ZnArgs {
  float* out;
  long long stride_b;
  int stride_c, stride_t;
  const int32_t* nf_eff;
  int C, n_chunks;
  double* stats;  // [B][2]
};
template <bool APPLY>
__global__ void __launch_bounds__(256) k_znorm(const ZnArgs a) {
  const int b = blockIdx.x / a.n_chunks, ch = blockIdx.x - b * a.n_chunks;
  const int T = a.nf_eff[b];
  const long long total = (long long)a.C * T;
  const long long i0 = (long long)ch * ZN_CHUNK;
  if (i0 >= total) return;
  const long long i1 = min(total, i0 + ZN_CHUNK);
  float* ob = a.out + (long long)b * a.stride_b;
  float mean = 0.f, rstd = 0.f;
  if (APPLY) {
    const double n = (double)total, m = a.stats[2 * b] / n;
    const double var = fmax(a.stats[2 * b + 1] / n - m * m, 0.0);
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var));  // std == 0 -> inf/nan, as numpy's division by zero gives
  }
  double s1 = 0.0, s2 = 0.0;
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const int c = (int)(i / T), t = (int)(i - (long long)c * T);
    float* p = ob + (long long)c * a.stride_c + (long long)t * a.stride_t;
    if (APPLY) {
      *p = (*p - mean) * rstd;
    } else {
      const double v = (double)*p;
      s1 += v;
      s2 += v * v;
    }
  }
  if (!APPLY) {
    __shared__ double red[2][8];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = s1;
      red[1][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t1 = 0, t2 = 0;
      for (int w = 0; w < 8; ++w) {
        t1 += red[0][w];
        t2 += red[1][w];
      }
      atomicAdd(a.stats + 2 * b, t1);
      atomicAdd(a.stats + 2 * b + 1, t2);
    }
  }
}

// Column statistics / standardisation over stacked feature rows: sklearn StandardScaler fitted on
// np.vstack(per-utterance arrays) (ASV_dl_func.py:1113-1129, :963-973).  x is [n_rows][row_stride] with
// W valid columns; stats = {sum[W], sumsq[W]} in double (atomicAdd per CTA; all-reduced across ranks by
// the host before the apply pass).  Thread (col = tid % 64, lane row = tid / 64): coalesced row reads.
constexpr int SC_ROWS = 256;  // rows per CTA
// Ragged batches (time-major [B][rows_per_utt][W] with n_frames[b] valid rows): n_frames != null masks the padding
// rows and the rows of utterances with a non-zero status, and the CTA adds its number of valid rows to
// stats[2 W] (np.vstack of the per-utterance arrays stacks valid frames only).
__global__ void __launch_bounds__(256) k_col_stats(const float* x, long long n_rows, int W, long long row_stride,
                                                   double* stats, const int32_t* n_frames, const int32_t* status,
                                                   int rows_per_utt) {
  __shared__ double red[2][4][64];
  const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const long long r0 = (long long)blockIdx.x * SC_ROWS, r1 = min(n_rows, r0 + SC_ROWS);
  auto row_ok = [&](long long r) {
    if (!n_frames) return true;
    const long long b = r / rows_per_utt;
    return (int)(r - b * rows_per_utt) < __ldg(n_frames + b) && (!status || __ldg(status + b) == 0);
  };
  if (n_frames && threadIdx.x == 0) {
    int cnt = 0;
    for (long long r = r0; r < r1; ++r) cnt += row_ok(r) ? 1 : 0;
    atomicAdd(stats + 2 * W, (double)cnt);
  }
  for (int c0 = 0; c0 < W; c0 += 64) {
    const int c = c0 + cl;
    double s1 = 0.0, s2 = 0.0;
    if (c < W)
      for (long long r = r0 + rl; r < r1; r += 4) {
        if (!row_ok(r)) continue;
        const double v = (double)__ldg(x + r * row_stride + c);
        s1 += v;
        s2 += v * v;
      }
    red[0][rl][cl] = s1;
    red[1][rl][cl] = s2;
    __syncthreads();
    if (rl == 0 && c < W) {
      atomicAdd(stats + c, red[0][0][cl] + red[0][1][cl] + red[0][2][cl] + red[0][3][cl]);
      atomicAdd(stats + W + c, red[1][0][cl] + red[1][1][cl] + red[1][2][cl] + red[1][3][cl]);
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) k_col_apply(float* x, long long n_rows, int W, long long row_stride,
                                                   const float* mean, const float* inv_scale) {
  const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const long long r0 = (long long)blockIdx.x * SC_ROWS, r1 = min(n_rows, r0 + SC_ROWS);
  for (int c = cl; c < W; c += 64) {
    const float m = __ldg(mean + c), is = __ldg(inv_scale + c);
    for (long long r = r0 + rl; r < r1; r += 4) {
      float* p = x + r * row_stride + c;
      *p = (*p - m) * is;
    }
  }
}

// mean over frames of feat[b][c][0..T_b): one warp per (b, c)
__global__ void __launch_bounds__(128) k_time_mean(const float* feat, long long stride_b, int stride_c,
                                                   const int32_t* nf_eff, int C, float* out,
                                                   long long out_stride_b) {
  const int b = blockIdx.x;
  const int c = blockIdx.y * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int T = nf_eff[b];
  if (c >= C || T == 0) return;
  const float* row = feat + (long long)b * stride_b + (long long)c * stride_c;
  float s = 0.f;
  for (int t = lane; t < T; t += 32) s += row[t];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[(long long)b * out_stride_b + c] = s / (float)T;
}

// standalone delta: librosa.feature.delta(x, width, order, mode='interp') on [B][C][t_stride]
struct DeltaArgs {
  const float* x;
  float* out;
  const int32_t* n_frames;
  int C, t_stride, width;
  float taps[CEP_MAXW];
};
__global__ void __launch_bounds__(256) k_delta(const DeltaArgs a) {
  const int b = blockIdx.x, c = blockIdx.y;
  const int T = a.n_frames[b];
  const int h = a.width >> 1;
  if (T < a.width) return;
  const long long base = ((long long)b * a.C + c) * a.t_stride;
  for (int t = blockIdx.z * blockDim.x + threadIdx.x; t < T; t += gridDim.z * blockDim.x) {
    int te = min(max(t, h), T - 1 - h);
    float d = 0.f;
    for (int i = 0; i < a.width; ++i) d = __fmaf_rn(a.taps[i], __ldg(a.x + base + te - h + i), d);
    a.out[base + t] = d;
  }
}

// dense FP32 FMA peak: 8 independent chains per thread, 2 flops per FFMA
__global__ void __launch_bounds__(256) k_fma_peak(float* sink, int iters) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
  float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  const float a = 0.999f, c = 1e-4f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = __fmaf_rn(x0, a, c); x1 = __fmaf_rn(x1, a, c); x2 = __fmaf_rn(x2, a, c); x3 = __fmaf_rn(x3, a, c);
      x4 = __fmaf_rn(x4, a, c); x5 = __fmaf_rn(x5, a, c); x6 = __fmaf_rn(x6, a, c); x7 = __fmaf_rn(x7, a, c);
    }
  }
  float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456f) sink[0] = s;
}

#endif  // AAD_STFT_ONLY

}  // namespace aad
