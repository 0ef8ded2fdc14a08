/*
 * aad.h -- C ABI of the B200-native spectral front-end (libaad_b200.so).
 *
 * Drop-in boundary for the feature-extraction hot path of
 * IzaP1k/AudioAnalysisDetector.  The reference has no FFI of its own (it is pure
 * Python); each entry point below cites the reference interface it replaces.
 * Signatures are plain C: pointers, sizes, a CUDA stream passed as void*.  All
 * data pointers are DEVICE pointers unless the name says host.  The device
 * entry points (aad_extract*, aad_logmel / mfcc / lfcc, aad_delta, ...) never
 * allocate (caller owns inputs, outputs, workspace and status) and never
 * synchronise the stream; the host convenience path aad_extract_host owns
 * internal device buffers that grow lazily on first use or up front with
 * aad_host_reserve, and is synchronous.  Nothing throws across the ABI:
 * every call returns 0 or a negative aad_error; per-utterance problems are
 * reported in status[B] (the reference's "print and return None" convention,
 * ASV_dl_func.py:418-420,437-439,536-538).
 */
#ifndef AAD_H_
#define AAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AAD_VERSION 200 /* 0.2.0: aad_params grew (spectrum), GTCC / CQCC / FLAC entry points */

typedef struct aad_plan aad_plan;

/* ---- enums ------------------------------------------------------------- */
enum aad_kind {        /* which reference extractor the plan mirrors */
  AAD_KIND_LOGMEL = 0, /* extract_mel_spectrogram  ASV_dl_func.py:522-538 */
  AAD_KIND_MFCC = 1,   /* extract_mfcc             ASV_dl_func.py:404-420 */
  AAD_KIND_LFCC = 2,   /* extract_lfcc             ASV_dl_func.py:423-439 */
  AAD_KIND_GTCC = 3    /* extract_gtcc (spafe gfcc) ASV_dl_func.py:484-499 */
};
enum aad_dtype { AAD_F32 = 0, AAD_I16 = 1 }; /* AAD_I16: see aad_params.i16_scale */
enum aad_window {
  AAD_WIN_HANN_PERIODIC = 0,    /* scipy get_window('hann', fftbins=True): librosa.stft */
  AAD_WIN_HAMMING_SYMMETRIC = 1 /* np.hamming(win_length): spafe windowing */
};
enum aad_fb {
  AAD_FB_MEL_SLANEY = 0,    /* librosa.filters.mel(htk=False, norm='slaney') */
  AAD_FB_LINEAR_INTBIN = 1, /* triangles on integer FFT bins (spafe 0.1.x / python_speech_features style) */
  AAD_FB_LINEAR_CONT = 2,   /* spafe 0.3.x linear_filter_banks: triangles on continuous bin frequencies; the
                               LFCC default (the reference pins spafe ~= 0.3.3, requirements.txt:5) */
  AAD_FB_CUSTOM = 3,        /* caller matrix, at most two adjacent filters per bin */
  AAD_FB_GAMMATONE = 4,     /* spafe 0.3.x gammatone_filter_banks (Slaney's ERB filters evaluated on the FFT bins, every
                               filter scaled to a maximum of 1): a DENSE matrix; n_fft 512 only, n_filt <= 64 */
  AAD_FB_CUSTOM_DENSE = 5   /* caller matrix without any structure (e.g. spafe's own gammatone / bark matrix);
                               n_fft 512 only, n_filt <= 64 */
};
enum aad_log {
  AAD_LOG_DB10 = 0, /* 10*log10(max(amin, S))            librosa.power_to_db */
  AAD_LOG_LN = 1,   /* ln(S == 0 ? eps : S)              spafe zero_handling + np.log */
  AAD_LOG_CBRT = 2  /* S^(1/3)                           spafe gfcc (dense filter banks only) */
};
enum aad_spectrum {
  AAD_SPEC_POWER = 0,    /* filter bank applied to |X|^2 (times power_scale) */
  AAD_SPEC_MAGNITUDE = 1 /* filter bank applied to |X| (times power_scale); dense filter banks only */
};
enum aad_ref {
  AAD_REF_ONE = 0,    /* power_to_db(S) inside librosa.feature.mfcc */
  AAD_REF_UTT_MAX = 1 /* power_to_db(S, ref=np.max)  ASV_dl_func.py:534 */
};
enum aad_layout {
  AAD_LAYOUT_CT = 0, /* out[b][c][t]  (librosa orientation; cnn_bilstm_hybrid.py:56 wants (B,F,T)) */
  AAD_LAYOUT_TC = 1  /* out[b][t][c]  (spafe orientation; BiLSTM collate ASV_dl_func.py:1206-1227) */
};

enum aad_error {
  AAD_OK = 0,
  AAD_ERR_INVALID_ARG = -1,
  AAD_ERR_UNSUPPORTED = -2, /* e.g. n_fft not in {256,512,1024,2048} */
  AAD_ERR_CUDA = -3,
  AAD_ERR_WORKSPACE = -4, /* workspace too small */
  AAD_ERR_KIND = -5,      /* plan kind does not match the entry point */
  AAD_ERR_FILTERBANK = -6, /* custom filterbank is not two-adjacent-filters-per-bin */
  AAD_ERR_PAIR = -7,       /* aad_extract_pair: the two plans do not share one STFT, or the second is not a plain log filter bank */
  AAD_ERR_FORMAT = -8      /* aad_flac_*: not a FLAC stream, or a frame fails its CRC / uses a reserved code */
};

enum aad_item_status { /* status[b] */
  AAD_ITEM_OK = 0,
  AAD_ITEM_EMPTY = 1,               /* length <= 0 */
  AAD_ITEM_TOO_SHORT_FOR_FRAME = 2, /* spafe framing: L < win_length */
  AAD_ITEM_TOO_SHORT_FOR_DELTA = 3, /* librosa.feature.delta: T < width */
  AAD_ITEM_OUT_TOO_SMALL = 4,       /* T > T_alloc of the output buffer */
  AAD_ITEM_NONFINITE = 5            /* NaN/Inf reached the log-energies (librosa.util.valid_audio) */
};

/* ---- plan parameters (POD) ---------------------------------------------- */
typedef struct aad_params {
  int32_t struct_size; /* sizeof(aad_params), for ABI versioning */
  int32_t kind;        /* aad_kind */
  int32_t sample_rate;
  int32_t n_fft;      /* 256 | 512 | 1024 | 2048 */
  int32_t win_length; /* <= n_fft */
  int32_t hop_length;
  int32_t window; /* aad_window */
  int32_t center; /* 1: librosa.stft(center=True, pad_mode='constant'): zero-pad n_fft/2
                        both sides, T = 1 + L/hop, window centred in the n_fft buffer.
                     0: spafe framing: no padding, T = (L-win)/hop + 1, window at the
                        start of the buffer, zero-extended to n_fft. */
  int32_t quantize_i16; /* float input only: y -> (int16)trunc(y*32767)  ASV_dl_func.py:434 */
  float pre_emph;       /* 0 = off; spafe default 0.97, first sample unchanged */
  int32_t n_filt;       /* n_mels / nfilts */
  int32_t fb_type;      /* aad_fb */
  float fmin;
  float fmax;         /* <= 0: sample_rate / 2 */
  float power_scale;  /* power spectrum multiplier folded into the filterbank (spafe: 1/nfft); 0 -> 1 */
  int32_t log_type;   /* aad_log */
  int32_t ref_type;   /* aad_ref (dB only) */
  float amin;         /* 1e-10 */
  float top_db;       /* 80; < 0 disables the floor */
  int32_t n_ceps;     /* 0: no DCT (log-mel); else DCT-II ortho, first n_ceps (scipy.fftpack.dct) */
  int32_t n_delta;    /* 0 | 1 | 2: append delta (and delta-delta) rows: librosa.feature.delta */
  int32_t delta_width; /* odd, 3..9 (9 = librosa default) */
  int32_t layout;      /* aad_layout */
  int32_t time_mean;   /* 1: out[b][c] = mean over frames (the reference's mean=True, axis=1 of (C,T)) */
  float i16_scale;     /* int16 input only: sample value = int16 * i16_scale.  0 -> default: 1/32768 for
                          LOGMEL / MFCC (16-bit PCM as librosa.load / soundfile decode it to float32),
                          1 for LFCC (spafe receives the raw int16 array, ASV_dl_func.py:434-435) */
  int32_t znorm;       /* 1: out = (x - mean(x)) / std(x) over all valid elements of the utterance (population
                          std), the compute_melspec variant of the reference (ASV_dataset.ipynb:1151,
                          cell [27]); applied last; not combinable with time_mean */
  const float* custom_fb; /* HOST pointer, row-major n_filt x (n_fft/2+1); AAD_FB_CUSTOM / AAD_FB_CUSTOM_DENSE only */
  int32_t spectrum;       /* aad_spectrum */
  int32_t reserved0;
} aad_params;

/* Fill *p with the parameters of the named reference extractor:
 *   AAD_KIND_LOGMEL: melspectrogram(n_fft 2048, hop 512, hann, center, n_mels 64, fmax sr/2)
 *                    + power_to_db(ref=np.max)                    ASV_dl_func.py:533-534
 *   AAD_KIND_MFCC:   librosa.feature.mfcc(n_mfcc 13; 128 mels)    ASV_dl_func.py:416
 *   AAD_KIND_LFCC:   int16 quantise + spafe lfcc(num_ceps 13, 24 filters, nfft 512,
 *                    25 ms / 10 ms hamming, pre-emphasis 0.97), TC layout  ASV_dl_func.py:434-435
 *   AAD_KIND_GTCC:   spafe gfcc(num_ceps 13, 40 gammatone filters as the reference passes them, nfft 512,
 *                    25 ms / 10 ms hamming, pre-emphasis 0.97, power spectrum / nfft, cube root, DCT-II ortho),
 *                    float input as librosa.load returns it, TC layout              ASV_dl_func.py:484-499 */
int aad_params_default(aad_params* p, int kind, int sample_rate);

/* ---- plan life cycle ------------------------------------------------------
 * Builds the device tables (window, twiddles, banded filterbank, DCT matrix,
 * Savitzky-Golay taps) once.  Replaces the per-call table construction inside
 * librosa.filters.mel / get_window / spafe linear_filter_banks. */
int aad_plan_create(const aad_params* p, int device, aad_plan** out);
int aad_plan_destroy(aad_plan* plan);

/* Output geometry and workspace size for a batch of B utterances of at most
 * max_len samples.  *t_max = frames of a max_len utterance; *c_out = rows per
 * frame (n_filt or n_ceps, times 1 + n_delta). */
int aad_query(const aad_plan* plan, int B, int64_t max_len, int32_t* t_max, int32_t* c_out,
              size_t* workspace_bytes);

/* ---- the hot call -----------------------------------------------------------
 * Batched replacement for one pass of extract_features' inner loop
 * (ASV_dl_func.py:1036-1045) over B in-memory utterances.
 *   wav        [B][wav_stride] float32 or int16 (wav_dtype), row b valid for lengths[b] samples
 *   lengths    [B] int32 (device)
 *   out        CT: [B][c_out][t_alloc]; TC: [B][t_alloc][c_out]; time_mean: [B][c_out]
 *              out_stride_b in elements (0 = dense)
 *   n_frames   [B] int32 out: frames of utterance b (the reference's T)
 *   status     [B] int32 out: aad_item_status; rows with non-zero status are left untouched
 *   workspace  device scratch of at least aad_query's workspace_bytes
 *   stream     cudaStream_t; work is enqueued, never synchronised
 */
int aad_extract(const aad_plan* plan, const void* wav, int wav_dtype, int64_t wav_stride,
                const int32_t* lengths, int B, int64_t max_len, float* out, int64_t out_stride_b,
                int32_t t_alloc, int32_t* n_frames, int32_t* status, void* workspace,
                size_t workspace_bytes, void* stream);

/* The same call over CHUNKS of decoded files that already sit in device memory: utterance b is the
 * lengths[b] samples starting at element row_off[b] of wav (int64, device); chunks may overlap and need
 * no padding.  Replaces the reference's per-chunk `librosa.load(filepath)` + `y[start_sample:end_sample]`
 * (ASV_dl_func.py:406-411, 425-429, 524-528: every 2-second chunk of prepare_dataframe :287-293 decodes
 * its whole file again) by one decode + one upload per file and a chunk table.  The caller guarantees
 * 0 <= row_off[b] and row_off[b] + lengths[b] <= elements of wav; chunk starts on even element offsets
 * (8-byte aligned float32 / 4-byte aligned int16 addresses) take the vector-load path, others the
 * per-sample path with identical results. */
int aad_extract_indexed(const aad_plan* plan, const void* wav, int wav_dtype, const int64_t* row_off,
                        const int32_t* lengths, int B, int64_t max_len, float* out, int64_t out_stride_b,
                        int32_t t_alloc, int32_t* n_frames, int32_t* status, void* workspace,
                        size_t workspace_bytes, void* stream);

/* One STFT, two features: `plan2` (a plain log filter-bank plan: no DCT, deltas, time mean or z-norm, CT layout,
 * e.g. the 64-mel log-mel of extract_mel_spectrogram) runs its filter bank on the power spectra that `plan`
 * (any plan over the same n_fft / hop / window / centring / input conversion, e.g. the MFCC of extract_mfcc)
 * computes in the same kernel launch.  The reference's extractor map runs librosa.stft once per feature and
 * chunk (ASV_deep_learning.ipynb:152-160 -> ASV_dl_func.py:416,533); here the second feature costs one more
 * pass over the power tile in shared memory.  row_off: null (rows of wav_stride) or a chunk table as in
 * aad_extract_indexed.  out2: [B][n_filt2][t_alloc]; workspace2: aad_query(plan2, ...) bytes; n_frames and
 * status are shared.  Returns AAD_ERR_PAIR when the plans cannot be paired. */
int aad_extract_pair(const aad_plan* plan, const aad_plan* plan2, const void* wav, int wav_dtype, int64_t wav_stride,
                     const int64_t* row_off, const int32_t* lengths, int B, int64_t max_len, float* out,
                     int64_t out_stride_b, float* out2, int64_t out2_stride_b, int32_t t_alloc, int32_t* n_frames,
                     int32_t* status, void* workspace, size_t workspace_bytes, void* workspace2,
                     size_t workspace2_bytes, void* stream);

/* Kind-checked aliases of aad_extract (return AAD_ERR_KIND on mismatch). */
int aad_logmel(const aad_plan* plan, const void* wav, int wav_dtype, int64_t wav_stride,
               const int32_t* lengths, int B, int64_t max_len, float* out, int64_t out_stride_b,
               int32_t t_alloc, int32_t* n_frames, int32_t* status, void* workspace,
               size_t workspace_bytes, void* stream);
int aad_mfcc(const aad_plan* plan, const void* wav, int wav_dtype, int64_t wav_stride,
             const int32_t* lengths, int B, int64_t max_len, float* out, int64_t out_stride_b,
             int32_t t_alloc, int32_t* n_frames, int32_t* status, void* workspace,
             size_t workspace_bytes, void* stream);
int aad_lfcc(const aad_plan* plan, const void* wav, int wav_dtype, int64_t wav_stride,
             const int32_t* lengths, int B, int64_t max_len, float* out, int64_t out_stride_b,
             int32_t t_alloc, int32_t* n_frames, int32_t* status, void* workspace,
             size_t workspace_bytes, void* stream);
/* extract_gtcc: spafe gfcc over B in-memory clips (ASV_dl_func.py:484-499) */
int aad_gtcc(const aad_plan* plan, const void* wav, int wav_dtype, int64_t wav_stride,
             const int32_t* lengths, int B, int64_t max_len, float* out, int64_t out_stride_b,
             int32_t t_alloc, int32_t* n_frames, int32_t* status, void* workspace,
             size_t workspace_bytes, void* stream);

/* Standalone delta stencil: librosa.feature.delta(x, width, order, axis=-1, mode='interp')
 * on x[B][C][t_stride] with per-utterance n_frames (edges replicate at each utterance's
 * own T).  Rows with n_frames[b] < width are left untouched. */
int aad_delta(const float* x, const int32_t* n_frames, int B, int C, int32_t t_stride,
              int width, int order, float* out, void* stream);

/* librosa.power_to_db's reference / floor (ASV_dl_func.py:534) as a separate step, for an utterance whose
 * frames were computed in several pieces (time-axis split of long-form audio across GPUs):
 * every piece is extracted with ref_type = AAD_REF_ONE and top_db < 0 (raw 10 log10(max(amin, S))), the
 * caller reduces the per-utterance maximum over the pieces (one MAX all-reduce of B floats), then
 *   x[b][f][t] = max(x - ref, (utt_max[b] - ref) - top_db),  ref = utt_max[b] (AAD_REF_UTT_MAX) or 0,
 * in place on x[B][n_filt][stride_f] (CT layout), frames t < n_frames[b].  top_db < 0: no floor. */
int aad_db_reference(float* x, int64_t stride_b, int32_t stride_f, const int32_t* n_frames, const float* utt_max,
                     int B, int n_filt, int32_t t_max, int ref_type, float top_db, void* stream);

/* Feature standardisation, the step right after the path in the reference: sklearn StandardScaler
 * fitted on np.vstack(per-utterance feature arrays) and applied per utterance
 * (prepare_train_test_data ASV_dl_func.py:1113-1129, train_all_features :963-973).
 *   aad_scaler_accumulate: stats[2*W] (device, double; zero it first) += {column sums, column sums of
 *       squares} of x[n_rows][row_stride] (W valid columns).  Ranks all-reduce stats (SUM) before the
 *       host turns them into mean / 1/scale (population variance; scale 1 where the variance is 0).
 *   aad_scaler_apply: x = (x - mean[c]) * inv_scale[c] in place. */
int aad_scaler_accumulate(const float* x, int64_t n_rows, int32_t W, int64_t row_stride, double* stats,
                          void* stream);
/* The same for a ragged time-major batch x[B][rows_per_utt][row_stride] straight from the extractor (TC
 * layout): only the n_frames[b] valid rows of utterances with status[b] == 0 (status may be NULL) enter the
 * sums, as np.vstack of the per-utterance arrays does, and the number of rows that did is added to
 * stats[2*W] (stats holds 2*W + 1 doubles here; zero it first). */
int aad_scaler_accumulate_ragged(const float* x, int B, int32_t rows_per_utt, int32_t W, int64_t row_stride,
                                 const int32_t* n_frames, const int32_t* status, double* stats, void* stream);
int aad_scaler_apply(float* x, int64_t n_rows, int32_t W, int64_t row_stride, const float* mean,
                     const float* inv_scale, void* stream);

/* CQCC, the sibling extractor the reference's CNN-BiLSTM is trained on (extract_cqcc, ASV_dl_func.py:442-481;
 * SURVEY 8f row 4): librosa.cqt (hop 512, fmin = C1, n_bins = floor(log2((sr/2 - 100) / fmin)) * bins_per_octave)
 * -> |.| -> amplitude_to_db(ref=np.max) -> per-frame linear interpolation onto a uniform frequency grid
 * -> log(x^2 + 1e-12) -> DCT-II ortho, first n_ceps rows.  out is [B][n_ceps][t_alloc] (T = 1 + len / 512 frames
 * per utterance); status as in aad_extract (1 empty, 4 output too small, 5 non-finite audio).  row_off (optional,
 * [B] element offsets): utterance b starts at wav + row_off[b] instead of b * wav_stride -- chunks of decoded files,
 * as in aad_extract_indexed.  cqt_mag_out
 * (optional, [B][n_bins][t_alloc]) receives |CQT|.  The octave recursion of librosa.cqt is followed; its
 * 'soxr_hq' resampler is replaced by a documented half-band FIR (csrc/aad_cqcc.cu, oracle/cqcc_ref.py). */
typedef struct aad_cqcc_plan aad_cqcc_plan;
int aad_cqcc_plan_create(int sample_rate, int bins_per_octave, int n_ceps, int device, aad_cqcc_plan** out);
int aad_cqcc_plan_destroy(aad_cqcc_plan* plan);
int aad_cqcc_query(const aad_cqcc_plan* plan, int B, int64_t max_len, int32_t* t_max, int32_t* n_ceps, int32_t* n_bins,
                   size_t* workspace_bytes);
int aad_cqcc(const aad_cqcc_plan* plan, const void* wav, int wav_dtype, int64_t wav_stride, const int64_t* row_off,
             const int32_t* lengths, int B, int64_t max_len, float* out, int64_t out_stride_b, int32_t t_alloc,
             int32_t* n_frames, int32_t* status, float* cqt_mag_out, void* workspace, size_t workspace_bytes, void* stream);

/* Compressed-audio decode on the input side of the path (HOST pointers, CPU code).  The
 * reference's corpus (ASVspoof 2019 / 2021) is 16-bit FLAC read through libsndfile: soundfile.info for the chunk
 * index (ASV_dl_func.py:280), librosa.load in every extractor call (ASV_dl_func.py:406,425,524).
 *   aad_flac_info:   STREAMINFO of a FLAC stream held in memory (total_samples is per channel; md5 = MD5 of the
 *                    interleaved little-endian PCM as the encoder saw it, all zero when absent).
 *   aad_flac_decode: the whole stream as interleaved int32 samples out[n][channels] (sample / 2^(bits-1) is the
 *                    float librosa.load returns before its mono mix); capacity_samples >= total_samples.  Every
 *                    frame is checked against its CRC-8 / CRC-16. */
typedef struct aad_flac_info_t {
  int32_t sample_rate, channels, bits_per_sample;
  int64_t total_samples;
  uint8_t md5[16];
} aad_flac_info_t;
int aad_flac_info(const uint8_t* data, size_t size, aad_flac_info_t* info);
int aad_flac_decode(const uint8_t* data, size_t size, int32_t* out, int64_t capacity_samples, int64_t* n_decoded);
/* The corpus's own format in one call: a MONO 16-BIT stream straight to int16 PCM (AAD_ERR_UNSUPPORTED otherwise), with the
 * MD5 of STREAMINFO checked in the library: *md5_state = 1 verified, 0 the stream stores no checksum, -1 mismatch. */
int aad_flac_decode_pcm16(const uint8_t* data, size_t size, int16_t* out, int64_t capacity_samples, int64_t* n_decoded,
                          int32_t* md5_state);

/* Host-buffer convenience path (what the reference-facing Python drop-ins use for
 * host arrays): pinned-or-pageable HOST wav/lengths in, HOST out/n_frames/status back,
 * chunked H2D -> kernels -> D2H pipelined on internal streams.  Synchronous.  Its internal
 * buffers (three chunk-sized device sets + pinned staging for the per-utterance arrays) are
 * allocated on the first call with a given shape, or ahead of time by aad_host_reserve (same
 * wav_dtype / B / max_len / t_alloc / chunk_utts as the calls that follow); later calls with
 * the same or smaller shapes allocate nothing.  aad_host_alloc / aad_host_free hand out pinned
 * host memory for the caller's buffers (write_combined != 0: faster for the device to read,
 * slow for the CPU to read back -- input staging only, never the output). */
int aad_host_reserve(aad_plan* plan, int wav_dtype, int B, int64_t max_len, int32_t t_alloc, int chunk_utts);
int aad_host_alloc(void** ptr, size_t bytes, int write_combined);
int aad_host_free(void* ptr);
int aad_extract_host(aad_plan* plan, const void* wav_host, int wav_dtype, int64_t wav_stride,
                     const int32_t* lengths_host, int B, int64_t max_len, float* out_host,
                     int64_t out_stride_b, int32_t t_alloc, int32_t* n_frames_host,
                     int32_t* status_host, int chunk_utts);

/* ---- the consumer of the features ------------------------------------------------------
 * Inference of the reference's AudioDeepfakeDetector (cnn_bilstm_hybrid.py:20-68, eval mode: dropout off,
 * BatchNorm on running statistics) on the front-end's output where it lies in device memory:
 * Conv1d(63 -> 64, k 3) over the F feature rows with the 63 frames as channels + BatchNorm + ReLU +
 * MaxPool(2) -> BiLSTM(64 -> 2 x 32) -> LayerNorm(1)-weighted max over time -> Linear(64, 64) + ReLU ->
 * Linear(64, 1) + Sigmoid.  Replaces `model(x)` at cnn_bilstm_hybrid.py:54-68 and the per-item
 * torch.tensor(...) collation of CQCCDataset (:4-15).  Weights: HOST pointers in the layouts of the
 * reference's state dict (conv_w [64][63][3]; w_ih [128][64], w_hh [128][32], gate order i, f, g, o; *_r the
 * reverse direction; fc1_w [64][64], fc2_w [1][64]).  attn_w / attn_b / ln_w are accepted for completeness:
 * LayerNorm over one element returns ln_b for every finite input, so they cannot influence the output. */
typedef struct aad_detector aad_detector;
typedef struct aad_detector_weights {
  int32_t struct_size;   /* sizeof(aad_detector_weights) */
  int32_t feature_dim;   /* F: 13 (MFCC), 19 (CQCC), 64 (log-mel) ...; pooled sequence length F / 2 */
  const float *conv_w, *conv_b, *bn_w, *bn_b, *bn_mean, *bn_var;
  float bn_eps;          /* 1e-5 */
  const float *w_ih, *w_hh, *b_ih, *b_hh, *w_ih_r, *w_hh_r, *b_ih_r, *b_hh_r;
  const float *attn_w, *attn_b, *ln_w, *ln_b;
  const float *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} aad_detector_weights;
int aad_detector_create(const aad_detector_weights* w, int device, aad_detector** out);
int aad_detector_destroy(aad_detector* det);
int aad_detector_query(const aad_detector* det, int B, size_t* workspace_bytes);
/* feats [B][F][stride_f >= 63] float32 (CT layout, frames 0..62 of every row are read), stride_b in elements
 * (0 = dense); scores [B] float32 out; enqueued on `stream`, never synchronised. */
int aad_detector_forward(const aad_detector* det, const float* feats, int64_t stride_b, int32_t stride_f, int B,
                         float* scores, void* workspace, size_t workspace_bytes, void* stream);

/* ---- introspection (tests / parity) ---------------------------------------- */
enum aad_table { AAD_TABLE_WINDOW = 0, AAD_TABLE_FILTERBANK = 1, AAD_TABLE_DCT = 2, AAD_TABLE_DELTA_TAPS = 3 };
/* Copies a plan table to HOST memory as float32: WINDOW [n_fft] (unscaled, zero-extended),
 * FILTERBANK dense [n_filt][n_fft/2+1] (power_scale folded in), DCT [n_ceps][n_filt],
 * DELTA_TAPS [2][delta_width].  Returns the element count or a negative error. */
int64_t aad_plan_table(const aad_plan* plan, int which, float* host_out, int64_t capacity);

/* Number of kernel launches one aad_extract call enqueues for this plan. */
int aad_plan_launches(const aad_plan* plan);

/* Per-kernel device timing for the roofline report.  When enabled, aad_extract records CUDA
 * events on the caller's stream around each launch; after the caller has synchronised,
 * aad_plan_kernel_times fills ms_out[4] = {prepare, stft_fb, cepstra|finalize, time_mean}
 * for the most recent call.  Off by default (the timed bench loop runs without it). */
int aad_plan_set_profiling(aad_plan* plan, int enable);
int aad_plan_kernel_times(const aad_plan* plan, float* ms_out);

/* Dense FP32 FMA micro-benchmark (roofline denominator; not in MEASURED_PEAKS.json):
 * runs `iters` dependent-chain FFMA blocks on `device`, returns achieved TFLOP/s. */
int aad_fp32_peak(int device, int iters, double* tflops_out);

const char* aad_strerror(int err);
/* Text of the most recent CUDA failure seen by this thread inside the library ("" if none). */
const char* aad_last_error_detail(void);
int aad_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AAD_H_ */
