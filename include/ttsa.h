/*
 * ttsa.h -- C ABI of the B200-native spectrogram-domain audio hot path
 *           ("TTS audio": STFT / iSTFT / mel / Griffin-Lim), libttsa_b200.so.
 *
 * Drop-in boundary for the reference's utils/audio.py::AudioProcessor
 * (prototypefund/your-voice-TTS).  Every entry point below names the reference
 * function (file:line under the reference tree) whose arithmetic it replaces.
 * The reference is pure Python (numpy + librosa + scipy); a maintainer binds this
 * library with ctypes and keeps AudioProcessor's Python
 * signature unchanged.
 *
 * Conventions
 *   - plain C: opaque handles, POD structs, raw pointers and sizes; no C++/torch types.
 *   - "dev" pointers are device (HBM) pointers on the plan's device, caller-owned;
 *     the library never allocates or frees caller memory inside a work call.
 *   - Work calls are asynchronous and ordered on `stream` (a cudaStream_t passed as void*;
 *     NULL = legacy default stream).  They launch only this library's own kernels.
 *   - Spectrogram-domain tensors are FRAME-MAJOR and packed over the batch:
 *       linear  [sum_T, num_freq]   fp32   (row = frame_off[u] + t)
 *       mel     [sum_T, num_mels]   fp32
 *       complex [sum_T, num_freq, 2] fp32 (re, im)
 *     (the acoustic models emit [B, T, D]; the reference's [D, T] numpy layout only exists
 *     because its callers transpose -- utils/synthesis.py:65, train.py:227).
 *   - Waveforms are packed fp32 [total_samples]; utterance u starts at wav_off[u]
 *     (multiple of 4 samples) and holds wav_len[u] samples.
 *   - Every function returns 0 on success or a negative ttsa_status; ttsa_last_error()
 *     gives a thread-local message.  No exception crosses the ABI.
 *   - num_freq == 1025 (n_fft == 2048), the value of every shipped config (config*.json:"num_freq"), runs the tuned
 *     warp-per-frame kernels (hop <= win <= 9*hop, win <= 2048).  Any other power-of-two n_fft in [256, 4096]
 *     (num_freq 129 ... 2049, hop <= win <= n_fft) runs an untuned one-CTA-per-frame path with the same semantics.
 *     Anything else -> TTSA_ERR_UNSUPPORTED.  There is no CPU fallback anywhere in this library.
 */
#ifndef TTSA_H_
#define TTSA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTSA_VERSION 100

typedef enum ttsa_status {
  TTSA_OK = 0,
  TTSA_ERR_BAD_ARG = -1,
  TTSA_ERR_BAD_CONFIG = -2,      /* e.g. mel_fmax > sample_rate/2 (assert at utils/audio.py:70-71) */
  TTSA_ERR_UNSUPPORTED = -3,
  TTSA_ERR_WORKSPACE = -4,       /* workspace too small */
  TTSA_ERR_CUDA = -5,
  TTSA_ERR_NO_DEVICE = -6
} ttsa_status;

/* POD mirror of the "audio" block of config.json:5-25 plus the derived STFT parameters of
 * utils/audio.py:114-119 (the caller derives n_fft/hop/win exactly like _stft_parameters). */
typedef struct ttsa_config {
  int32_t sample_rate;
  int32_t num_mels;
  int32_t num_freq;
  int32_t n_fft;            /* (num_freq - 1) * 2 */
  int32_t hop_length;       /* int(frame_shift_ms / 1000 * sample_rate) */
  int32_t win_length;       /* int(frame_length_ms / 1000 * sample_rate) */
  int32_t signal_norm;      /* bool */
  int32_t symmetric_norm;   /* bool */
  int32_t clip_norm;        /* bool */
  int32_t griffin_lim_iters;
  double min_level_db;
  double ref_level_db;
  double power;
  double preemphasis;
  double max_norm;
  double mel_fmin;
  double mel_fmax;          /* <= 0 means None (sample_rate / 2) */
} ttsa_config;

typedef struct ttsa_plan ttsa_plan;    /* immutable per (config, device): window, twiddles, mel basis, pinv */
typedef struct ttsa_batch ttsa_batch;  /* immutable per batch shape: per-utterance frame counts / offsets    */

/* ---- library ---------------------------------------------------------------------------- */
int ttsa_version(void);
const char* ttsa_last_error(void);
/* Number of kernel launches this library has issued in this process (bench.py "gpu_launches"). */
uint64_t ttsa_launch_count(void);

/* ---- plan ------------------------------------------------------------------------------- */
/* AudioProcessor.__init__ + _stft_parameters + _build_mel_basis (+ pinv)  utils/audio.py:12-77,114-119 */
int ttsa_plan_create(const ttsa_config* cfg, int device, ttsa_plan** out);
int ttsa_plan_destroy(ttsa_plan* plan);
/* _build_mel_basis (utils/audio.py:68-77): float64 [num_mels, num_freq] into a HOST buffer. */
int ttsa_plan_mel_basis(const ttsa_plan* plan, double* host_out);
/* np.linalg.pinv(_build_mel_basis()) (utils/audio.py:65): float64 [num_freq, num_mels], HOST buffer. */
int ttsa_plan_inv_mel_basis(const ttsa_plan* plan, double* host_out);
/* The schedule by which the feature kernel contracts |X| with _build_mel_basis() (_linear_to_mel inside
 * melspectrogram, utils/audio.py:60-62,145-152), for inspection and tests; works on host-only plans.
 * pairs_out[3]: step pairs of the three slots (all 0: this basis is served by the per-filter lane schedule);
 * words_out (HOST, may be NULL): float4 [NP][32] | u32 [NP][32] | u32 [3][32] as described in
 * csrc/host_tables.hpp (mel_segment_schedule).  Returns the number of 32-bit words (<= cap_words when
 * words_out is given) or a negative ttsa_status. */
int64_t ttsa_plan_mel_schedule(const ttsa_plan* plan, int32_t* pairs_out, uint32_t* words_out, int64_t cap_words);

/* ---- batch layout ----------------------------------------------------------------------- */
/* From per-utterance frame counts T[u] (spectrogram-domain inputs; the waveform of utterance u then
 * has hop*(T[u]-1) samples, the length librosa.istft returns -- utils/audio.py:199-201). */
int ttsa_batch_from_frames(const ttsa_plan* plan, const int32_t* n_frames_host, int32_t n_utts, ttsa_batch** out);
/* From per-utterance waveform lengths L[u] (wav inputs; T[u] = 1 + L[u] / hop as librosa.stft with
 * center=True yields -- utils/audio.py:191-197). */
int ttsa_batch_from_wav_lengths(const ttsa_plan* plan, const int32_t* wav_len_host, int32_t n_utts, ttsa_batch** out);
/* Strided variants: utterance u's frames live at rows [u*frame_stride, u*frame_stride + T[u]) of the spectrogram
 * tensors, i.e. the tensors are padded [n_utts, frame_stride, D] blocks as the acoustic models emit them
 * (models/tacotron2.py:62-73) and as collate_fn builds them (datasets/TTSDataset.py:209-217, utils/data.py:25-31).
 * Rows beyond T[u] are never read or written.  frame_stride >= max T[u]. */
int ttsa_batch_from_frames_strided(const ttsa_plan* plan, const int32_t* n_frames_host, int32_t n_utts,
                                   int64_t frame_stride, ttsa_batch** out);
int ttsa_batch_from_wav_lengths_strided(const ttsa_plan* plan, const int32_t* wav_len_host, int32_t n_utts,
                                        int64_t frame_stride, ttsa_batch** out);
int ttsa_batch_destroy(ttsa_batch* batch);
int64_t ttsa_batch_total_frames(const ttsa_batch* batch);   /* sum_T: rows of the packed spectrogram tensors */
int64_t ttsa_batch_total_samples(const ttsa_batch* batch);  /* floats in the packed waveform buffer          */
/* Host copies of the offsets: frame_off[n_utts+1], wav_off[n_utts+1], wav_len[n_utts] (any may be NULL). */
int ttsa_batch_offsets(const ttsa_batch* batch, int64_t* frame_off, int64_t* wav_off, int32_t* wav_len);

/* ---- forward: wav -> features ----------------------------------------------------------- */
#define TTSA_FEAT_PREEMPHASIS 1   /* apply_preemphasis before the STFT (utils/audio.py:139-140) */
/* spectrogram() and melspectrogram() in ONE pass over the STFT (utils/audio.py:138-152; the reference
 * computes the STFT twice per wav, datasets/TTSDataset.py:191-192).  lin_out [sum_T,num_freq] and/or
 * mel_out [sum_T,num_mels] receive _normalize(_amp_to_db(.) - ref_level_db); either may be NULL. */
int ttsa_stft_features(const ttsa_plan* plan, const ttsa_batch* batch, const float* wav_dev,
                       float* lin_out_dev, float* mel_out_dev, uint32_t flags, void* stream);
/* _stft (utils/audio.py:191-197): complex [sum_T, num_freq, 2]; no pre-emphasis, no dB. */
int ttsa_stft(const ttsa_plan* plan, const ttsa_batch* batch, const float* wav_dev,
              float* stft_out_dev, void* stream);

/* ---- inverse: spectrogram -> wav -------------------------------------------------------- */
/* _istft (utils/audio.py:199-201): complex [sum_T,num_freq,2] -> packed wav (hop*(T-1) samples each). */
int ttsa_istft(const ttsa_plan* plan, const ttsa_batch* batch, const float* stft_dev,
               float* wav_out_dev, void* stream);

#define TTSA_SPEC_MAGNITUDE 0   /* input is S, used as |S|           (_griffin_lim(S), utils/audio.py:182) */
#define TTSA_SPEC_NORM_DB   1   /* input is the normalised dB spectrogram; _denormalize, +ref_level_db,
                                   _db_to_amp and **power are fused  (inv_spectrogram, utils/audio.py:154-160) */
#define TTSA_GL_DEEMPHASIS  1   /* apply_inv_preemphasis on the result (utils/audio.py:160, 170) */
/* _griffin_lim (utils/audio.py:182-189) over a packed batch: iters STFTs + (iters+1) iSTFTs.
 *   spec_dev      [sum_T, num_freq] fp32, meaning given by spec_kind
 *   init_angles   [sum_T, num_freq] fp32 radians, or NULL -> 2*pi*U[0,1) from a counter RNG keyed by seed
 *                 (the reference draws np.random.rand, utils/audio.py:183)
 *   sc_log_dev    NULL or [iters, n_utts, 2] fp32: (sum (|stft(y_{i-1})| - S)^2, sum S^2) per iteration/utterance
 *   workspace     >= ttsa_griffin_lim_workspace_bytes() bytes, 256-byte aligned, device memory
 *   wav_out_dev must be 16-byte aligned (TTSA_ERR_BAD_ARG otherwise): waveform spans are read as sample pairs from
 *   16-byte-aligned utterance slots; spectrogram rows may start at any 4-byte boundary */
size_t ttsa_griffin_lim_workspace_bytes(const ttsa_plan* plan, const ttsa_batch* batch);
int ttsa_griffin_lim(const ttsa_plan* plan, const ttsa_batch* batch, const float* spec_dev, int spec_kind,
                     int iters, const float* init_angles_dev, uint64_t seed, uint32_t flags,
                     float* wav_out_dev, float* sc_log_dev, void* workspace_dev, size_t workspace_bytes,
                     void* stream);

/* Fast Griffin-Lim (Perraudin, Balazs, Sondergaard 2013; librosa >= 0.7 griffinlim(momentum=...)).  NOT in the
 * reference -- it changes the result, so it is a separate, opt-in entry point: iteration i takes the phases of
 * stft(y_{i-1}) - momentum / (1 + momentum) * stft(y_{i-2}) (y_{-1} = 0), i.e. the kernel transforms
 * y_{i-1} - beta * y_{i-2} (the STFT is linear).  momentum in [0, 1); 0 is ttsa_griffin_lim exactly. */
int ttsa_griffin_lim_fast(const ttsa_plan* plan, const ttsa_batch* batch, const float* spec_dev, int spec_kind,
                          int iters, const float* init_angles_dev, uint64_t seed, uint32_t flags, double momentum,
                          float* wav_out_dev, float* sc_log_dev, void* workspace_dev, size_t workspace_bytes,
                          void* stream);

/* ---- mel <-> linear (the dense contraction) --------------------------------------------- */
#define TTSA_MEL_IN_AMPLITUDE 0  /* input already amplitude                                            */
#define TTSA_MEL_IN_NORM_DB   1  /* input is normalised dB: _denormalize, +ref, _db_to_amp fused on load */
#define TTSA_MEL_OUT_PLAIN    0  /* plain GEMM result (after max(1e-10,.) for mel_to_linear)             */
#define TTSA_MEL_OUT_POWER    1  /* mel_to_linear only: additionally ** power (inv_mel_spectrogram:170)  */
#define TTSA_MEL_OUT_NORM_DB  2  /* linear_to_mel only: _normalize(_amp_to_db(.) - ref) (out_linear_to_mel) */
/* _mel_to_linear: max(1e-10, pinv(mel_basis) @ mel)  (utils/audio.py:64-66); mel [sum_T,num_mels] ->
 * lin [sum_T,num_freq]. */
int ttsa_mel_to_linear(const ttsa_plan* plan, const ttsa_batch* batch, const float* mel_dev, int in_kind,
                       float* lin_out_dev, int out_kind, void* stream);
/* _linear_to_mel: mel_basis @ S (utils/audio.py:60-62) and out_linear_to_mel (utils/audio.py:174-180). */
int ttsa_linear_to_mel(const ttsa_plan* plan, const ttsa_batch* batch, const float* lin_dev, int in_kind,
                       float* mel_out_dev, int out_kind, void* stream);

/* ---- pre/de-emphasis -------------------------------------------------------------------- */
/* apply_preemphasis: y[n] = x[n] - p x[n-1]  (utils/audio.py:128-131); packed wav in/out (may not alias). */
int ttsa_preemphasis(const ttsa_plan* plan, const ttsa_batch* batch, const float* x_dev, float* y_dev, void* stream);
/* apply_inv_preemphasis: y[n] = x[n] + p y[n-1]  (utils/audio.py:133-136); blocked linear-recurrence scan.
 * workspace >= ttsa_deemphasis_workspace_bytes(); x and y may alias. */
size_t ttsa_deemphasis_workspace_bytes(const ttsa_plan* plan, const ttsa_batch* batch);
int ttsa_deemphasis(const ttsa_plan* plan, const ttsa_batch* batch, const float* x_dev, float* y_dev,
                    void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- waveform post-processing (the steps right after the synthesis path) ------------------ */
/* max |wav_u| per utterance (the peak save_wav normalises by, utils/audio.py:57).  lens_dev: optional [B] int32
 * length override (clamped to the batch's lengths), e.g. the endpoints below. */
int ttsa_wav_peaks(const ttsa_plan* plan, const ttsa_batch* batch, const float* wav_dev, const int32_t* lens_dev,
                   float* peaks_dev, void* stream);
/* find_endpoint (utils/audio.py:203-210): window = int(sample_rate * min_silence_sec), hop = int(window / 4); the
 * first x in range(hop, len - window, hop) whose SIGNED maximum over wav[x : x + window] is below
 * 10^(threshold_db / 20) gives x + hop, else len.  endpoints_dev: [B] int32. */
int ttsa_find_endpoint(const ttsa_plan* plan, const ttsa_batch* batch, const float* wav_dev, double threshold_db,
                       double min_silence_sec, int32_t* endpoints_dev, void* stream);
/* save_wav's sample conversion (utils/audio.py:56-58): int16(wav * (32767 / max(0.01, max |wav|))), truncating like
 * numpy's astype.  Utterance u lands at out_off[u] = sum_{v<u} (len_v + gap_samples) followed by gap_samples zeros
 * (server/synthesizer.py:157-158 appends 10 000 zeros after every sentence and normalises the concatenation by ONE
 * peak: TTSA_PCM_JOINT_PEAK).  TTSA_PCM_F32_ARITH forms the product in float32 (the reference's arithmetic for a
 * float32 waveform); the default is float64 (float64 waveform: after de-emphasis, or the server's list).
 * out_off_dev: [B+1] int64 (out_off[B] = samples written); writes beyond out_capacity samples are dropped. */
#define TTSA_PCM_JOINT_PEAK 1u
#define TTSA_PCM_F32_ARITH  2u
size_t ttsa_pcm16_workspace_bytes(const ttsa_plan* plan, const ttsa_batch* batch);
int ttsa_wav_to_pcm16(const ttsa_plan* plan, const ttsa_batch* batch, const float* wav_dev, const int32_t* lens_dev,
                      uint32_t flags, int64_t gap_samples, int64_t* out_off_dev, int16_t* out_dev, int64_t out_capacity,
                      void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- elementwise steps of the API ------------------------------------------------------- */
#define TTSA_PW_NORMALIZE   0   /* _normalize    utils/audio.py:79-94   */
#define TTSA_PW_DENORMALIZE 1   /* _denormalize  utils/audio.py:96-112  */
#define TTSA_PW_AMP_TO_DB   2   /* _amp_to_db    utils/audio.py:121-123 */
#define TTSA_PW_DB_TO_AMP   3   /* _db_to_amp    utils/audio.py:125-126 */
int ttsa_pointwise(const ttsa_plan* plan, int op, const float* x_dev, float* y_dev, int64_t n, void* stream);
/* Layout shim for the reference's [D, T] numpy convention: out[c, r] = in[r, c]. */
int ttsa_transpose(const float* in_dev, float* out_dev, int64_t rows, int64_t cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTSA_H_ */
