/*
 * evfeat.h -- C ABI of libevfeat.so: the B200 (sm_100a) implementation of EveryVoice's
 * preprocessing feature-extraction hot path.
 *
 * The reference has no FFI for this path: it is a Python operator surface over
 * torchaudio.  Each entry point below names the reference
 * operator it replaces (paths relative to the EveryVoice repository root); the Python
 * shim `everyvoice_b200` keeps the reference's names on top of these calls through
 * ctypes (INTEGRATION.md shows the binding a maintainer would add).
 *
 * Conventions
 *   - plain C types only; every `*_dev` pointer is a CUDA device pointer owned by the
 *     caller (e.g. torch.Tensor.data_ptr()); every `*_host` pointer is host memory that
 *     is only read during the call;
 *   - every function returns an evf_status (0 == EVF_OK); evf_last_error() returns a
 *     thread-local, human readable description of the last failure;
 *   - no exception crosses the ABI; nothing is computed on the CPU; there is no fallback;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  All
 *     compute entry points are asynchronous with respect to the host;
 *   - a plan / batch is immutable after creation and may be used from several host
 *     threads on distinct streams.
 */
#ifndef EVFEAT_H_
#define EVFEAT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVF_ABI_VERSION 11

#if defined(__GNUC__)
#define EVF_API __attribute__((visibility("default")))
#else
#define EVF_API
#endif

typedef enum evf_status {
  EVF_OK = 0,
  EVF_ERR_INVALID_ARGUMENT = 1,
  EVF_ERR_UNSUPPORTED = 2,   /* spec_type unknown; n_fft > 14 000 (no such audio config) */
  EVF_ERR_SHORT_INPUT = 3,   /* an utterance has L <= n_fft/2: reflect padding undefined
                                (torch.stft raises RuntimeError for the same input)       */
  EVF_ERR_FILTERBANK = 4,    /* (no longer returned: a bank that is not triangular runs as a dense projection) */
  EVF_ERR_CUDA = 5,          /* a CUDA runtime call failed; see evf_last_error()          */
  EVF_ERR_NO_DEVICE = 6,     /* no sm_100 device: this library has no other code path     */
  EVF_ERR_OUT_OF_MEMORY = 7
} evf_status;

/* everyvoice/config/preprocessing_config.py:18-22  AudioSpecTypeEnum */
typedef enum evf_spec_type {
  EVF_SPEC_MEL = 0,          /* "mel"          utils/heavy.py:57-68   (htk scale, slaney norm, power) */
  EVF_SPEC_MEL_LIBROSA = 1,  /* "mel-librosa"  utils/heavy.py:69-100  (basis @ sqrt(power + 1e-9))   */
  EVF_SPEC_LINEAR = 2,       /* "linear"       utils/heavy.py:101-106 (power spectrogram)             */
  EVF_SPEC_RAW = 3           /* "raw"          utils/heavy.py:107-113 (complex STFT, no log/energy)   */
} evf_spec_type;

/* Which FFT kernel a plan uses.  AUTO: the warp-per-FFT kernel for n_fft 1024 (hop <= n_fft) and 2048 (even hop),
 * the any-size shared-memory mixed-radix kernel for every other n_fft / hop torch.stft accepts.  GENERIC forces the
 * any-size kernel (the parity tests cross-check the two implementations against each other). */
typedef enum evf_fft_path {
  EVF_FFT_AUTO = 0,
  EVF_FFT_GENERIC = 1
} evf_fft_path;

typedef enum evf_sample_format {
  EVF_SAMPLES_F32 = 0,       /* float32 in [-1, 1] (what torchaudio.load returns)                   */
  EVF_SAMPLES_S16 = 1        /* PCM16 as stored by process_audio; converted as s / 32768.0f on load */
} evf_sample_format;

/* The parameter contract of everyvoice/config/preprocessing_config.py:25-91 (AudioConfig)
 * as consumed by get_spectral_transform (utils/heavy.py:47-56). */
typedef struct evf_config {
  int32_t spec_type;      /* evf_spec_type                                                     */
  int32_t sample_rate;    /* informational (the mel basis is passed in explicitly)             */
  int32_t n_fft;          /* any n_fft >= 1 torch.stft accepts (up to 14 000)                  */
  int32_t win_length;     /* <= n_fft; informational (window is passed in explicitly)          */
  int32_t hop_length;     /* fft_hop_size                                                      */
  int32_t n_mels;         /* rows of the mel basis; ignored for linear / raw                   */
  int32_t apply_log;      /* 1: log(max(x, log_clip))  == dynamic_range_compression_torch,
                                utils/heavy.py:39-40 (Preprocessor.extract_spectral_features,
                                preprocessor/preprocessor.py:230-233, normalize=True)          */
  int32_t keep_last_frame;/* 0: T = L // hop frames (process_spec drops the last frame,
                                preprocessor/preprocessor.py:921-927);
                             1: T = L // hop + 1 (what the bare transform returns)             */
  int32_t sample_format;  /* evf_sample_format                                                 */
  float   log_clip;       /* 1e-5f                                                             */
  int32_t fft_path;       /* evf_fft_path; 0 = automatic                                       */
} evf_config;

typedef struct evf_plan evf_plan;
typedef struct evf_batch evf_batch;

EVF_API int         evf_abi_version(void);
EVF_API const char* evf_last_error(void);

/* Replaces the construction done by get_spectral_transform (utils/heavy.py:47-119).
 *   window_host : n_fft floats -- torch.hann_window(win_length) centre-padded to n_fft,
 *                 exactly the window torch.stft applies;
 *   mel_fb_host : [n_fft/2+1][n_mels] floats, row-major (frequency-major), the matrix
 *                 torchaudio.functional.melscale_fbanks returns / librosa.filters.mel
 *                 transposed; NULL for linear / raw.
 * Uploads window, FFT twiddles (computed in fp64) and the compressed mel table. */
EVF_API int evf_plan_create(const evf_config* cfg, const float* window_host, const float* mel_fb_host,
                    int device, evf_plan** plan_out);
EVF_API int evf_plan_destroy(evf_plan* plan);
/* floats written per frame: n_mels (mel types), n_fft/2+1 (linear), 2*(n_fft/2+1) (raw) */
EVF_API int evf_plan_row_floats(const evf_plan* plan, int32_t* row_floats_out);
/* frames produced for an utterance of n_samples: L // hop (keep_last_frame 0), else what torch.stft(center=True)
 * yields, 1 + (L + 2 * (n_fft // 2) - n_fft) // hop  (= L // hop + 1 for even n_fft) */
EVF_API int64_t evf_plan_num_frames(const evf_plan* plan, int64_t n_samples);

/* A ragged batch descriptor: utterance b occupies samples [sample_offsets[b],
 * sample_offsets[b+1]) of the packed sample buffer.  Computes per-utterance frame
 * counts / offsets and the frame-tile work list and uploads them. */
EVF_API int evf_batch_create(const evf_plan* plan, const int64_t* sample_offsets_host, int32_t n_utts,
                     evf_batch** batch_out);
EVF_API int evf_batch_destroy(evf_batch* batch);
EVF_API int evf_batch_total_frames(const evf_batch* batch, int64_t* total_frames_out);
/* copies frame_offsets[n_utts + 1] to host memory */
EVF_API int evf_batch_frame_offsets(const evf_batch* batch, int64_t* frame_offsets_host_out);
/* device copy of the same array (int64[n_utts + 1]); owned by the batch */
EVF_API int evf_batch_frame_offsets_dev(const evf_batch* batch, const int64_t** frame_offsets_dev_out);

/* Replaces, for every utterance of the batch at once,
 *   Preprocessor.extract_spectral_features(audio, transform)[:, :L // hop]
 *     (preprocessor/preprocessor.py:220-233, 921-927; utils/heavy.py:39-113) and
 *   Preprocessor.extract_energy(spec)  (preprocessor/preprocessor.py:302-309).
 *   samples_dev : packed float32 or int16 samples (cfg.sample_format)
 *   spec_out_dev: packed, time-major [total_frames][row_floats] float32
 *   energy_out_dev: [total_frames] float32, or NULL (must be NULL-able; ignored for raw) */
EVF_API int evf_features_run(const evf_plan* plan, const evf_batch* batch, const void* samples_dev,
                     float* spec_out_dev, float* energy_out_dev, void* stream);

/* The same for utterances [utt_begin, utt_end) of the batch only -- the chunked host pipeline
 * (one batch descriptor per shard, one launch per resident chunk).  The three pointers are the
 * addresses that sample 0 / frame 0 of the WHOLE batch would have: a caller holding only the
 * chunk [utt_begin, utt_end) in a device buffer passes `buffer - offset_of_chunk`; nothing outside
 * the chunk is touched.  For the 16-byte bulk-copy path the chunk must sit in its buffer at the
 * same (sample index * sample size) mod 16 as in the packed layout. */
EVF_API int evf_features_run_range(const evf_plan* plan, const evf_batch* batch, int32_t utt_begin, int32_t utt_end,
                           const void* samples_base_dev, float* spec_base_dev, float* energy_base_dev,
                           void* stream);

/* Convenience wrapper: batch_create + features_run + batch_destroy.
 * frame_offsets_host_out may be NULL. */
EVF_API int evf_features_ragged(const evf_plan* plan, const void* samples_dev,
                        const int64_t* sample_offsets_host, int32_t n_utts, float* spec_out_dev,
                        float* energy_out_dev, int64_t* frame_offsets_host_out, void* stream);

/* Replaces Preprocessor.extract_energy (preprocessor/preprocessor.py:302-309) on an
 * already computed time-major spectrogram [n_frames][row_floats]. */
EVF_API int evf_energy_from_spec(const float* spec_dev, int64_t n_frames, int32_t row_floats,
                         float* energy_out_dev, void* stream);

/* Replaces dynamic_range_compression_torch (utils/heavy.py:39-40) as a stand-alone
 * operator: out = log(max(in, clip_val) * c).  in_dev may equal out_dev. */
EVF_API int evf_log_compress(const float* in_dev, float* out_dev, int64_t n, float c, float clip_val,
                     void* stream);

/* Replaces Preprocessor.average_data_by_durations (preprocessor/preprocessor.py:287-300)
 * for a ragged batch: utterance b has values [value_offsets[b], value_offsets[b+1]) and
 * durations [phone_offsets[b], phone_offsets[b+1]); out has one float per duration.
 * Python slice semantics are reproduced exactly (clipping, d <= 0 -> 1e-7f, empty -> NaN). */
EVF_API int evf_segment_mean(const float* values_dev, const int64_t* value_offsets_dev,
                     const int64_t* durations_dev, const int64_t* phone_offsets_dev,
                     int32_t n_utts, float* out_dev, void* stream);

/* Replaces the reductions of Scaler.calculate_stats (preprocessor/helpers.py:86-106):
 * out5_dev = {count, sum, sum of squares, min, max} over the non-NaN values, float64.
 * With accumulate != 0 the result is merged into what out5_dev already holds (so several
 * buffers, and -- after an all-reduce -- several GPUs, can be combined). */
EVF_API int evf_stats_partial(const float* values_dev, int64_t n, double* out5_dev, int32_t accumulate,
                      void* stream);

/* Replaces Scaler.normalize (preprocessor/helpers.py:78-80) over a whole shard, in place. */
EVF_API int evf_normalize_inplace(float* values_dev, int64_t n, float mean, float std, void* stream);

/* Same, with mean = sum/n and std = sqrt((sumsq - sum^2/n)/(n-1)) derived on the device from
 * the (possibly all-reduced) five numbers of evf_stats_partial: no host round trip between
 * the reduction and the normalisation. */
EVF_API int evf_normalize_by_stats(float* values_dev, int64_t n, const double* stats5_dev, void* stream);

/* The multi-GPU exchange step (replaces gathering every file into one Scaler,
 * preprocessor/preprocessor.py:378-451): parts_dev holds n_parts five-number summaries, one per
 * rank, `stride_doubles` apart -- the output of ONE all-gather of evf_stats_partial results.
 * evf_stats_merge reduces them to one summary (SUM, SUM, SUM, MIN, MAX);
 * evf_normalize_by_gathered_stats normalises a shard in place straight from the gathered buffer
 * (ranks are merged in index order, so every rank derives bit-identical mean / std). */
EVF_API int evf_stats_merge(const double* parts_dev, int32_t n_parts, int32_t stride_doubles, double* out5_dev,
                    void* stream);
EVF_API int evf_normalize_by_gathered_stats(float* values_dev, int64_t n, const double* parts_dev, int32_t n_parts,
                                    int32_t stride_doubles, void* stream);

/* ---- pitch tracking: what Preprocessor.extract_pitch gets from pyworld,
 * everyvoice/preprocessor/preprocessor.py:257-277 -- pw.dio(x, sr, frame_period = hop / sr * 1000, speed = 4) followed by
 * pw.stonemask(x, f0, t, sr) -- for a ragged batch, in float64 like WORLD.  PARITY UNPINNED: pyworld (a wrapper of the
 * WORLD vocoder, C++) is a third-party dependency that is not available offline; the kernels restate WORLD's published
 * algorithm (dio.cpp, stonemask.cpp, matlabfunctions.cpp) and are tested against the CPU restatement oracle/world_pitch.py.
 *   evf_pitch_num_frames  : WORLD's GetSamplesForDIO, (int)(1000.0 * n / fs / frame_period) + 1, the same double arithmetic
 *   offsets_host          : [n_utts + 1] sample offsets into x_dev (HOST memory: the layout is planned on the host)
 *   f0_out_dev            : packed float64 tracks, evf_pitch_num_frames entries per utterance (0 = unvoiced)
 *   scratch_dev           : evf_pitch_scratch_bytes(...) bytes, 8-byte aligned
 * pyworld's defaults: f0_floor 71, f0_ceil 800, channels_in_octave 2, allowed_range 0.1. */
EVF_API int64_t evf_pitch_num_frames(int32_t sample_rate, double frame_period_ms, int64_t n_samples);
EVF_API int64_t evf_pitch_scratch_bytes(const int64_t* offsets_host, int32_t n_utts, int32_t sample_rate,
                                        double frame_period_ms, int32_t speed);
EVF_API int evf_pitch_dio_stonemask(const void* x_dev, int32_t x_format, const int64_t* offsets_host, int32_t n_utts,
                                    int32_t sample_rate, double frame_period_ms, int32_t speed, double f0_floor,
                                    double f0_ceil, double channels_in_octave, double allowed_range, void* scratch_dev,
                                    int64_t scratch_bytes, double* f0_out_dev, void* stream);

/* Pitch post-processing of Preprocessor.extract_pitch, everyvoice/preprocessor/preprocessor.py:278-285 (everything
 * after pyworld's dio / stonemask, which stay on the CPU with the reference): zeros (unvoiced) are filled by linear
 * interpolation over the frame index like np.interp, constant beyond the first / last voiced frame; an utterance
 * without a voiced frame becomes zeros.  pitch_dev: packed float64 tracks; out_dev: packed float32. */
EVF_API int evf_pitch_fill_unvoiced(const double* pitch_dev, const int64_t* offsets_dev, int32_t n_utts, float* out_dev,
                                    void* stream);

/* ---- backward of the transform: HiFiGAN trains through the mel spectrogram of the
 * generated audio, everyvoice/model/vocoder/HiFiGAN_iSTFT_lightning/hfgl/model.py:581-590, 719-721 -------------
 * grad_spec_dev: d loss / d spectrogram in the layout evf_features_run writes ([total_frames][row_floats],
 * linear domain: the plan must have apply_log == 0); grad_samples_dev: d loss / d samples, packed like the samples;
 * scratch_dev: evf_features_backward_scratch_floats(plan, batch) float32.  float32 samples, spec types mel /
 * mel-librosa / linear (raw is complex: unsupported).  Deterministic (no atomics). */
EVF_API int64_t evf_features_backward_scratch_floats(const evf_plan* plan, const evf_batch* batch);
EVF_API int evf_features_backward(const evf_plan* plan, const evf_batch* batch, const float* samples_dev,
                                  const float* grad_spec_dev, float* scratch_dev, float* grad_samples_dev,
                                  void* stream);
/* The same with the two steps that surround it in HiFiGAN's loss folded in (hfgl/model.py:719-721:
 * dynamic_range_compression_torch(transform(wav))[:, :, 1:] against the target mel):
 *   grad_layout   EVF_GRAD_FRAME_MAJOR: grad_dev is laid out like the forward output ([total_frames][row_floats]);
 *                 EVF_GRAD_BIN_MAJOR: per utterance [row_floats][T_b], utterance b at frame_offset[b] * row_floats --
 *                 what autograd hands back for the [B, F, T] tensor the transform returns (no transposing copy);
 *   log_spec_dev  NULL, or the log output of evf_features_run for the same batch (frame-major; the plan may then have
 *                 apply_log == 1): grad_dev is d loss / d log-spectrogram and is multiplied by d log / d spec =
 *                 exp(-log_spec) where log_spec > log(log_clip), 0 where the clamp was active (an element exactly AT
 *                 the clip value counts as clamped), NaN where log_spec is NaN. */
#define EVF_GRAD_FRAME_MAJOR 0
#define EVF_GRAD_BIN_MAJOR 1
EVF_API int evf_features_backward_ex(const evf_plan* plan, const evf_batch* batch, const float* samples_dev,
                                     const float* grad_dev, int32_t grad_layout, const float* log_spec_dev,
                                     float* scratch_dev, float* grad_samples_dev, void* stream);
/* backward of dynamic_range_compression_torch (utils/heavy.py:39-40): grad_in = grad_out / x where x >= clip_val,
 * else 0 (the clamp blocks the gradient) */
EVF_API int evf_log_compress_backward(const float* in_dev, const float* grad_out_dev, float* grad_in_dev, int64_t n,
                                      float clip_val, void* stream);

/* ---- audio front-end: the numerics of Preprocessor.process_audio ----------------
 * everyvoice/preprocessor/preprocessor.py:131-218.  Ragged batches: utterance b owns [offsets[b], offsets[b+1])
 * of a packed buffer; every offsets array has n_utts + 1 entries.  `x_dev` buffers are float32 or int16 PCM
 * (`x_format`, evf_sample_format; int16 is s / 32768 like torchaudio.load of a PCM16 wav). */
typedef struct evf_resampler evf_resampler;

/* torchaudio.functional.resample(audio, orig_freq, new_freq) as process_audio calls it (preprocessor.py:196-198):
 * sinc_interp_hann; torchaudio's defaults are lowpass_filter_width = 6, rolloff = 0.99.  The kernel bank is built
 * in double precision on the host exactly as torchaudio builds it. */
EVF_API int evf_resampler_create(int32_t orig_freq, int32_t new_freq, int32_t lowpass_filter_width, double rolloff,
                                 int32_t device, evf_resampler** resampler_out);
EVF_API int evf_resampler_destroy(evf_resampler* resampler);
/* ceil(new_freq * n_in / orig_freq): the length torchaudio crops the resampled waveform to; -1 on bad input */
EVF_API int64_t evf_resampler_out_length(const evf_resampler* resampler, int64_t n_in);
/* in_dev: packed float32 or int16 PCM (evf_sample_format; int16 is s / 32768 like torchaudio.load); out_dev: packed
 * float32 laid out by out_offsets_dev, whose lengths must be evf_resampler_out_length of the input lengths;
 * max_out_len = the longest of them (grid sizing). */
EVF_API int evf_audio_resample(const evf_resampler* resampler, const void* in_dev, int32_t in_format,
                               const int64_t* in_offsets_dev, const int64_t* out_offsets_dev, int32_t n_utts,
                               int64_t max_out_len, float* out_dev, void* stream);
/* absmax_dev[b] = max |x| of utterance b (NaN if it holds a NaN): torch.max(torch.abs(audio)), preprocessor.py:200 */
EVF_API int evf_audio_absmax(const void* x_dev, int32_t x_format, const int64_t* offsets_dev, int32_t n_utts,
                             int64_t max_len, float* absmax_dev, void* stream);
/* Peak normalisation + truncation + output format in one pass (preprocessor.py:199-201, 216-218; helpers.py:31-44):
 * for j < dst length: v = x[src_offsets[b] + j]; if absmax_dev: v = (v / absmax[b]) * 0.95f (two roundings, as the
 * reference's two in-place ops); written as float32 (out_f32_dev) and / or PCM16 (out_s16_dev: round-half-even of
 * v * 32768, clipped) at dst_offsets[b] + j.  dst lengths are the kept lengths (L // hop) * hop <= src lengths. */
EVF_API int evf_audio_finalize(const void* x_dev, int32_t x_format, const int64_t* src_offsets_dev,
                               const int64_t* dst_offsets_dev, int32_t n_utts, int64_t max_kept_len,
                               const float* absmax_dev, float* out_f32_dev, int16_t* out_s16_dev, void* stream);
/* torchaudio.transforms.Loudness(sr)(audio) for mono utterances (preprocessor.py:177-186: skipped when NaN or
 * < -36): ITU-R BS.1770-4 K-weighting, 400 ms blocks with 75 % overlap, absolute (-70) and relative (-10) gates.
 *   Fast pass: the filters' recursion is restarted every 100 ms with a 100 ms run-in, so all steps of all utterances
 *   run in parallel; about 2e-3 LKFS from torchaudio.
 *   Exact pass (refine_band_lkfs > 0): utterances whose loudness lies within the band of gate_lkfs (the caller's
 *   threshold, -36), or that hold a block within the band of a block-gating threshold, are re-evaluated with
 *   torchaudio's float32 arithmetic sample by sample (K-weighted signal bit-identical to the reference's), so the
 *   keep / skip decision is the reference's.  refine_flags_dev (int32[n_utts]) receives which ones were.
 * biquad_coeffs_host: 10 floats {b0, b1, b2, a1, a2} / a0 of the treble shelf and of the high-pass as the caller's
 *   torch build evaluates torchaudio's formulas (its sin / cos / exp can differ from libm's by an ulp, which the
 *   38 Hz high-pass amplifies); NULL: computed here with libm.
 * scratch_dev: float32, evf_audio_loudness_scratch_floats(sr, L_b) entries per utterance at scratch_offsets_dev;
 * max_len = the longest utterance (grid sizing). */
EVF_API int64_t evf_audio_loudness_scratch_floats(int32_t sample_rate, int64_t n_samples);
/* the 100 ms step in samples as torchaudio rounds it (round(round(0.4 * sr) * 0.25), half to even); -1 on bad input */
EVF_API int32_t evf_audio_loudness_step(int32_t sample_rate);
EVF_API int evf_audio_loudness(const void* x_dev, int32_t x_format, const int64_t* offsets_dev, int32_t n_utts,
                               int64_t max_len, int32_t sample_rate, const float* biquad_coeffs_host,
                               float refine_band_lkfs, float gate_lkfs, float* scratch_dev,
                               const int64_t* scratch_offsets_dev, int32_t* refine_flags_dev, float* lkfs_dev,
                               void* stream);

/* The loudness gate consumed on the device, without a host round trip (preprocessor.py:177-186: a file is skipped when
 * its loudness is NaN or < gate_lkfs = -36): keep_out_dev[b] = 1 / 0, and the per-utterance values of a skipped
 * utterance (values_dev[value_offsets[b] .. value_offsets[b + 1]), e.g. its phone-level energies; may be NULL) become
 * NaN, which the statistics (evf_stats_partial) skip.  The host drops the skipped utterances after the batch. */
EVF_API int evf_audio_gate_mask(const float* lkfs_dev, float gate_lkfs, float* values_dev,
                                const int64_t* value_offsets_dev, int32_t n_utts, int32_t* keep_out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EVFEAT_H_ */
