// Internal (non-ABI) declarations shared by the translation units of libevfeat.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "evfeat.h"

namespace evf {

// warps per CTA of the warp-per-FFT kernel; one FFT job per warp per tile.  Two CTAs of 8 warps per SM instead of one
// of 16: the same 16 resident warps, but tiles of 16 frames (less idle job slots at utterance ends) and two rings that
// drift independently: 0.689 -> 0.663 ms on configs[1] (profiles/r02y_kbench_cta_shape.txt)
constexpr int kMaxWarps = 8;
constexpr int kCtasPerSm = 16 / kMaxWarps;
constexpr int kFftSize = 1024;       // complex points per warp-level FFT
constexpr int kScrStride = 34;       // padded row stride of the per-warp transpose scratch (even: LDS.64 rows)

enum FftMode : int {
  MODE_PACK2 = 0,  // n_fft == 1024: two real frames packed as re/im of one complex FFT
  MODE_HALF = 1,   // n_fft == 2048: one real frame as a 1024-point complex FFT + split
  MODE_GENERIC = 2,  // any n_fft / hop: shared-memory mixed-radix FFT over frame pairs (evfeat_generic.cu)
  MODE_PACK2_512 = 3,  // n_fft == 512: a warp runs TWO packed 512-point jobs (4 frames) as 2 x (16 x 32)
  MODE_PACK2_256 = 4,  // n_fft == 256: a warp runs FOUR packed 256-point jobs (8 frames) as 4 x (8 x 32)
  MODE_DECIMATED = 5,  // n_fft == 3072 / 4096 with hop % (n_fft / 1024) == 0: phase streams through MODE_PACK2 + a combine
};
inline bool mode_is_pack2(int mode) { return mode == MODE_PACK2 || mode == MODE_PACK2_512 || mode == MODE_PACK2_256; }
// packed jobs (frame pairs) a warp runs side by side, and frames a warp owns per tile
inline int mode_jobs_per_warp(int mode) { return mode == MODE_PACK2_512 ? 2 : (mode == MODE_PACK2_256 ? 4 : 1); }
inline int mode_frames_per_warp(int mode) { return mode == MODE_HALF ? 1 : 2 * mode_jobs_per_warp(mode); }

// One frame tile of one utterance; built on the host by evf_batch_create so that the kernel
// needs no dependent loads (tile -> utterance -> offsets) to start staging a tile.
struct alignas(16) TileDesc {
  long long s_off;       // first sample of the utterance in the packed sample buffer
  long long out_frame0;  // first output frame of the tile in the packed output
  int L;                 // samples in the utterance
  int start;             // utterance-relative sample index of tile word 0 (may be negative)
  int nvalid;            // frames of the tile that exist
  int span;              // words of the input tile that are consumed
};
static_assert(sizeof(TileDesc) == 32, "TileDesc is loaded as two 16-byte words");

// Kernel parameters (passed by value; lives in the constant bank).
struct FeatParams {
  const void* samples;
  const TileDesc* tiles;
  int n_tiles;
  float* spec_out;
  float* energy_out;
  // plan tables (global memory; copied to shared memory once per CTA)
  const float* window;   // [n_fft], pre-scaled by 0.5; packed modes: stored as pairs {w[32r + lane], w[32(r+R/2) + lane]} at [r][lane], r < R/2, R = n_fft/32
  const float4* tw4;     // [16][32] four-step twiddles {t[n][lane], t[n+16][lane]}, t[n2][lane] = W_N^(n2*k1), N = min(n_fft, 1024), k1 = lane % (N/32)
  const float2* wpost;   // MODE_HALF: exp(-2 pi i k / n_fft), k = 0..512
  // mel projection (per-warp bin walk, see evfeat_features.cu):
  const float2* wtab;    // [n_chunk][32] {rising weight (sign bit = flush after this bin), falling weight} of bin n_chunk*lane + i
  const uint2* gtab;  // [(n_heads + 1)][m_pad] per filter: byte offsets (from the warp's slot area) of its c-th rising and c-th falling partial
  const unsigned* ltab;  // [32] per lane: slot of the first flush | slot of the second flush << 16 (later ones follow consecutively)
  int hop;
  int n_mels;
  int n_freq;
  int k_used;            // bins [0, k_used) carry a non-zero mel weight
  int n_chunk;           // bins walked by one lane (odd, so that the lanes' reads fall into distinct banks)
  int n_heads;           // most lanes that continue one interval started by an earlier lane
  int m_pad;             // n_mels rounded up to 32
  int n_slots;           // partial-sum slots per warp (non-empty intervals + 32 heads + 1 zero slot)
  int row_floats;
  int apply_log;
  float log_clip;
  // shared-memory carve-up, in 4-byte words from the start of dynamic shared memory
  int off_bar, off_in, off_in2, off_win, off_tw, off_wpost, off_wtab, off_gtab, off_ltab, off_warp;
  int warp_words;        // per-warp region: transpose scratch (aliased by the P column) + partial-sum slots
  int scr_words;         // words of that region before the slots (32 * kScrStride unless several jobs' P columns need more)
  int nbuf;              // input tile buffers in the ring (2 or 1)
  int in_words;          // capacity of one input tile
};

struct PlanTables {
  std::vector<float> window;      // n_fft, pre-scaled
  std::vector<float4> tw4;        // 512
  std::vector<float2> wpost;      // 513 (MODE_HALF) or empty
  std::vector<float2> melw;       // k_used {rising, falling}
  std::vector<int> kstart;        // n_mels + 2: interval j owns bins [kstart[j], kstart[j + 1])
  std::vector<int> jk;            // k_used + 1 interval of every bin (+ sentinel)
  std::vector<float2> wtab;       // n_chunk * 32
  std::vector<uint2> gtab;        // (n_heads + 1) * m_pad
  std::vector<unsigned> ltab;     // 32
  int k_used = 0, n_chunk = 1, n_heads = 0, m_pad = 32, n_slots = 34;
};

// Makes `dev` the current device for the lifetime of the guard and restores the previous one.
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define EVF_CUDA(call)                                          \
  do {                                                          \
    cudaError_t e_ = (call);                                    \
    if (e_ != cudaSuccess) return ::evf::cuda_fail(e_, #call);  \
  } while (0)

// evfeat_features.cu
int features_smem_bytes(int mode, int spec_type, int warps, int hop, int n_fft, const PlanTables& t,
                        FeatParams* carve);
int features_configure(int mode, int spec_type, int sample_format, int smem_bytes);
int features_launch(int mode, int spec_type, int sample_format, const FeatParams& p, int grid, int smem_bytes,
                    cudaStream_t stream);

// evfeat_generic.cu: the any-size kernel
constexpr int kGenMaxStages = 24;
struct GenStages {
  int n;
  int radix[kGenMaxStages];
};
struct GenParams {
  const void* samples;
  const TileDesc* tiles;
  int n_tiles;
  float* spec_out;
  float* energy_out;
  const float* window;     // [n_fft], pre-scaled by 0.5 (natural order)
  const float2* tw;        // [n_fft] W_N^m = exp(-2 pi i m / n_fft), fp64-built
  const double2* tw64;     // the same in float64; only when a prime factor > 5 needs a direct-DFT stage
  const float2* melw;      // [k_used] {rising, falling} weight of every bin (triangular banks)
  const int* kstart;       // [n_mels + 2] interval j (between filter centres j - 1 and j) owns bins [kstart[j], kstart[j + 1])
  const float* fb_dense;   // [n_freq][n_mels]: set instead of melw / kstart for a bank that is not triangular
  int n_fft, hop, n_freq, n_mels, row_floats, apply_log;
  float log_clip;
  int pairs;               // frame pairs (FFTs) per CTA iteration; a tile holds 2 * pairs frames
  GenStages st;
};
int generic_factorize(int n_fft, GenStages* st, bool* needs_tw64);
int generic_smem_bytes(int n_fft, int* pairs_out);
int generic_configure(int spec_type, int sample_format, int smem_bytes);
int generic_launch(int spec_type, int sample_format, const GenParams& p, int grid, int smem_bytes, cudaStream_t stream);
// Where the backward kernels read d loss / d spectrogram from (evf_features_backward_ex)
struct GradSrc {
  const float* grad;       // frame-major [total_frames][row] or, bin_major, per utterance [row][T_b]
  const float* log_spec;   // NULL, or the forward's log output (frame-major): grad is w.r.t. the log
  float log_clip;          // the forward's clip value
  int bin_major;
  int keep_last;           // frames of an utterance: L / hop (+ 1)
  int row;                 // row_floats
};
int generic_backward_launch(int spec_type, const GenParams& p, const GradSrc& gs, float* frame_grad, const int* jk,
                            int k_used, int grid, int smem_bytes, cudaStream_t st);
// evfeat_decimated.cu: n_fft = R * 1024 (R = 3, 4) as R phase-stream transforms of 1024 points + a radix-R combine
int decimated_deinterleave(const void* samples, int sample_format, const long long* sample_off_dev,
                           const long long* stream_off_dev, int utt0, int n_utts, long long chunk_soff0,
                           long long plane_stride, long long max_stream_len, int R, int n_fft, float* planes,
                           cudaStream_t st);
int decimated_combine(int R, int spec_type, const float* raw, long long plane_stride, long long n_frames,
                      float* spec_out, float* energy_out, const float2* wcomb, const float2* melw, const int* kstart,
                      int n_mels, int n_freq, int k_used, int row_floats, int apply_log, float log_clip, int num_sms,
                      cudaStream_t st);
// evfeat_backward.cu: folds the per-frame gradient rows back onto the samples (reflect padding included)
int overlap_add_launch(const float* frame_grad, const long long* sample_off, const long long* frame_off, int n_utts,
                       long long max_len, int n_fft, int hop, float* grad_samples, cudaStream_t st);

// evfeat_backward.cu
struct BwdParams {
  const float* samples;          // packed float32 audio
  const TileDesc* tiles;
  int n_tiles;
  GradSrc gs;                    // d loss / d spectrogram: layout, optional fused log
  float* frame_grad;             // scratch [total_frames][n_fft]
  float* grad_samples;           // out, packed like samples
  const float* window;           // plan tables (MODE_PACK2 layouts)
  const float4* tw4;
  const float2* wpost;           // MODE_HALF
  const float2* melw;            // [k_used] {rising weight -> filter j(k), falling weight -> filter j(k) - 1}
  const int* jk;                 // [k_used] interval index j(k)
  const long long* sample_off;   // [n_utts + 1]
  const long long* frame_off;    // [n_utts + 1]
  int n_utts;
  long long max_len;             // longest utterance (grid sizing)
  int hop, n_mels, k_used, row_floats, spec_type, n_fft;
};
int features_backward_launch(const BwdParams& p, cudaStream_t stream);
int launch_log_compress_backward(const float* x, const float* g, float* out, int64_t n, float clip, cudaStream_t s);

// evfeat_aux.cu
int launch_energy_from_spec(const float* spec, int64_t n_frames, int row, float* out, cudaStream_t s);
int launch_segment_mean(const float* values, const int64_t* value_off, const int64_t* durations,
                        const int64_t* phone_off, int n_utts, float* out, cudaStream_t s);
int launch_stats_partial(const float* values, int64_t n, double* out5, int accumulate, cudaStream_t s);
int launch_normalize(float* values, int64_t n, float mean, float std, cudaStream_t s);
int launch_normalize_by_stats(float* values, int64_t n, const double* stats5, cudaStream_t s);
int launch_normalize_by_gathered(float* values, int64_t n, const double* parts, int n_parts, int stride,
                                 cudaStream_t s);
int launch_stats_merge(const double* parts, int n_parts, int stride, double* out5, cudaStream_t s);
int launch_log_compress(const float* in, float* out, int64_t n, float c, float clip, cudaStream_t s);
int launch_pitch_fill_unvoiced(const double* pitch, const int64_t* offsets, int n_utts, float* out, cudaStream_t s);
int launch_gate_mask(const float* lkfs, float gate, float* values, const int64_t* offsets, int n_utts, int* keep_out,
                     cudaStream_t s);

}  // namespace evf
