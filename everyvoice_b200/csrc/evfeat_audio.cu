// Audio front-end of the feature path (sm_100a): what Preprocessor.process_audio does to a loaded
// waveform before process_spec sees it.
//
//   everyvoice/preprocessor/preprocessor.py:177-186   BS.1770 loudness gate  (torchaudio.transforms.Loudness)
//   everyvoice/preprocessor/preprocessor.py:196-198   torchaudio.functional.resample(audio, sr, resample_rate)
//   everyvoice/preprocessor/preprocessor.py:199-201   audio /= max|audio| ; audio *= 0.95
//   everyvoice/preprocessor/preprocessor.py:216-218   truncation to a multiple of the hop size
//   everyvoice/preprocessor/helpers.py:31-44          save_wav(..., PCM_S, 16 bit)  -> int16 samples
//
// Ragged batches throughout: utterance b owns [offsets[b], offsets[b + 1]) of a packed buffer.  The
// arithmetic of the third-party pieces (torchaudio 2.7.1 functional.resample / functional.loudness /
// lfilter) is restated from their published algorithm; oracle/ev_oracle.py holds the CPU restatement
// and tests/golden/frontend_*.npz the outputs of the live reference.
#include <cmath>
#include <new>
#include <vector>

#include "evfeat_internal.h"

struct evf_resampler {
  int device = 0;
  int orig = 1, neu = 1;  // reduced by their gcd
  int width = 0;          // zero padding on the left, torchaudio's `width`
  int taps = 0;           // 2 * width + orig
  float* d_kt = nullptr;  // [nk][neu] transposed, compacted kernel bank: consecutive output phases are consecutive words
  int* d_k0 = nullptr;    // [neu] first tap of each phase's support
  int nk = 0;             // taps kept per phase (the longest support; everything outside is exactly 0.0f)
  std::vector<float> h_kt;  // host copy of the full [taps][neu] bank: the small-ratio kernels take it by value
};

namespace evf {
namespace {

__device__ __forceinline__ float sample_to_float(float v) { return v; }
__device__ __forceinline__ float sample_to_float(short v) { return (float)v * (1.0f / 32768.0f); }

constexpr int kRsThreads = 256;
constexpr int kMaxGridY = 65535;  // gridDim.y limit: ragged (chunk, utterance) grids are launched in slices of utterances

// y[j] = sum_k K[j % new][k] * xpad[(j / new) * orig + k],  xpad = x shifted by `width` zeros
// (torchaudio _apply_sinc_resample_kernel: pad (width, width + orig), conv1d with stride orig, transpose, crop).
// torchaudio's bank has 2 * width + orig taps per phase, but the Hann window clamps the argument at
// +-lowpass_filter_width, where it is zero: for 48 -> 22.05 kHz only ~27 of the 348 float32 taps of a phase are
// non-zero, the rest are EXACTLY 0.0f.  The bank is stored compacted ([nk][new] from tap k0[phase] on), which skips
// exact-zero products in the same ascending order -- bit-identical for finite input.
// One block = 256 consecutive outputs of one utterance; their input span is staged in shared memory once
// (coalesced), the kernel bank is read transposed so that the 32 lanes of a warp (consecutive phases) read
// consecutive words.
template <typename SampleT>
__global__ void __launch_bounds__(kRsThreads) resample_kernel(const SampleT* __restrict__ x,
                                                              const long long* __restrict__ in_off,
                                                              const long long* __restrict__ out_off,
                                                              const float* __restrict__ kt,
                                                              const int* __restrict__ k0, int nk, int orig, int neu,
                                                              int width, int span_cap, float* __restrict__ y) {
  extern __shared__ float s_x[];
  const int b = blockIdx.y;
  const long long i0 = in_off[b], L = in_off[b + 1] - i0;
  const long long o0 = out_off[b], Lo = out_off[b + 1] - o0;
  const long long j0 = (long long)blockIdx.x * kRsThreads;
  if (j0 >= Lo) return;
  const long long blk0 = j0 / neu;                  // first input block of this tile
  const long long base = blk0 * orig - width;       // utterance-relative index of s_x[0]
  for (int e = threadIdx.x; e < span_cap; e += kRsThreads) {
    const long long n = base + e;
    s_x[e] = (n >= 0 && n < L) ? sample_to_float(x[i0 + n]) : 0.f;
  }
  __syncthreads();
  const long long j = j0 + threadIdx.x;
  if (j >= Lo) return;
  const int phase = (int)(j % neu);
  const int rel = (int)(j / neu - blk0) * orig;
  const float* kp = kt + phase;
  const float* xp = s_x + rel + __ldg(k0 + phase);
  float acc = 0.f;
#pragma unroll 4
  for (int k = 0; k < nk; ++k) acc = fmaf(__ldg(kp + (long long)k * neu), xp[k], acc);
  y[o0 + j] = acc;
}

// Small integer ratios (44.1 -> 22.05 kHz and back): a thread owns R consecutive input blocks = R * NEU consecutive
// outputs, whose (R - 1) * ORIG + TAPS input samples sit in registers.  The block's input span is staged in shared
// memory with coalesced loads; sample n lives at word n + n / (R * ORIG), so the windows of neighbouring threads start
// R * ORIG + 1 words apart (odd: the 32 lanes of a window read hit 32 distinct banks).  The kernel bank (<= 30
// coefficients) is a by-value kernel parameter, i.e. constant-bank operands of the FFMAs: no load instructions.
// Same summation order as resample_kernel (bit-identical results).
template <int N>
struct SmallBank {
  float k[N];
};
template <typename SampleT, int ORIG, int NEU, int TAPS, int R>
__global__ void __launch_bounds__(kRsThreads) resample_small_kernel(const SampleT* __restrict__ x,
                                                                    const long long* __restrict__ in_off,
                                                                    const long long* __restrict__ out_off,
                                                                    const SmallBank<TAPS * NEU> kt, int width,
                                                                    float* __restrict__ y) {
  constexpr int S = R * ORIG;                          // samples between the windows of neighbouring threads
  constexpr int W = (R - 1) * ORIG + TAPS;             // window of one thread
  constexpr int SPAN = (kRsThreads - 1) * S + W;       // samples of one block
  __shared__ float s_x[SPAN + SPAN / S + 1];
  const int b = blockIdx.y;
  const long long i0 = in_off[b], L = in_off[b + 1] - i0;
  const long long o0 = out_off[b], Lo = out_off[b + 1] - o0;
  const long long ib0 = (long long)blockIdx.x * kRsThreads * R;  // first input block of this CTA
  if (ib0 * NEU >= Lo) return;
  const long long base = ib0 * ORIG - width;           // utterance-relative sample index of staged word 0
  const SampleT* xs = x + i0;
  for (int e = threadIdx.x; e < SPAN; e += kRsThreads) {
    const long long n = base + e;
    s_x[e + e / S] = (n >= 0 && n < L) ? sample_to_float(__ldg(xs + n)) : 0.f;
  }
  __syncthreads();
  const long long ib = ib0 + (long long)threadIdx.x * R;
  if (ib * NEU >= Lo) return;
  float w[W];
  const float* sw = s_x + threadIdx.x * (S + 1);       // word of sample threadIdx.x * S
#pragma unroll
  for (int e = 0; e < W; ++e) w[e] = sw[e + e / S];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int ph = 0; ph < NEU; ++ph) {
      const long long j = (ib + r) * NEU + ph;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < TAPS; ++k) acc = fmaf(kt.k[k * NEU + ph], w[r * ORIG + k], acc);
      if (j < Lo) y[o0 + j] = acc;
    }
}

// max |x| per utterance (NaN-propagating like torch.max(torch.abs(.)): a NaN sample yields NaN).  A block owns 2048
// consecutive samples of one utterance; the 8 loads of a thread are independent (64 KB in flight per SM).
constexpr int kStreamItems = 8;
template <typename SampleT>
__global__ void __launch_bounds__(256) absmax_kernel(const SampleT* __restrict__ x, const long long* __restrict__ off,
                                                     unsigned* __restrict__ out_bits) {
  const int b = blockIdx.y;
  const long long o0 = off[b], L = off[b + 1] - o0;
  const long long j0 = (long long)blockIdx.x * (256 * kStreamItems) + threadIdx.x;
  if ((long long)blockIdx.x * (256 * kStreamItems) >= L) return;
  float v[kStreamItems];
#pragma unroll
  for (int i = 0; i < kStreamItems; ++i) {
    const long long j = j0 + i * 256;
    v[i] = (j < L) ? fabsf(sample_to_float(__ldg(x + o0 + j))) : 0.f;
  }
  float m = 0.f;
  bool nan = false;
#pragma unroll
  for (int i = 0; i < kStreamItems; ++i) {
    nan |= (v[i] != v[i]);
    m = fmaxf(m, v[i]);
  }
  unsigned bits = nan ? 0x7fc00000u : __float_as_uint(m);  // quiet NaN orders above every finite |x| and +inf
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
  __shared__ unsigned s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = bits;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned r = s[0];
    for (int w = 1; w < 8; ++w) r = max(r, s[w]);
    atomicMax(out_bits + b, r);
  }
}

// out = (x / max) * 0.95 (two roundings, exactly `audio /= max; audio *= 0.95` in fp32), truncated to the kept
// length, as float32 and / or PCM16 (round to nearest even of x * 32768, clipped: ffmpeg swresample flt -> s16)
template <typename SampleT>
__global__ void __launch_bounds__(256) finalize_kernel(const SampleT* __restrict__ x,
                                                       const long long* __restrict__ src_off,
                                                       const long long* __restrict__ dst_off,
                                                       const float* __restrict__ absmax, float* __restrict__ out_f32,
                                                       short* __restrict__ out_s16) {
  const int b = blockIdx.y;
  const long long s0 = src_off[b], d0 = dst_off[b], n = dst_off[b + 1] - d0;
  if ((long long)blockIdx.x * (256 * kStreamItems) >= n) return;
  const float m = absmax ? absmax[b] : 1.f;
  const long long j0 = (long long)blockIdx.x * (256 * kStreamItems) + threadIdx.x;
  float v[kStreamItems];
#pragma unroll
  for (int i = 0; i < kStreamItems; ++i) {
    const long long j = j0 + i * 256;
    v[i] = (j < n) ? sample_to_float(__ldg(x + s0 + j)) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < kStreamItems; ++i) {
    const long long j = j0 + i * 256;
    if (j >= n) break;
    float t = v[i];
    if (absmax) t = __fmul_rn(__fdiv_rn(t, m), 0.95f);
    if (out_f32) out_f32[d0 + j] = t;
    if (out_s16) {
      float q = rintf(t * 32768.0f);
      q = fminf(fmaxf(q, -32768.f), 32767.f);
      out_s16[d0 + j] = (t == t) ? (short)q : (short)0;
    }
  }
}

// ---- BS.1770-4 loudness (torchaudio.functional.loudness) ---------------------------------------
struct Biquad {
  float b0, b1, b2, a1, a2;  // already divided by a0 (torchaudio's _lfilter normalises the coefficients first)
};
struct LoudnessParams {
  Biquad shelf, highpass;
  int gate, step;  // 400 ms block and 100 ms step in samples: round(0.4 sr), round(gate / 4); gate - 4 step is in [-2, 2]
};
constexpr int kLdRec = 5;  // scratch floats per step: sum of z^2, z^2 of its first two and of its last two samples

// Pass 1 (fast): the two K-weighting biquads (torchaudio's lfilter(clamp=True) clamps each filter's OUTPUT to [-1, 1],
// the recursion itself runs on the unclamped state), squared and summed per 100 ms step.  Every sample leaves DRAM once
// and the time recursion is run as a SCAN inside a warp:
//   * a warp owns 8 consecutive steps of one utterance (plus one run-in step from a zero state: the slowest pole of
//     the 38 Hz high-pass -- a double pole at 1 - 2*pi*38/sr -- has decayed to n * r^n < 1e-7 after one step at every
//     sampling rate) and walks them in pieces of 32 * C samples (K pieces per step, the last one shorter); a piece is
//     copied once, coalesced, into a shared-memory tile, of which lane l owns the chunk [l * C, l * C + C) (C odd: the
//     32 lanes read 32 distinct banks);
//   * a biquad is linear in (input, state), so a lane runs its chunk from a ZERO state (transposed direct form II),
//     the warp then propagates the true states across the 32 chunk ends -- state_end[l] = M state_end[l - 1] + E[l],
//     M = A^C the free evolution of the state over a chunk: five shuffle rounds with the host-built powers M^(2^k)
//     -- and the lane adds the free response g[i] . state_start to its zero-state outputs;
//   * the clamp between the filters is pointwise, so the same three steps run twice: shelf -> clamp -> high-pass
//     (zero state) -> scan -> correct, clamp, square, accumulate.
// Three passes over the tile in shared memory instead of two passes over DRAM (the former run-in of every step).
// Accuracy: about 1e-3 LKFS against torchaudio (restart every 8 steps + transposed form); utterances whose gate
// decision could depend on that are re-evaluated by loudness_exact_kernel.
constexpr int kLdWarps = 1;        // warps per block: a warp is independent, and the ragged grid packs best in single warps
constexpr int kLdSteps = 8;        // steps per warp
constexpr int kLdMaxChunk = 80;    // samples per lane and piece (the host aims at <= 72)
// A 2 x 2 matrix as a float32 value plus a float32 correction (hi + lo of the float64 entry).  The high-pass' state
// transition has a DOUBLE pole at 1 - 2*pi*38/sr, i.e. it is a defective matrix, whose eigenvalues move with the square
// root of a perturbation: entries rounded to float32 alone (relative 6e-8 of entries of size 50) split the poles by
// 1e-3 and bias every step energy by 1e-3 at 96 kHz.  With the correction term the bias is below 1e-6.
struct Mat2 {
  float m00, m01, m10, m11;
  float l00, l01, l10, l11;
};
// M v + e, small terms first
__device__ __forceinline__ void mat_apply(const Mat2& M, float v1, float v2, float& e1, float& e2) {
  const float t1 = fmaf(M.l00, v1, fmaf(M.l01, v2, e1));
  const float t2 = fmaf(M.l10, v1, fmaf(M.l11, v2, e2));
  e1 = fmaf(M.m00, v1, fmaf(M.m01, v2, t1));
  e2 = fmaf(M.m10, v1, fmaf(M.m11, v2, t2));
}
struct LoudnessScan {
  int C, K;                        // chunk length (odd), pieces per step (K - 1 of 32 * C samples, then the rest)
  int n_act, c_last;               // last piece of a step: lanes that own samples, samples of the last of them
  Mat2 Ms[5], Mh[5];               // A^(C * 2^k), k = 0..4, of the shelf / the high-pass
  Mat2 Ms_last, Mh_last;           // A^c_last
  float2 gs[kLdMaxChunk];          // free response of the shelf: output i = gs[i].x * s1 + gs[i].y * s2
  float2 gh[kLdMaxChunk];          // ... of the high-pass
};

// state_end[l] = M state_end[l - 1] + E[l] over the lanes of a warp (carry = state before lane 0); returns the state
// at the START of every lane's chunk in (e1, e2)
__device__ __forceinline__ void scan_states(const Mat2 (&Mp)[5], float c1, float c2, float& e1, float& e2, int lane) {
  if (lane == 0) mat_apply(Mp[0], c1, c2, e1, e2);
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float p1 = __shfl_up_sync(0xffffffffu, e1, 1 << k);
    const float p2 = __shfl_up_sync(0xffffffffu, e2, 1 << k);
    if (lane >= (1 << k)) mat_apply(Mp[k], p1, p2, e1, e2);
  }
  const float i1 = __shfl_up_sync(0xffffffffu, e1, 1);
  const float i2 = __shfl_up_sync(0xffffffffu, e2, 1);
  e1 = lane ? i1 : c1;
  e2 = lane ? i2 : c2;
}

template <typename SampleT>
__global__ void __launch_bounds__(kLdWarps * 32) loudness_scan_kernel(const SampleT* __restrict__ x,
                                                                      const long long* __restrict__ off,
                                                                      const LoudnessParams P,
                                                                      const __grid_constant__ LoudnessScan S,
                                                                      float* __restrict__ scratch,
                                                                      const long long* __restrict__ scratch_off) {
  extern __shared__ __align__(16) float ld_smem[];
  const int C = S.C;
  float2* s_gs = reinterpret_cast<float2*>(ld_smem);
  float2* s_gh = s_gs + C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // float32 input is copied global -> shared without passing through registers (cp.async).  Copying the next piece
  // into a second tile during the recursion was measured and is no faster (0.78 ms against 0.72 ms at the same chunk
  // length, 0.73 ms with half the chunk length: the second tile costs resident warps).
  constexpr bool kAsync = (sizeof(SampleT) == 4);
  float* tile = ld_smem + 4 * C + warp * (32 * C);
  float* mw = tile + lane * C;
  for (int i = threadIdx.x; i < C; i += kLdWarps * 32) {
    s_gs[i] = S.gs[i];
    s_gh[i] = S.gh[i];
  }
  __syncthreads();
  const int b = blockIdx.y;
  const SampleT* xs = x + off[b];
  const long long L = off[b + 1] - off[b];
  const long long n_sub = (L + P.step - 1) / P.step;  // the last, partial step too: a block may reach 2 samples into it
  const long long q_first = ((long long)blockIdx.x * kLdWarps + warp) * kLdSteps;
  if (q_first >= n_sub) return;
  const long long q_end = (q_first + kLdSteps < n_sub) ? q_first + kLdSteps : n_sub;
  const int step = P.step, K = S.K, piece = 32 * C;
  const int base = lane * C;
  const Biquad s = P.shelf, h = P.highpass;
  float cs1 = 0.f, cs2 = 0.f, ch1 = 0.f, ch2 = 0.f;  // states at the start of the piece (zero at t = 0 and at the run-in)
  float* rec_base = scratch + scratch_off[b];
  const long long q_start = (q_first > 0) ? q_first - 1 : 0;
  // samples of piece (q, k): count, and how many of them lie inside the utterance
  auto piece_geom = [&](long long q, int k, long long& t0, int& n_here, int& n_ok) {
    n_here = (k == K - 1) ? step - (K - 1) * piece : piece;
    t0 = q * step + (long long)k * piece;
    const long long in_utt0 = L - t0;
    n_ok = in_utt0 < n_here ? (in_utt0 < 0 ? 0 : (int)in_utt0) : n_here;
  };
  auto copy_async = [&](float* dst, long long t0, int n_here, int n_ok) {
    const SampleT* xp = xs + (t0 < L ? t0 : 0);
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(dst);
    int j = lane;
    if (n_ok == n_here) {  // the whole piece lies inside the utterance: one instruction per copy
      for (; j + 7 * 32 < n_here; j += 8 * 32) {
        const uint32_t d = dst0 + 4u * (uint32_t)j;
        const SampleT* g = xp + j;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 128u * u), "l"(g + 32 * u) : "memory");
      }
    }
    for (; j < n_here; j += 32) {  // zero fill past the end of the utterance
      const int ok = j < n_ok;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst0 + 4u * (uint32_t)j),
                   "l"(xp + (ok ? j : 0)), "r"(ok ? 4 : 0)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (long long q = q_start; q < q_end; ++q) {
    const bool run_in = q < q_first;
    float acc_step = 0.f, head0 = 0.f, head1 = 0.f, tail0 = 0.f, tail1 = 0.f;
    for (int k = 0; k < K; ++k) {
      const bool last = (k == K - 1);
      const int n_act = last ? S.n_act : 32;
      const int c_last = last ? S.c_last : C;
      long long t0;
      int n_here, n_ok;
      piece_geom(q, k, t0, n_here, n_ok);
      __syncwarp();
      // One memory round trip per piece: float32 samples go global -> shared asynchronously (all of a lane's copies
      // in flight, zero fill past the end of the utterance); int16 samples pass through registers, 24 loads at a time.
      if constexpr (kAsync) {
        copy_async(tile, t0, n_here, n_ok);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      } else {
        const SampleT* xp = xs + t0;
        int j = lane;
        if (n_ok == n_here) {  // the whole piece lies inside the utterance: no bounds tests
          for (; j + 23 * 32 < n_here; j += 24 * 32) {
            float v[24];
#pragma unroll
            for (int u = 0; u < 24; ++u) v[u] = sample_to_float(__ldg(xp + j + 32 * u));
#pragma unroll
            for (int u = 0; u < 24; ++u) tile[j + 32 * u] = v[u];
          }
        }
        for (; j + 23 * 32 < n_here; j += 24 * 32) {
          float v[24];
#pragma unroll
          for (int u = 0; u < 24; ++u) v[u] = (j + 32 * u < n_ok) ? sample_to_float(__ldg(xp + j + 32 * u)) : 0.f;
#pragma unroll
          for (int u = 0; u < 24; ++u) tile[j + 32 * u] = v[u];
        }
        for (; j + 7 * 32 < n_here; j += 8 * 32) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = (j + 32 * u < n_ok) ? sample_to_float(__ldg(xp + j + 32 * u)) : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) tile[j + 32 * u] = v[u];
        }
        for (; j < n_here; j += 32) tile[j] = (j < n_ok) ? sample_to_float(__ldg(xp + j)) : 0.f;
      }
      {  // the next piece's lines towards L2 while this one is in the recursion
        const long long tn = t0 + n_here + (long long)lane * 32;
        if (tn < L && tn < q_end * step) asm volatile("prefetch.global.L2 [%0];" ::"l"(xs + tn));
        if (tn + 1024 < L && tn + 1024 < q_end * step) asm volatile("prefetch.global.L2 [%0];" ::"l"(xs + tn + 1024));
      }
      __syncwarp();
      // The chunk loops carry no per-lane bounds (branches would serialise each load behind the previous store):
      // every lane runs [0, c_last) and [c_last, C); the last lane's state is captured in between (its chunk ends
      // there); what it and the idle lanes compute afterwards -- on tile words past the piece -- never reaches a
      // lane that is used: the scan only passes states upwards, and the sums below are masked.
      // ---- shelf from a zero state; the outputs replace the samples -------------------------------------------------
      float e1 = 0.f, e2 = 0.f, z1, z2;
      auto shelf_run = [&](int i0, int i1) {
#pragma unroll 4
        for (int i = i0; i < i1; ++i) {
          const float x0 = mw[i];
          const float u0 = fmaf(s.b0, x0, e1);
          e1 = fmaf(s.b1, x0, fmaf(-s.a1, u0, e2));
          e2 = fmaf(s.b2, x0, -s.a2 * u0);
          mw[i] = u0;
        }
      };
      shelf_run(0, c_last);
      z1 = e1;
      z2 = e2;
      shelf_run(c_last, C);
      if (lane < n_act - 1) {
        z1 = e1;
        z2 = e2;
      }
      {
        const Mat2 Ml = last ? S.Ms_last : S.Ms[0];
        e1 = z1;
        e2 = z2;
        scan_states(S.Ms, cs1, cs2, e1, e2, lane);
        float n1 = z1, n2 = z2;  // the piece's end state, on its last lane
        mat_apply(Ml, e1, e2, n1, n2);
        cs1 = __shfl_sync(0xffffffffu, n1, n_act - 1);
        cs2 = __shfl_sync(0xffffffffu, n2, n_act - 1);
      }
      // ---- true shelf output -> clamp -> high-pass from a zero state ------------------------------------------------
      float f1 = 0.f, f2 = 0.f;
      auto hp_run = [&](int i0, int i1) {
#pragma unroll 4
        for (int i = i0; i < i1; ++i) {
          const float2 g = s_gs[i];
          const float u = fmaf(g.x, e1, fmaf(g.y, e2, mw[i]));
          const float cc = fminf(fmaxf(u, -1.f), 1.f);
          const float v0 = fmaf(h.b0, cc, f1);
          f1 = fmaf(h.b1, cc, fmaf(-h.a1, v0, f2));
          f2 = fmaf(h.b2, cc, -h.a2 * v0);
          mw[i] = v0;
        }
      };
      hp_run(0, c_last);
      z1 = f1;
      z2 = f2;
      hp_run(c_last, C);
      if (lane < n_act - 1) {
        z1 = f1;
        z2 = f2;
      }
      {
        const Mat2 Ml = last ? S.Mh_last : S.Mh[0];
        f1 = z1;
        f2 = z2;
        scan_states(S.Mh, ch1, ch2, f1, f2, lane);
        float n1 = z1, n2 = z2;
        mat_apply(Ml, f1, f2, n1, n2);
        ch1 = __shfl_sync(0xffffffffu, n1, n_act - 1);
        ch2 = __shfl_sync(0xffffffffu, n2, n_act - 1);
      }
      if (run_in) continue;
      // ---- true high-pass output -> clamp -> square -> sum; nothing past the piece or the end of the utterance ------
      // (per-lane trip count: only the lanes at the end of a step or of the utterance stop early)
      float acc = 0.f;
      const int left = min(max(n_ok - base, 0), C);
      int i = 0;
      for (; i + 4 <= left; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 g = s_gh[i + u];
          const float v = fmaf(g.x, f1, fmaf(g.y, f2, mw[i + u]));
          const float z = fminf(fmaxf(v, -1.f), 1.f);
          const float zz = z * z;
          acc += zz;
          mw[i + u] = zz;
        }
      }
      for (; i < left; ++i) {
        const float2 g = s_gh[i];
        const float v = fmaf(g.x, f1, fmaf(g.y, f2, mw[i]));
        const float z = fminf(fmaxf(v, -1.f), 1.f);
        const float zz = z * z;
        acc += zz;
        mw[i] = zz;
      }
      __syncwarp();
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      acc_step += acc;
      // z^2 of the step's first two and last two samples (tile word j = sample j of the piece; zero past the
      // end of the utterance, where the tile holds something else)
      auto zz_at = [&](int j) { return (j < n_ok) ? tile[j] : 0.f; };
      if (k == 0) {
        head0 = zz_at(0);
        head1 = zz_at(1);
      }
      if (last) {
        tail1 = zz_at(n_here - 1);
        if (n_here >= 2) tail0 = zz_at(n_here - 2);
      } else {
        tail0 = zz_at(piece - 1);  // the step's second-to-last sample if the last piece holds a single one
      }
    }
    if (!run_in && lane == 0) {
      float* rec = rec_base + kLdRec * q;
      rec[0] = acc_step;
      rec[1] = head0;
      rec[2] = head1;
      rec[3] = tail0;
      rec[4] = tail1;
    }
  }
}

// Host side of the scan: chunk geometry, powers of the state-transition matrix and the free responses, in float64 from
// the float32 coefficients the recursion uses.
static void mat_mul(const double (&a)[4], const double (&b)[4], double (&c)[4]) {
  const double r[4] = {a[0] * b[0] + a[1] * b[2], a[0] * b[1] + a[1] * b[3], a[2] * b[0] + a[3] * b[2],
                       a[2] * b[1] + a[3] * b[3]};
  for (int i = 0; i < 4; ++i) c[i] = r[i];
}
static Mat2 mat_split(const double (&P)[4]) {
  Mat2 M;
  float* hi = &M.m00;
  float* lo = &M.l00;
  for (int i = 0; i < 4; ++i) {
    hi[i] = (float)P[i];
    lo[i] = (float)(P[i] - (double)hi[i]);
  }
  return M;
}
static void scan_tables(const Biquad& q, int C, int c_last, Mat2 (&Mp)[5], Mat2& M_last, float2* g) {
  // zero input: u0 = s1; s1' = -a1 u0 + s2; s2' = -a2 u0  ->  state' = A state, output i = row 0 of A^i
  const double A[4] = {-(double)q.a1, 1.0, -(double)q.a2, 0.0};
  double P[4] = {1.0, 0.0, 0.0, 1.0};
  for (int i = 0; i < C; ++i) {
    g[i] = make_float2((float)P[0], (float)P[1]);
    if (i == c_last) M_last = mat_split(P);
    mat_mul(A, P, P);
  }
  if (c_last == C) M_last = mat_split(P);
  for (int k = 0; k < 5; ++k) {
    Mp[k] = mat_split(P);
    mat_mul(P, P, P);
  }
}

// Exact pass for the utterances flagged by the gate kernel: one warp per utterance, lane 0 runs torchaudio's float32
// arithmetic sample by sample from t = 0 -- coefficients normalised by a0, feed-forward part as
// fma(b0, x[t], fma(b1, x[t-1], b2 * x[t-2])) (conv1d's accumulation), recursion y = (ff - a2 y[t-2]) - a1 y[t-1] with
// separate roundings on the UNclamped outputs (_lfilter_core_loop), each filter's output clamped to [-1, 1] -- so the
// K-weighted signal is bit-identical to the reference's; the other lanes only fetch the samples (coalesced) and hand
// them over by shuffle.  Sequential, ~5 ms for a 10 s utterance: it only ever runs for utterances whose keep / skip
// decision is within the fast pass' error of a gating threshold.
template <typename SampleT>
__global__ void __launch_bounds__(32) loudness_exact_kernel(const SampleT* __restrict__ x,
                                                            const long long* __restrict__ off, LoudnessParams P,
                                                            const int* __restrict__ refine, float* __restrict__ scratch,
                                                            const long long* __restrict__ scratch_off) {
  const int b = blockIdx.x;
  if (!refine[b]) return;
  const SampleT* xs = x + off[b];
  const long long L = off[b + 1] - off[b];
  const int lane = threadIdx.x;
  float* rec = scratch + scratch_off[b];
  const Biquad s = P.shelf, h = P.highpass;
  float x1 = 0.f, x2 = 0.f, ya1 = 0.f, ya2 = 0.f;  // shelf: previous inputs / unclamped outputs
  float u1 = 0.f, u2 = 0.f, yb1 = 0.f, yb2 = 0.f;  // high-pass: previous (clamped shelf) inputs / unclamped outputs
  double acc = 0.0;
  float h0 = 0.f, h1 = 0.f, l0 = 0.f, l1 = 0.f;
  int k = 0;         // index inside the current step
  long long q = 0;   // current step
  const long long n_used = L;
  float nxt = (lane < n_used) ? sample_to_float(__ldg(xs + lane)) : 0.f;
  for (long long t0 = 0; t0 < n_used; t0 += 32) {
    const float cur = nxt;
    const long long tn = t0 + 32 + lane;
    nxt = (tn < n_used) ? sample_to_float(__ldg(xs + tn)) : 0.f;  // in flight during the 32 sequential steps below
    const int n_it = (int)((n_used - t0 < 32) ? (n_used - t0) : 32);
    for (int i = 0; i < n_it; ++i) {
      const float x0 = __shfl_sync(0xffffffffu, cur, i);
      if (lane == 0) {
        const float ffa = fmaf(s.b0, x0, fmaf(s.b1, x1, __fmul_rn(s.b2, x2)));
        const float ya = __fsub_rn(__fsub_rn(ffa, __fmul_rn(s.a2, ya2)), __fmul_rn(s.a1, ya1));
        x2 = x1;
        x1 = x0;
        ya2 = ya1;
        ya1 = ya;
        const float u0 = fminf(fmaxf(ya, -1.f), 1.f);
        const float ffb = fmaf(h.b0, u0, fmaf(h.b1, u1, __fmul_rn(h.b2, u2)));
        const float yb = __fsub_rn(__fsub_rn(ffb, __fmul_rn(h.a2, yb2)), __fmul_rn(h.a1, yb1));
        u2 = u1;
        u1 = u0;
        yb2 = yb1;
        yb1 = yb;
        const float z = fminf(fmaxf(yb, -1.f), 1.f);
        const float zz = __fmul_rn(z, z);
        acc += (double)zz;
        if (k == 0) h0 = zz;
        if (k == 1) h1 = zz;
        l0 = l1;
        l1 = zz;
        if (++k == P.step) {
          float* r = rec + kLdRec * q;
          r[0] = (float)acc;
          r[1] = h0;
          r[2] = h1;
          r[3] = l0;
          r[4] = l1;
          acc = 0.0;
          k = 0;
          ++q;
          h0 = h1 = 0.f;
        }
      }
    }
  }
  if (lane == 0 && k > 0) {  // the last, partial step (a block may reach 2 samples into it)
    float* r = rec + kLdRec * q;
    r[0] = (float)acc;
    r[1] = h0;
    r[2] = h1;
    r[3] = 0.f;
    r[4] = 0.f;
  }
}

// Pass 2, one thread per utterance: 400 ms block energies (four steps, 75 % overlap; gate - 4 * step samples are
// added from the next step's head or removed from the last step's tail) and the two gating passes.  With
// refine_out != NULL it also flags the utterances whose result the fast pass cannot be trusted with: loudness within
// `band` LKFS of the -36 LKFS gate, or a block within `band` of one of the two block-gating thresholds (a block that
// changes sides moves the mean by 1 / n_blocks).  With only_refined != NULL only flagged utterances are evaluated.
__global__ void __launch_bounds__(64) loudness_gate_kernel(const long long* __restrict__ off, int n_utts,
                                                           LoudnessParams P, const float* __restrict__ scratch,
                                                           const long long* __restrict__ scratch_off,
                                                           float* __restrict__ lkfs, float band, float gate_lkfs,
                                                           int* __restrict__ refine_out,
                                                           const int* __restrict__ only_refined) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_utts) return;
  if (only_refined != nullptr && !only_refined[b]) return;
  const long long L = off[b + 1] - off[b];
  const float* sub = scratch + scratch_off[b];
  const long long n_blk = (L >= P.gate) ? (L - P.gate) / P.step + 1 : 0;
  if (n_blk == 0) {
    lkfs[b] = __int_as_float(0x7fc00000);  // shorter than one gating block: the reference cannot measure it
    if (refine_out) refine_out[b] = 0;
    return;
  }
  const float inv_gate = 1.0f / (float)P.gate;
  const int extra = P.gate - 4 * P.step;  // in [-2, 2]
  auto energy = [&](long long i) {
    const float* r = sub + kLdRec * i;
    float e = (r[0] + r[kLdRec]) + (r[2 * kLdRec] + r[3 * kLdRec]);
    if (extra > 0) e += r[4 * kLdRec + 1] + (extra > 1 ? r[4 * kLdRec + 2] : 0.f);      // head of step i + 4
    if (extra < 0) e -= r[3 * kLdRec + 4] + (extra < -1 ? r[3 * kLdRec + 3] : 0.f);     // tail of step i + 3
    return e * inv_gate;
  };
  auto lk = [](float e) { return -0.691f + 10.0f * log10f(e); };
  bool fragile = false;
  float sum = 0.f;
  int cnt = 0;
  for (long long i = 0; i < n_blk; ++i) {  // absolute gate (-70 LKFS)
    const float e = energy(i);
    const float l = lk(e);
    if (fabsf(l + 70.0f) < band) fragile = true;
    if (l > -70.0f) {
      sum += e;
      ++cnt;
    }
  }
  const float gamma_rel = lk(sum / (float)cnt) - 10.0f;  // cnt == 0 -> NaN, like the reference
  sum = 0.f;
  cnt = 0;
  for (long long i = 0; i < n_blk; ++i) {  // relative gate (-10 LU below the absolute-gated mean)
    const float e = energy(i);
    const float l = lk(e);
    if (fabsf(l - gamma_rel) < band) fragile = true;
    if (l > -70.0f && l > gamma_rel) {
      sum += e;
      ++cnt;
    }
  }
  const float out = lk(sum / (float)cnt);
  lkfs[b] = out;
  if (refine_out) refine_out[b] = (band > 0.f && (fragile || fabsf(out - gate_lkfs) < band)) ? 1 : 0;
}

// RBJ cookbook biquads as torchaudio.functional.{treble_biquad, highpass_biquad} build them: every operation in
// the waveform's dtype (float32), then a1, a2 (and here b0..b2, which lfilter applies before dividing) over a0.
Biquad make_treble(float sr, float gain_db, float fc, float Q) {
  const float w0 = 6.283185307179586f * fc / sr, alpha = sinf(w0) / 2.0f / Q;
  const float A = expf(gain_db / 40.0f * 2.302585092994046f);
  const float t1 = 2.0f * sqrtf(A) * alpha, t2 = (A - 1.0f) * cosf(w0), t3 = (A + 1.0f) * cosf(w0);
  const float b0 = A * ((A + 1.0f) + t2 + t1), b1 = -2.0f * A * ((A - 1.0f) + t3), b2 = A * ((A + 1.0f) + t2 - t1);
  const float a0 = (A + 1.0f) - t2 + t1, a1 = 2.0f * ((A - 1.0f) - t3), a2 = (A + 1.0f) - t2 - t1;
  return Biquad{b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
}
Biquad make_highpass(float sr, float fc, float Q) {
  const float w0 = 6.283185307179586f * fc / sr, alpha = sinf(w0) / 2.0f / Q;
  const float b0 = (1.0f + cosf(w0)) / 2.0f, b1 = -1.0f - cosf(w0), b2 = b0;
  const float a0 = 1.0f + alpha, a1 = -2.0f * cosf(w0), a2 = 1.0f - alpha;
  return Biquad{b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
}

long long gcd_ll(long long a, long long b) {
  while (b) {
    const long long t = a % b;
    a = b;
    b = t;
  }
  return a;
}

}  // namespace
}  // namespace evf

using namespace evf;

extern "C" {

int evf_resampler_create(int32_t orig_freq, int32_t new_freq, int32_t lowpass_filter_width, double rolloff,
                         int32_t device, evf_resampler** out) {
  if (out) *out = nullptr;
  if (!out || orig_freq < 1 || new_freq < 1 || lowpass_filter_width < 1 || !(rolloff > 0.0)) {
    set_error("evf_resampler_create: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
    set_error("evf_resampler_create: no such CUDA device; libevfeat has no CPU path");
    return EVF_ERR_NO_DEVICE;
  }
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  evf_resampler* r = new (std::nothrow) evf_resampler();
  if (!r) return EVF_ERR_OUT_OF_MEMORY;
  r->device = device;
  const long long g = gcd_ll(orig_freq, new_freq);
  r->orig = (int)(orig_freq / g);
  r->neu = (int)(new_freq / g);
  // torchaudio _get_sinc_resample_kernel, sinc_interp_hann, evaluated in double then cast to float
  const double base_freq = (double)(r->orig < r->neu ? r->orig : r->neu) * rolloff;
  r->width = (int)std::ceil((double)lowpass_filter_width * r->orig / base_freq);
  r->taps = 2 * r->width + r->orig;
  if ((long long)r->taps * r->neu > (64ll << 20)) {
    delete r;
    set_error("evf_resampler_create: rate ratio needs a kernel bank of more than 64 Mi entries");
    return EVF_ERR_UNSUPPORTED;
  }
  std::vector<float> kt((size_t)r->taps * r->neu);
  const double lpw = (double)lowpass_filter_width, scale = base_freq / r->orig;
  for (int p = 0; p < r->neu; ++p)
    for (int k = 0; k < r->taps; ++k) {
      double t = (double)(-p) / r->neu + (double)(k - r->width) / r->orig;
      t *= base_freq;
      t = t < -lpw ? -lpw : (t > lpw ? lpw : t);
      const double c = std::cos(t * M_PI / lpw / 2.0);
      const double window = c * c;
      t *= M_PI;
      const double sinc = (t == 0.0) ? 1.0 : std::sin(t) / t;
      kt[(size_t)k * r->neu + p] = (float)(sinc * window * scale);
    }
  // support of every phase: first / last tap that is not exactly 0.0f
  std::vector<int> k0(r->neu, 0);
  int nk = 1;
  for (int p = 0; p < r->neu; ++p) {
    int first = r->taps, last = -1;
    for (int k = 0; k < r->taps; ++k)
      if (kt[(size_t)k * r->neu + p] != 0.0f) {
        if (first == r->taps) first = k;
        last = k;
      }
    if (last < 0) first = last = 0;
    k0[p] = first;
    nk = (last - first + 1 > nk) ? last - first + 1 : nk;
  }
  for (int p = 0; p < r->neu; ++p)
    if (k0[p] + nk > r->taps) k0[p] = r->taps - nk;  // keep every phase's window inside the bank (extra taps are zeros)
  std::vector<float> kc((size_t)nk * r->neu);
  for (int p = 0; p < r->neu; ++p)
    for (int t = 0; t < nk; ++t) kc[(size_t)t * r->neu + p] = kt[(size_t)(k0[p] + t) * r->neu + p];
  r->nk = nk;
  cudaError_t e = cudaMalloc(&r->d_kt, kc.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(r->d_kt, kc.data(), kc.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&r->d_k0, k0.size() * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(r->d_k0, k0.data(), k0.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(r->d_kt);
    cudaFree(r->d_k0);
    delete r;
    return cuda_fail(e, "resampler kernel bank upload");
  }
  r->h_kt.swap(kt);
  *out = r;
  return EVF_OK;
}

int evf_resampler_destroy(evf_resampler* r) {
  if (!r) return EVF_OK;
  DeviceGuard guard(r->device);
  cudaFree(r->d_kt);
  cudaFree(r->d_k0);
  delete r;
  return EVF_OK;
}

int64_t evf_resampler_out_length(const evf_resampler* r, int64_t n_in) {
  if (!r || n_in < 0) return -1;
  return (n_in * r->neu + r->orig - 1) / r->orig;  // ceil(new * n / orig), exact in integers
}

int evf_audio_resample(const evf_resampler* r, const void* in_dev, int32_t in_format, const int64_t* in_offsets_dev,
                       const int64_t* out_offsets_dev, int32_t n_utts, int64_t max_out_len, float* out_dev,
                       void* stream) {
  if (!r || (n_utts > 0 && (!in_dev || !in_offsets_dev || !out_offsets_dev || !out_dev)) || n_utts < 0 ||
      max_out_len < 0 || (in_format != EVF_SAMPLES_F32 && in_format != EVF_SAMPLES_S16)) {
    set_error("evf_audio_resample: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (n_utts == 0 || max_out_len == 0) return EVF_OK;
  DeviceGuard guard(r->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* io = reinterpret_cast<const long long*>(in_offsets_dev);
  const long long* oo = reinterpret_cast<const long long*>(out_offsets_dev);
  const bool s16 = (in_format == EVF_SAMPLES_S16);
  // register-window kernels for the two common ratios (lowpass_filter_width 6, rolloff 0.99: 28 resp. 15 taps)
  constexpr int R = 4;
  auto small_grid = [&](int neu, int ny) {
    const long long blocks_in = (max_out_len + neu - 1) / neu;  // input blocks of the longest utterance
    return dim3((unsigned)((blocks_in + (long long)R * kRsThreads - 1) / ((long long)R * kRsThreads)), (unsigned)ny);
  };
  if (r->orig == 2 && r->neu == 1 && r->taps == 28) {
    SmallBank<28> bank;
    for (int i = 0; i < 28; ++i) bank.k[i] = r->h_kt[i];
    for (int y0 = 0; y0 < n_utts; y0 += kMaxGridY) {
      const int ny = n_utts - y0 < kMaxGridY ? n_utts - y0 : kMaxGridY;
      if (s16)
        resample_small_kernel<short, 2, 1, 28, R><<<small_grid(1, ny), kRsThreads, 0, st>>>(
            static_cast<const short*>(in_dev), io + y0, oo + y0, bank, r->width, out_dev);
      else
        resample_small_kernel<float, 2, 1, 28, R><<<small_grid(1, ny), kRsThreads, 0, st>>>(
            static_cast<const float*>(in_dev), io + y0, oo + y0, bank, r->width, out_dev);
      EVF_CUDA(cudaGetLastError());
    }
    return EVF_OK;
  }
  if (r->orig == 1 && r->neu == 2 && r->taps == 15) {
    SmallBank<30> bank;
    for (int i = 0; i < 30; ++i) bank.k[i] = r->h_kt[i];
    for (int y0 = 0; y0 < n_utts; y0 += kMaxGridY) {
      const int ny = n_utts - y0 < kMaxGridY ? n_utts - y0 : kMaxGridY;
      if (s16)
        resample_small_kernel<short, 1, 2, 15, R><<<small_grid(2, ny), kRsThreads, 0, st>>>(
            static_cast<const short*>(in_dev), io + y0, oo + y0, bank, r->width, out_dev);
      else
        resample_small_kernel<float, 1, 2, 15, R><<<small_grid(2, ny), kRsThreads, 0, st>>>(
            static_cast<const float*>(in_dev), io + y0, oo + y0, bank, r->width, out_dev);
      EVF_CUDA(cudaGetLastError());
    }
    return EVF_OK;
  }
  const int span = ((kRsThreads - 1) / r->neu + 2) * r->orig + r->taps;
  const size_t smem = (size_t)span * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("evf_audio_resample: rate ratio needs more than 200 KB of shared memory per block");
    return EVF_ERR_UNSUPPORTED;
  }
  for (int y0 = 0; y0 < n_utts; y0 += kMaxGridY) {
    const int ny = n_utts - y0 < kMaxGridY ? n_utts - y0 : kMaxGridY;
    const dim3 grid((unsigned)((max_out_len + kRsThreads - 1) / kRsThreads), (unsigned)ny);
    if (s16) {
      auto k = resample_kernel<short>;
      if (smem > 48 * 1024) EVF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, kRsThreads, smem, st>>>(static_cast<const short*>(in_dev), io + y0, oo + y0, r->d_kt, r->d_k0, r->nk,
                                        r->orig, r->neu, r->width, span, out_dev);
    } else {
      auto k = resample_kernel<float>;
      if (smem > 48 * 1024) EVF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, kRsThreads, smem, st>>>(static_cast<const float*>(in_dev), io + y0, oo + y0, r->d_kt, r->d_k0, r->nk,
                                        r->orig, r->neu, r->width, span, out_dev);
    }
    EVF_CUDA(cudaGetLastError());
  }
  return EVF_OK;
}

int evf_audio_absmax(const void* x_dev, int32_t x_format, const int64_t* offsets_dev, int32_t n_utts, int64_t max_len,
                     float* absmax_dev, void* stream) {
  if (n_utts < 0 || max_len < 0 || (x_format != EVF_SAMPLES_F32 && x_format != EVF_SAMPLES_S16) ||
      (n_utts > 0 && (!x_dev || !offsets_dev || !absmax_dev))) {
    set_error("evf_audio_absmax: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (n_utts == 0) return EVF_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EVF_CUDA(cudaMemsetAsync(absmax_dev, 0, (size_t)n_utts * sizeof(float), st));
  if (max_len == 0) return EVF_OK;
  const long long gx = (max_len + 256 * kStreamItems - 1) / (256 * kStreamItems);
  const long long* off = reinterpret_cast<const long long*>(offsets_dev);
  unsigned* out = reinterpret_cast<unsigned*>(absmax_dev);
  for (int y0 = 0; y0 < n_utts; y0 += kMaxGridY) {
    const dim3 grid((unsigned)gx, (unsigned)(n_utts - y0 < kMaxGridY ? n_utts - y0 : kMaxGridY));
    if (x_format == EVF_SAMPLES_S16)
      absmax_kernel<short><<<grid, 256, 0, st>>>(static_cast<const short*>(x_dev), off + y0, out + y0);
    else
      absmax_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x_dev), off + y0, out + y0);
    EVF_CUDA(cudaGetLastError());
  }
  return EVF_OK;
}

int evf_audio_finalize(const void* x_dev, int32_t x_format, const int64_t* src_offsets_dev,
                       const int64_t* dst_offsets_dev, int32_t n_utts, int64_t max_kept_len, const float* absmax_dev,
                       float* out_f32_dev, int16_t* out_s16_dev, void* stream) {
  if (n_utts < 0 || max_kept_len < 0 || (x_format != EVF_SAMPLES_F32 && x_format != EVF_SAMPLES_S16)) {
    set_error("evf_audio_finalize: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (n_utts == 0 || max_kept_len == 0) return EVF_OK;  // nothing kept: no output buffer needed
  if (!x_dev || !src_offsets_dev || !dst_offsets_dev || (!out_f32_dev && !out_s16_dev)) {
    set_error("evf_audio_finalize: null pointer");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  const long long gx = (max_kept_len + 256 * kStreamItems - 1) / (256 * kStreamItems);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* so = reinterpret_cast<const long long*>(src_offsets_dev);
  const long long* dd = reinterpret_cast<const long long*>(dst_offsets_dev);
  short* o16 = reinterpret_cast<short*>(out_s16_dev);
  for (int y0 = 0; y0 < n_utts; y0 += kMaxGridY) {
    const dim3 grid((unsigned)gx, (unsigned)(n_utts - y0 < kMaxGridY ? n_utts - y0 : kMaxGridY));
    const float* am = absmax_dev ? absmax_dev + y0 : nullptr;
    if (x_format == EVF_SAMPLES_S16)
      finalize_kernel<short><<<grid, 256, 0, st>>>(static_cast<const short*>(x_dev), so + y0, dd + y0, am, out_f32_dev, o16);
    else
      finalize_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x_dev), so + y0, dd + y0, am, out_f32_dev, o16);
    EVF_CUDA(cudaGetLastError());
  }
  return EVF_OK;
}

int32_t evf_audio_loudness_step(int32_t sample_rate) {
  if (sample_rate < 1) return -1;
  const long long gate = (long long)std::nearbyint(0.4 * (double)sample_rate);
  const long long step = (long long)std::nearbyint((double)gate * (1.0 - 0.75));
  return step < 1 ? -1 : (int32_t)step;
}

int64_t evf_audio_loudness_scratch_floats(int32_t sample_rate, int64_t n_samples) {
  const long long step = evf_audio_loudness_step(sample_rate);
  if (step < 1 || n_samples < 0) return -1;
  // kLdRec floats per 100 ms step; a block energy reads up to 4 steps past its index
  return kLdRec * (n_samples / step + 5);
}

int evf_audio_loudness(const void* x_dev, int32_t x_format, const int64_t* offsets_dev, int32_t n_utts,
                       int64_t max_len, int32_t sample_rate, const float* biquad_coeffs_host,
                       float refine_band_lkfs, float gate_lkfs, float* scratch_dev,
                       const int64_t* scratch_offsets_dev, int32_t* refine_flags_dev, float* lkfs_dev, void* stream) {
  if (n_utts < 0 || sample_rate < 1 || max_len < 0 || (x_format != EVF_SAMPLES_F32 && x_format != EVF_SAMPLES_S16) ||
      !(refine_band_lkfs >= 0.f) ||
      (n_utts > 0 && (!x_dev || !offsets_dev || !scratch_dev || !scratch_offsets_dev || !lkfs_dev)) ||
      (n_utts > 0 && refine_band_lkfs > 0.f && !refine_flags_dev)) {
    set_error("evf_audio_loudness: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  LoudnessParams P;
  // Python's round() is round-half-even; 0.4 * sr and gate * 0.25 are compared on the same doubles
  P.gate = (int)std::nearbyint(0.4 * (double)sample_rate);
  P.step = (int)std::nearbyint((double)P.gate * (1.0 - 0.75));
  if (P.step < 2 || P.gate - 4 * P.step < -2 || P.gate - 4 * P.step > 2) {
    set_error("evf_audio_loudness: sampling rate too low for 100 ms steps");
    return EVF_ERR_UNSUPPORTED;
  }
  if (n_utts == 0) return EVF_OK;
  if (biquad_coeffs_host != nullptr) {  // {b0, b1, b2, a1, a2} / a0 of the shelf, then of the high-pass
    const float* c = biquad_coeffs_host;
    P.shelf = Biquad{c[0], c[1], c[2], c[3], c[4]};
    P.highpass = Biquad{c[5], c[6], c[7], c[8], c[9]};
  } else {
    P.shelf = make_treble((float)sample_rate, 4.0f, 1500.0f, (float)(1.0 / std::sqrt(2.0)));
    P.highpass = make_highpass((float)sample_rate, 38.0f, 0.5f);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* off = reinterpret_cast<const long long*>(offsets_dev);
  const long long* soff = reinterpret_cast<const long long*>(scratch_offsets_dev);
  const long long max_sub = (max_len + P.step - 1) / P.step;
  if (max_sub > 0) {
    LoudnessScan S;  // by-value kernel parameters, rebuilt per call (a few hundred flops)
    // pieces of 32 * C samples with C <= 72 (odd): short chunks keep the per-warp tile small (more resident warps)
    // (chunk targets of 40 / 24 samples -- more resident warps, more scans -- measured 0.63 / 0.70 ms against 0.60 ms)
    S.K = (P.step + 32 * 72 - 1) / (32 * 72);
    S.C = ((P.step + 32 * S.K - 1) / (32 * S.K)) | 1;
    while (S.K > 1 && (S.K - 1) * 32 * S.C >= P.step) --S.K;  // the last piece must hold at least one sample
    const int rest = P.step - (S.K - 1) * 32 * S.C;
    S.n_act = (rest + S.C - 1) / S.C;
    S.c_last = rest - (S.n_act - 1) * S.C;
    if (S.C > kLdMaxChunk || P.step < 64 || rest < 1 || S.n_act < 1 || S.n_act > 32) {
      set_error("evf_audio_loudness: sampling rate too low for the 100 ms steps of the loudness pass");
      return EVF_ERR_UNSUPPORTED;
    }
    scan_tables(P.shelf, S.C, S.c_last, S.Ms, S.Ms_last, S.gs);
    scan_tables(P.highpass, S.C, S.c_last, S.Mh, S.Mh_last, S.gh);
    const int smem = (4 * S.C + kLdWarps * 32 * S.C) * (int)sizeof(float);
    auto kern_s16 = loudness_scan_kernel<short>;
    auto kern_f32 = loudness_scan_kernel<float>;
    if (x_format == EVF_SAMPLES_S16)
      EVF_CUDA(cudaFuncSetAttribute(kern_s16, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else
      EVF_CUDA(cudaFuncSetAttribute(kern_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long per_block = (long long)kLdWarps * kLdSteps;
    for (int y0 = 0; y0 < n_utts; y0 += kMaxGridY) {
      const dim3 grid((unsigned)((max_sub + per_block - 1) / per_block),
                      (unsigned)(n_utts - y0 < kMaxGridY ? n_utts - y0 : kMaxGridY));
      if (x_format == EVF_SAMPLES_S16)
        kern_s16<<<grid, kLdWarps * 32, smem, st>>>(static_cast<const short*>(x_dev), off + y0, P, S, scratch_dev,
                                                    soff + y0);
      else
        kern_f32<<<grid, kLdWarps * 32, smem, st>>>(static_cast<const float*>(x_dev), off + y0, P, S, scratch_dev,
                                                    soff + y0);
      EVF_CUDA(cudaGetLastError());
    }
  }
  const bool refine = refine_band_lkfs > 0.f;
  loudness_gate_kernel<<<(n_utts + 63) / 64, 64, 0, st>>>(off, n_utts, P, scratch_dev, soff, lkfs_dev, refine_band_lkfs,
                                                          gate_lkfs, refine ? refine_flags_dev : nullptr, nullptr);
  EVF_CUDA(cudaGetLastError());
  if (refine) {
    if (x_format == EVF_SAMPLES_S16)
      loudness_exact_kernel<short><<<n_utts, 32, 0, st>>>(static_cast<const short*>(x_dev), off, P, refine_flags_dev,
                                                          scratch_dev, soff);
    else
      loudness_exact_kernel<float><<<n_utts, 32, 0, st>>>(static_cast<const float*>(x_dev), off, P, refine_flags_dev,
                                                          scratch_dev, soff);
    EVF_CUDA(cudaGetLastError());
    loudness_gate_kernel<<<(n_utts + 63) / 64, 64, 0, st>>>(off, n_utts, P, scratch_dev, soff, lkfs_dev, 0.f, gate_lkfs,
                                                            nullptr, refine_flags_dev);
    EVF_CUDA(cudaGetLastError());
  }
  return EVF_OK;
}

}  // extern "C"
