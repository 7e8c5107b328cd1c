// Pitch tracking on the device (sm_100a): what Preprocessor.extract_pitch gets from pyworld,
//   everyvoice/preprocessor/preprocessor.py:257-277   pw.dio(x, sr, frame_period = hop / sr * 1000, speed = 4)
//                                                     pw.stonemask(x, f0, t, sr)
// (SURVEY.md section 8f, row N3).  pyworld wraps M. Morise's WORLD vocoder (C++: dio.cpp, stonemask.cpp,
// matlabfunctions.cpp); neither the wheel nor its sources are available offline, so the kernels restate WORLD's
// published algorithm -- PARITY UNPINNED (oracle/world_pitch.py is the CPU restatement they are tested against; it
// follows WORLD's FFT-domain formulation, the kernels the equivalent time-domain one).  Everything is float64, like
// WORLD.
//
//   dio_decimate_*_kernel decimate(): 9 reflected margin samples, 3rd-order Chebyshev IIR forwards and backwards
//                         (zero phase), every r-th sample.  One thread per 256-sample segment with a run-in as long
//                         as the filter's (measured) memory; dio_mean_kernel: mean of the decimated signal.
//   dio_lowcut_kernel     DC removal + the 50 Hz low-cut FIR (minus a normalised Hann window plus a unit impulse; WORLD
//                         multiplies spectra of a zero-padded FFT, which IS this linear convolution).
//   dio_band_kernel       one warp per (utterance, band): Nuttall low-pass FIR of the band, the four zero-crossing
//                         engines (negative / positive going, peaks, dips) streamed through WORLD's interp1 at the frame
//                         times, candidate = mean of the four interval-frequencies, score = their deviation.
//   dio_fix_kernel        best candidate per frame, then FixF0Contour's four steps (jump removal, short-section
//                         removal, forward / backward extension): sequential, one thread per utterance.
//   stonemask_kernel      one warp per frame: Blackman-windowed segment and its derivative window, their spectra at the
//                         (at most 8) harmonic bins by direct DFT, instantaneous-frequency refinement.
#include <cmath>
#include <vector>

#include "evfeat_internal.h"

namespace evf {
namespace {

constexpr double kPi = 3.1415926535897932384626433832795;
constexpr double kMaximumValue = 100000.0;
constexpr double kSafeGuard = 1e-12;
constexpr int kMaxBands = 16;
constexpr int kNFact = 9;

struct DioParams {
  int fs, r, n_bands, cutoff;   // cutoff = matlab_round(actual_fs / 50)
  double actual_fs, frame_period, f0_floor, f0_ceil, allowed_range;
  double boundary_f0[kMaxBands];
  int half[kMaxBands];          // matlab_round(actual_fs / boundary_f0 / 2)
  double dec_a[3], dec_b[2];
};

// per-utterance layout (host built, device copies): everything int64
struct PitchLayout {
  const long long* x_off;    // [n + 1] samples
  const long long* y_off;    // [n + 1] decimated samples (y_length = 1 + L / r)
  const long long* f_off;    // [n + 1] frames (f0_length)
};

__device__ __forceinline__ double sample_d(const float* x, long long i) { return (double)x[i]; }
__device__ __forceinline__ double sample_d(const short* x, long long i) { return (double)((float)x[i] * (1.0f / 32768.0f)); }

// decimate(): the 3rd-order IIR runs over the margin-extended signal forwards, then over the reversed result.  The
// recursion is sequential, but the filter forgets: its impulse response falls below 1e-18 of its peak within `runin`
// samples (host-measured from the coefficients), so a thread that starts `runin` samples early from a zero state
// reproduces the sequential recursion to the last bit or two of float64.  One thread per segment of kDecSeg samples.
constexpr int kDecSeg = 256;
template <typename SampleT>
__device__ __forceinline__ double dec_input(const SampleT* xs, long long n, long long i, double x_first, double x_last) {
  if (i < kNFact) return 2.0 * x_first - sample_d(xs, kNFact - i);
  if (i < kNFact + n) return sample_d(xs, i - kNFact);
  return 2.0 * x_last - sample_d(xs, n - 2 - (i - (kNFact + n)));
}

template <typename SampleT>
__global__ void __launch_bounds__(128) dio_decimate_fwd_kernel(const SampleT* __restrict__ x, PitchLayout lay, DioParams P,
                                                               int runin, double* __restrict__ fwd) {
  const int b = blockIdx.y;
  const SampleT* xs = x + lay.x_off[b];
  const long long n = lay.x_off[b + 1] - lay.x_off[b];
  const long long total = n + 2 * kNFact;
  const long long s0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kDecSeg;
  if (s0 >= total) return;
  double* f = fwd + lay.x_off[b] + 2ll * kNFact * b;
  const double a0 = P.dec_a[0], a1 = P.dec_a[1], a2 = P.dec_a[2], b0 = P.dec_b[0], b1 = P.dec_b[1];
  const double x_first = sample_d(xs, 0), x_last = sample_d(xs, n - 1);
  double w0 = 0.0, w1 = 0.0, w2 = 0.0;
  const long long start = s0 - runin > 0 ? s0 - runin : 0;
  const long long end = s0 + kDecSeg < total ? s0 + kDecSeg : total;
  for (long long i = start; i < end; ++i) {
    const double wt = dec_input(xs, n, i, x_first, x_last) + a0 * w0 + a1 * w1 + a2 * w2;
    if (i >= s0) f[i] = b0 * wt + b1 * w0 + b1 * w1 + b0 * w2;
    w2 = w1;
    w1 = w0;
    w0 = wt;
  }
}

__global__ void __launch_bounds__(128) dio_decimate_bwd_kernel(PitchLayout lay, DioParams P, int runin,
                                                               const double* __restrict__ fwd, double* __restrict__ y) {
  const int b = blockIdx.y;
  const long long n = lay.x_off[b + 1] - lay.x_off[b];
  const long long total = n + 2 * kNFact;
  const long long s0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kDecSeg;  // in the REVERSED sequence
  if (s0 >= total) return;
  const double* f = fwd + lay.x_off[b] + 2ll * kNFact * b;
  const long long y_len = lay.y_off[b + 1] - lay.y_off[b];
  double* yb = y + lay.y_off[b];
  const double a0 = P.dec_a[0], a1 = P.dec_a[1], a2 = P.dec_a[2], b0 = P.dec_b[0], b1 = P.dec_b[1];
  const long long nout = (n - 1) / P.r + 1;
  const long long nbeg = P.r - P.r * nout + n;
  double w0 = 0.0, w1 = 0.0, w2 = 0.0;
  const long long start = s0 - runin > 0 ? s0 - runin : 0;
  const long long end = s0 + kDecSeg < total ? s0 + kDecSeg : total;
  for (long long i = start; i < end; ++i) {
    const double wt = f[total - 1 - i] + a0 * w0 + a1 * w1 + a2 * w2;
    const double v = b0 * wt + b1 * w0 + b1 * w1 + b0 * w2;
    w2 = w1;
    w1 = w0;
    w0 = wt;
    if (i >= s0) {
      const long long j = total - 1 - i;         // position in the re-reversed result
      const long long idx = j - (kNFact - 1);    // decimate() reads tmp1[idx + kNFact - 1] for idx = nbeg, nbeg + r, ... < n + kNFact
      if (idx >= nbeg && idx < n + kNFact && (idx - nbeg) % P.r == 0) {
        const long long c = (idx - nbeg) / P.r;
        if (c < y_len) yb[c] = v;
      }
    }
  }
}

// r == 1 copy, zero fill past the decimated samples, and the mean (summed in index order, like WORLD's loop)
template <typename SampleT>
__global__ void __launch_bounds__(64) dio_mean_kernel(const SampleT* __restrict__ x, PitchLayout lay, int n_utts, DioParams P,
                                                      double* __restrict__ y, double* __restrict__ mean_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_utts) return;
  const SampleT* xs = x + lay.x_off[b];
  const long long n = lay.x_off[b + 1] - lay.x_off[b];
  const long long y_len = lay.y_off[b + 1] - lay.y_off[b];
  double* yb = y + lay.y_off[b];
  long long n_set;  // entries of y the decimation wrote
  if (P.r == 1) {
    for (long long i = 0; i < y_len; ++i) yb[i] = (i < n) ? sample_d(xs, i) : 0.0;
    n_set = y_len;
  } else {
    const long long nout = (n - 1) / P.r + 1;
    const long long nbeg = P.r - P.r * nout + n;
    n_set = (n + kNFact - nbeg + P.r - 1) / P.r;
    for (long long i = n_set; i < y_len; ++i) yb[i] = 0.0;
  }
  double sum = 0.0;
  for (long long i = 0; i < y_len; ++i) sum += yb[i];
  mean_out[b] = sum / (double)y_len;
}

// y_lc[j] = sum_k lc[k] * (y - mean)[j - (k - C)], j in [-C, y_len + C): stored at index j + C of the utterance's row
__global__ void __launch_bounds__(256) dio_lowcut_kernel(const double* __restrict__ y, PitchLayout lay, int C,
                                                         const double* __restrict__ lc, const double* __restrict__ mean,
                                                         double* __restrict__ ylc) {
  extern __shared__ double s_y[];  // [256 + 2 C]
  const int b = blockIdx.y;
  const long long y_len = lay.y_off[b + 1] - lay.y_off[b];
  const long long j0 = (long long)blockIdx.x * 256 - C;  // first output of this block
  if (j0 >= y_len + C) return;
  const double* yb = y + lay.y_off[b];
  const double m = mean[b];
  for (int e = threadIdx.x; e < 256 + 2 * C; e += 256) {
    const long long i = j0 - C + e;  // input index
    s_y[e] = (i >= 0 && i < y_len) ? yb[i] - m : 0.0;
  }
  __syncthreads();
  const long long j = j0 + threadIdx.x;
  if (j >= y_len + C) return;
  double acc = 0.0;
  // input index j - (k - C), k = 0 .. 2C  <->  tile element threadIdx.x + 2C - k
  for (int k = 0; k <= 2 * C; ++k) acc = fma(lc[k], s_y[threadIdx.x + 2 * C - k], acc);
  ylc[lay.y_off[b] + 2ll * C * b + (j + C)] = acc;
}

// One warp per (utterance, band).
constexpr int kBandWarps = 4;
__global__ void __launch_bounds__(kBandWarps * 32) dio_band_kernel(const double* __restrict__ ylc, PitchLayout lay, int n_utts,
                                                                   DioParams P, const double* __restrict__ nuttall,
                                                                   const int* __restrict__ nut_off,
                                                                   double* __restrict__ interp /* [n_bands][4][frames] per utt */,
                                                                   double* __restrict__ cand, double* __restrict__ score) {
  __shared__ double s_tile[kBandWarps][32 + 4 * 64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long job = (long long)blockIdx.x * kBandWarps + warp;
  if (job >= (long long)n_utts * P.n_bands) return;
  const int b = (int)(job / P.n_bands), band = (int)(job % P.n_bands);
  const long long y_len = lay.y_off[b + 1] - lay.y_off[b];
  const long long F = lay.f_off[b + 1] - lay.f_off[b];
  const int C = P.cutoff;
  const double* yl = ylc + lay.y_off[b] + 2ll * C * b + C;  // yl[j] = y_lc[j], j in [-C, y_len + C)
  const int h = P.half[band], taps = 4 * h;
  const double* nut = nuttall + nut_off[band];
  const double fs = P.actual_fs;
  double* tile = s_tile[warp];
  double* out4 = interp + ((long long)P.n_bands * lay.f_off[b] + (long long)band * F) * 4;  // [4][F]
  const double fp_ms = P.frame_period;

  // streaming state of the four zero-crossing engines (identical in every lane)
  long long n_edges[4] = {0, 0, 0, 0};
  double prev_fine[4], loc_prev[4], itv_prev[4], loc_last[4], itv_last[4];
  long long n_events[4] = {0, 0, 0, 0}, fptr[4] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 4; ++k) prev_fine[k] = loc_prev[k] = itv_prev[k] = loc_last[k] = itv_last[k] = 0.0;

  auto assign = [&](int kind, double x0, double y0, double x1, double y1, long long f_end) {
    // frames [fptr, f_end) take WORLD's interp1 on the segment (x0, y0) - (x1, y1); lanes share the frames
    for (long long f = fptr[kind] + lane; f < f_end; f += 32) {
      const double t = (double)f * fp_ms / 1000.0;
      out4[kind * F + f] = y0 + (t - x0) / (x1 - x0) * (y1 - y0);
    }
    fptr[kind] = f_end > fptr[kind] ? f_end : fptr[kind];
  };
  auto first_frame_at_or_after = [&](double loc) -> long long {  // smallest f with f * fp / 1000 >= loc
    long long f = (long long)(loc * 1000.0 / fp_ms);
    if (f < 0) f = 0;
    while (f > 0 && (double)(f - 1) * fp_ms / 1000.0 >= loc) --f;
    while ((double)f * fp_ms / 1000.0 < loc) ++f;
    return f;
  };
  auto on_edge = [&](int kind, double fine) {
    if (n_edges[kind] >= 1) {  // a new interval (event)
      const double itv = fs / (fine - prev_fine[kind]);
      const double loc = (prev_fine[kind] + fine) / 2.0 / fs;
      if (n_events[kind] >= 1) {
        long long f_end = first_frame_at_or_after(loc);  // frames with t < loc: histc puts t == loc into the next bin
        if (f_end > F) f_end = F;
        assign(kind, loc_last[kind], itv_last[kind], loc, itv, f_end);
      }
      loc_prev[kind] = loc_last[kind];
      itv_prev[kind] = itv_last[kind];
      loc_last[kind] = loc;
      itv_last[kind] = itv;
      ++n_events[kind];
    }
    prev_fine[kind] = fine;
    ++n_edges[kind];
  };

  double s_prev = 0.0;  // lane's filtered sample of the previous chunk
  // chunk c covers samples i0 .. i0 + 31; crossings are evaluated for positions p = i0 - 2 + lane (needs s_p, s_p+1, s_p+2)
  for (long long i0 = 0; i0 < y_len + 2; i0 += 32) {
    // stage y_lc[i0 + 2h - taps + 1 .. i0 + 31 + 2h] (= 32 + taps - 1 values)
    const long long base = i0 + 2 * h - taps + 1;
    __syncwarp();
    for (int e = lane; e < 32 + taps - 1; e += 32) {
      const long long j = base + e;
      tile[e] = (j >= -C && j < y_len + C) ? yl[j] : 0.0;
    }
    __syncwarp();
    double s_cur = 0.0;
    {
      // s_i = sum_m nut[m] * y_lc[i + 2h - m]; tile element of y_lc[i + 2h - m] is (lane + taps - 1 - m)
      const double* tp = tile + lane + taps - 1;
      for (int m = 0; m < taps; ++m) s_cur = fma(nut[m], tp[-m], s_cur);
    }
    const long long p = i0 - 2 + lane;
    // s_p, s_{p+1}, s_{p+2}
    const double v_m2 = __shfl_sync(0xffffffffu, s_prev, (lane + 30) & 31);  // prev chunk lane + 30 (for lane < 2)
    const double c_m2 = __shfl_sync(0xffffffffu, s_cur, (lane + 30) & 31);   // cur chunk lane - 2 (for lane >= 2)
    const double sp0 = (lane < 2) ? v_m2 : c_m2;
    const double v_m1 = __shfl_sync(0xffffffffu, s_prev, (lane + 31) & 31);
    const double c_m1 = __shfl_sync(0xffffffffu, s_cur, (lane + 31) & 31);
    const double sp1 = (lane < 1) ? v_m1 : c_m1;
    const double sp2 = s_cur;
    const double d0 = sp0 - sp1, d1 = sp1 - sp2;
    const bool in0 = p >= 0 && p <= y_len - 2, in2 = p >= 0 && p <= y_len - 3;
    const bool fl[4] = {in0 && sp0 > 0.0 && sp1 <= 0.0, in0 && -sp0 > 0.0 && -sp1 <= 0.0,
                        in2 && d0 > 0.0 && d1 <= 0.0, in2 && -d0 > 0.0 && -d1 <= 0.0};
    const double fine_s = (fl[0] || fl[1]) ? (double)(p + 1) - sp0 / (sp1 - sp0) : 0.0;
    const double fine_d = (fl[2] || fl[3]) ? (double)(p + 1) - d0 / (d1 - d0) : 0.0;
#pragma unroll
    for (int kind = 0; kind < 4; ++kind) {
      unsigned mask = __ballot_sync(0xffffffffu, fl[kind]);
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const double fine = __shfl_sync(0xffffffffu, kind < 2 ? fine_s : fine_d, src);
        on_edge(kind, fine);
      }
    }
    s_prev = s_cur;
  }
  // frames at or beyond the last event: the last segment, extrapolated; then candidate and score
  bool ok = true;
#pragma unroll
  for (int kind = 0; kind < 4; ++kind) {
    if (n_events[kind] >= 2) assign(kind, loc_prev[kind], itv_prev[kind], loc_last[kind], itv_last[kind], F);
    ok = ok && (n_events[kind] - 2 > 0);  // CheckEvent(n - 2) for every kind
  }
  __syncwarp();
  double* cb = cand + (long long)P.n_bands * lay.f_off[b] + (long long)band * F;
  double* sb = score + (long long)P.n_bands * lay.f_off[b] + (long long)band * F;
  const double bf = P.boundary_f0[band];
  for (long long f = lane; f < F; f += 32) {
    double c = 0.0, sc = kMaximumValue;
    if (ok) {
      const double a0 = out4[f], a1 = out4[F + f], a2 = out4[2 * F + f], a3 = out4[3 * F + f];
      c = (a0 + a1 + a2 + a3) / 4.0;
      sc = sqrt(((a0 - c) * (a0 - c) + (a1 - c) * (a1 - c) + (a2 - c) * (a2 - c) + (a3 - c) * (a3 - c)) / 3.0);
      if (c > bf || c < bf / 2.0 || c > P.f0_ceil || c < P.f0_floor) {
        c = 0.0;
        sc = kMaximumValue;
      }
    }
    cb[f] = c;
    sb[f] = sc;
  }
}

__device__ double select_best_f0(double current_f0, double past_f0, const double* cand, long long F, int n_bands,
                                 long long target, double allowed_range) {
  const double reference_f0 = (current_f0 * 3.0 - past_f0) / 2.0;
  double minimum_error = fabs(reference_f0 - cand[target]);
  double best_f0 = cand[target];
  for (int i = 1; i < n_bands; ++i) {
    const double e = fabs(reference_f0 - cand[(long long)i * F + target]);
    if (e < minimum_error) {
      minimum_error = e;
      best_f0 = cand[(long long)i * F + target];
    }
  }
  if (fabs(1.0 - best_f0 / reference_f0) > allowed_range) return 0.0;
  return best_f0;
}

__global__ void __launch_bounds__(64) dio_fix_kernel(PitchLayout lay, int n_utts, DioParams P, const double* __restrict__ cand,
                                                     const double* __restrict__ score, double* __restrict__ tmp /* [3][frames] per utt */,
                                                     int* __restrict__ itmp /* [2][frames] per utt */, double* __restrict__ f0_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_utts) return;
  const long long F = lay.f_off[b + 1] - lay.f_off[b];
  const double* cb = cand + (long long)P.n_bands * lay.f_off[b];
  const double* sb = score + (long long)P.n_bands * lay.f_off[b];
  double* best = tmp + 3 * lay.f_off[b];
  double* s1 = best + F;
  double* s2 = s1 + F;
  int* positive = itmp + 2 * lay.f_off[b];
  int* negative = positive + F;
  double* out = f0_out + lay.f_off[b];
  for (long long i = 0; i < F; ++i) {  // GetBestF0Contour
    double t = sb[i], v = cb[i];
    for (int j = 1; j < P.n_bands; ++j)
      if (t > sb[(long long)j * F + i]) {
        t = sb[(long long)j * F + i];
        v = cb[(long long)j * F + i];
      }
    best[i] = v;
    out[i] = 0.0;
  }
  const long long vrm = (long long)(0.5 + 1000.0 / P.frame_period / P.f0_floor) * 2 + 1;  // voice_range_minimum
  if (F <= vrm) return;
  // step 1: f0_base keeps the contour away from both ends; frames that jump by more than allowed_range go
  auto base = [&](long long i) { return (i >= vrm && i < F - vrm) ? best[i] : 0.0; };
  for (long long i = 0; i < F; ++i)
    s1[i] = (i >= vrm && fabs((base(i) - base(i - 1)) / (kSafeGuard + base(i))) < P.allowed_range) ? base(i) : 0.0;
  // step 2: voiced sections shorter than the minimum go
  const long long center = (vrm - 1) / 2;
  for (long long i = 0; i < F; ++i) {
    double v = s1[i];
    if (i >= center && i < F - center)
      for (long long j = -center; j <= center; ++j)
        if (s1[i + j] == 0.0) {
          v = 0.0;
          break;
        }
    s2[i] = v;
  }
  int n_pos = 0, n_neg = 0;
  for (long long i = 1; i < F; ++i) {
    if (s2[i] == 0.0 && s2[i - 1] != 0.0) negative[n_neg++] = (int)(i - 1);
    else if (s2[i - 1] == 0.0 && s2[i] != 0.0) positive[n_pos++] = (int)i;
  }
  // step 3: every section grows forwards while a candidate continues it (in place in s2)
  for (int i = 0; i < n_neg; ++i) {
    const long long limit = (i == n_neg - 1) ? F - 1 : negative[i + 1];
    for (long long j = negative[i]; j < limit; ++j) {
      s2[j + 1] = select_best_f0(s2[j], s2[j - 1], cb, F, P.n_bands, j + 1, P.allowed_range);
      if (s2[j + 1] == 0.0) break;
    }
  }
  // step 4: ... and backwards
  for (int i = n_pos - 1; i >= 0; --i) {
    const long long limit = (i == 0) ? 1 : positive[i - 1];
    for (long long j = positive[i]; j > limit; --j) {
      s2[j - 1] = select_best_f0(s2[j], s2[j + 1], cb, F, P.n_bands, j - 1, P.allowed_range);
      if (s2[j - 1] == 0.0) break;
    }
  }
  for (long long i = 0; i < F; ++i) out[i] = s2[i];
}

__device__ __forceinline__ int matlab_round_d(double x) { return x > 0.0 ? (int)(x + 0.5) : (int)(x - 0.5); }

// StoneMask, one warp per frame.
template <typename SampleT>
__global__ void __launch_bounds__(128) stonemask_kernel(const SampleT* __restrict__ x, PitchLayout lay, int n_utts, int fs_i,
                                                        double frame_period, const long long* __restrict__ frame_utt,
                                                        long long total_frames, double* __restrict__ f0 /* in / out */) {
  const long long g = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (g >= total_frames) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)frame_utt[g];
  const long long f = g - lay.f_off[b];
  const double initial_f0 = f0[g];
  const double fs = (double)fs_i;
  __syncwarp();
  if (initial_f0 <= 40.0 || initial_f0 > fs / 12.0) {  // kFloorF0StoneMask
    if (lane == 0) f0[g] = 0.0;
    return;
  }
  const SampleT* xs = x + lay.x_off[b];
  const long long x_len = lay.x_off[b + 1] - lay.x_off[b];
  const double current_position = (double)f * frame_period / 1000.0;
  const int half = (int)(1.5 * fs / initial_f0 + 1.0);
  const double wlen = (2.0 * half + 1.0) / fs;  // window_length_in_time
  const int n = 2 * half + 1;
  const int fft_size = 1 << (2 + (int)(log(half * 2.0 + 1.0) / 0.69314718055994530942));
  const double base_time0 = (double)(-half) / fs;
  const int basic_index = matlab_round_d((current_position + base_time0) * fs + 0.001);
  auto main_w = [&](int i) {  // Blackman window at sample index_raw = basic_index + i
    const double t = ((double)(basic_index + i) - 1.0) / fs - current_position;
    return 0.42 + 0.5 * cos(2.0 * kPi * t / wlen) + 0.08 * cos(4.0 * kPi * t / wlen);
  };
  auto seg = [&](int i) {
    long long idx = (long long)basic_index + i - 1;
    idx = idx < 0 ? 0 : (idx > x_len - 1 ? x_len - 1 : idx);
    return sample_d(xs, idx);
  };
  // spectra of (segment * main window) and (segment * diff window) at a list of bins, by direct DFT: each lane takes
  // a contiguous share of the samples and rotates its phasor (one sincos per lane and bin)
  const int per = (n + 31) / 32;
  const int i_lo = lane * per, i_hi = min(n, i_lo + per);
  auto spectra = [&](const int* bins, int n_bins, double* mr, double* mi, double* dr, double* di) {
    for (int q = 0; q < n_bins; ++q) mr[q] = mi[q] = dr[q] = di[q] = 0.0;
    double w_prev = (i_lo > 0 && i_lo <= n) ? main_w(i_lo - 1) : 0.0, w_cur = (i_lo < n) ? main_w(i_lo) : 0.0;
    double cr[8], ci[8], rr[8], ri[8];
    for (int q = 0; q < n_bins; ++q) {
      const double th = -2.0 * kPi * (double)bins[q] / (double)fft_size;
      sincos(th * (double)i_lo, &ci[q], &cr[q]);
      sincos(th, &ri[q], &rr[q]);
    }
    for (int i = i_lo; i < i_hi; ++i) {
      const double w_next = (i + 1 < n) ? main_w(i + 1) : 0.0;
      double dw;
      if (i == 0) dw = -w_next / 2.0;
      else if (i == n - 1) dw = w_prev / 2.0;
      else dw = -(w_next - w_prev) / 2.0;
      const double s = seg(i);
      const double vm = s * w_cur, vd = s * dw;
      for (int q = 0; q < n_bins; ++q) {
        mr[q] = fma(vm, cr[q], mr[q]);
        mi[q] = fma(vm, ci[q], mi[q]);
        dr[q] = fma(vd, cr[q], dr[q]);
        di[q] = fma(vd, ci[q], di[q]);
        const double nr = cr[q] * rr[q] - ci[q] * ri[q];
        ci[q] = cr[q] * ri[q] + ci[q] * rr[q];
        cr[q] = nr;
      }
      w_prev = w_cur;
      w_cur = w_next;
    }
    for (int q = 0; q < n_bins; ++q)
      for (int o = 16; o > 0; o >>= 1) {
        mr[q] += __shfl_xor_sync(0xffffffffu, mr[q], o);
        mi[q] += __shfl_xor_sync(0xffffffffu, mi[q], o);
        dr[q] += __shfl_xor_sync(0xffffffffu, dr[q], o);
        di[q] += __shfl_xor_sync(0xffffffffu, di[q], o);
      }
  };
  auto fix_f0 = [&](double f_init, int n_harm) -> double {
    int bins[8];
    double mr[8], mi[8], dr[8], di[8];
    for (int i = 0; i < n_harm; ++i) {
      int idx = matlab_round_d(f_init * fft_size / fs * (i + 1));
      bins[i] = idx > fft_size / 2 ? fft_size / 2 : idx;
    }
    spectra(bins, n_harm, mr, mi, dr, di);
    double num = 0.0, den = 0.0;
    for (int i = 0; i < n_harm; ++i) {
      const double power = mr[i] * mr[i] + mi[i] * mi[i];
      const double numerator_i = mr[i] * di[i] - mi[i] * dr[i];
      const double inst = power == 0.0 ? 0.0 : (double)bins[i] * fs / fft_size + numerator_i / power * fs / 2.0 / kPi;
      const double amp = sqrt(power);
      num += amp * inst;
      den += amp * (i + 1.0);
    }
    return num / (den + kSafeGuard);
  };
  const double tentative = fix_f0(initial_f0, 2);
  double mean_f0 = 0.0;
  if (!(tentative <= 0.0 || tentative > initial_f0 * 2.0)) {
    int n_harm = (int)(fs / 2.0 / tentative);
    n_harm = n_harm < 6 ? n_harm : 6;
    mean_f0 = fix_f0(tentative, n_harm);
  }
  if (fabs(mean_f0 - initial_f0) > initial_f0 * 0.2) mean_f0 = initial_f0;
  if (lane == 0) f0[g] = mean_f0;
}

// cheby1(3, 0.05, 0.8 / r): WORLD's FilterForDecimate table, a = {-a1, -a2, -a3}, b = {b0, b1}
const double kDecimateCoeffs[13][5] = {
    {0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0},
    {0.04115673456775716, -0.4259911245918959, 0.04103721547996115, 0.1679746468180222, 0.5039239404540666},
    {0.9503937898323742, -0.674291467415268, 0.15412211621346472, 0.07122194517117862, 0.21366583551353585},
    {1.4499664446880223, -0.9894349708095054, 0.245782523406902, 0.03671075033932264, 0.11013225101796792},
    {1.761093965428056, -1.255491484385977, 0.32371865077882145, 0.02133485852238745, 0.06400457556716235},
    {1.971535274951214, -1.4686795689225343, 0.38939084349657005, 0.013469181309343806, 0.04040754392803142},
    {2.12252390195347, -1.6395144861046296, 0.44469707800587344, 0.009036688268160781, 0.027110064804482345},
    {2.2357462340187593, -1.7780899984041356, 0.491525553659687, 0.006352276340711179, 0.01905682902213354},
    {2.323600349175958, -1.89215456174636, 0.5314892813372907, 0.004633116404138924, 0.013899349212416773},
    {2.3936475118069382, -1.9873904075111852, 0.5658879979027052, 0.0034818622251927374, 0.010445586675578211},
    {2.450743295230728, -2.0679490460197805, 0.5957477443833211, 0.002682250800716404, 0.008046752402149212},
    {2.4981398605924205, -2.1368928194784025, 0.6218751381622148, 0.002109727590470877, 0.006329182771412631},
};

int matlab_round_h(double x) { return x > 0 ? (int)(x + 0.5) : (int)(x - 0.5); }

}  // namespace
}  // namespace evf

using namespace evf;

extern "C" {

int64_t evf_pitch_num_frames(int32_t sample_rate, double frame_period_ms, int64_t n_samples) {
  if (sample_rate < 1 || !(frame_period_ms > 0.0) || n_samples < 0) return -1;
  return (int64_t)(1000.0 * (double)n_samples / (double)sample_rate / frame_period_ms) + 1;  // GetSamplesForDIO
}

int64_t evf_pitch_scratch_bytes(const int64_t* offsets_host, int32_t n_utts, int32_t sample_rate, double frame_period_ms,
                                int32_t speed) {
  if (!offsets_host || n_utts < 0 || sample_rate < 1 || !(frame_period_ms > 0.0)) return -1;
  const int r = speed < 1 ? 1 : (speed > 12 ? 12 : speed);
  const double actual_fs = (double)sample_rate / r;
  const int C = matlab_round_h(actual_fs / 50.0);
  long long xs = 0, ys = 0, fr = 0;
  for (int b = 0; b < n_utts; ++b) {
    const long long L = offsets_host[b + 1] - offsets_host[b];
    xs += L + 2 * kNFact;
    ys += 1 + L / r;
    fr += evf_pitch_num_frames(sample_rate, frame_period_ms, L);
  }
  const long long nb = kMaxBands;
  long long bytes = 0;
  bytes += 8 * xs;                         // forward pass of the decimation filter
  bytes += 8 * ys;                         // decimated signal
  bytes += 8 * (ys + 2ll * C * n_utts);    // low-cut filtered signal with its margins
  bytes += 8 * nb * fr * 4;                // interpolated interval-frequencies [bands][4][frames]
  bytes += 8 * nb * fr * 2;                // candidates and scores
  bytes += 8 * fr * 3 + 4 * fr * 2;        // contour fixing
  bytes += 8 * fr;                         // frame -> utterance map
  bytes += 8 * (3ll * (n_utts + 1) + n_utts);  // layout tables, means
  bytes += 8 * (2 * C + 1 + 4 * 64 * nb + nb) + 4096;
  return bytes;
}

int evf_pitch_dio_stonemask(const void* x_dev, int32_t x_format, const int64_t* offsets_host, int32_t n_utts,
                            int32_t sample_rate, double frame_period_ms, int32_t speed, double f0_floor, double f0_ceil,
                            double channels_in_octave, double allowed_range, void* scratch_dev, int64_t scratch_bytes,
                            double* f0_out_dev, void* stream) {
  if (n_utts < 0 || sample_rate < 1 || !(frame_period_ms > 0.0) || !(f0_floor > 0.0) || !(f0_ceil > f0_floor) ||
      !(channels_in_octave > 0.0) || (x_format != EVF_SAMPLES_F32 && x_format != EVF_SAMPLES_S16) ||
      (n_utts > 0 && (!x_dev || !offsets_host || !scratch_dev || !f0_out_dev))) {
    set_error("evf_pitch_dio_stonemask: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (n_utts == 0) return EVF_OK;
  DioParams P{};
  P.fs = sample_rate;
  P.r = speed < 1 ? 1 : (speed > 12 ? 12 : speed);
  P.actual_fs = (double)sample_rate / P.r;
  P.frame_period = frame_period_ms;
  P.f0_floor = f0_floor;
  P.f0_ceil = f0_ceil;
  P.allowed_range = allowed_range;
  P.n_bands = 1 + (int)(std::log(f0_ceil / f0_floor) / 0.69314718055994530942 * channels_in_octave);
  if (P.n_bands > kMaxBands || P.n_bands < 1) {
    set_error("evf_pitch_dio_stonemask: more than 16 bands");
    return EVF_ERR_UNSUPPORTED;
  }
  P.cutoff = matlab_round_h(P.actual_fs / 50.0);
  std::vector<double> nut;
  std::vector<int> nut_off(kMaxBands, 0);
  for (int i = 0; i < P.n_bands; ++i) {
    P.boundary_f0[i] = f0_floor * std::pow(2.0, (i + 1) / channels_in_octave);
    P.half[i] = matlab_round_h(P.actual_fs / P.boundary_f0[i] / 2.0);
    if (P.half[i] < 1 || P.half[i] > 64) {
      set_error("evf_pitch_dio_stonemask: band filter length outside [4, 256] taps (sampling rate / speed / f0 range)");
      return EVF_ERR_UNSUPPORTED;
    }
    nut_off[i] = (int)nut.size();
    const int len = 4 * P.half[i];
    for (int m = 0; m < len; ++m) {  // NuttallWindow
      const double t = m / (len - 1.0);
      nut.push_back(0.355768 - 0.487396 * std::cos(2.0 * kPi * t) + 0.144232 * std::cos(4.0 * kPi * t) -
                    0.012604 * std::cos(6.0 * kPi * t));
    }
  }
  for (int k = 0; k < 3; ++k) P.dec_a[k] = kDecimateCoeffs[P.r][k];
  P.dec_b[0] = kDecimateCoeffs[P.r][3];
  P.dec_b[1] = kDecimateCoeffs[P.r][4];
  const int C = P.cutoff;
  std::vector<double> lc(2 * C + 1);  // DesignLowCutFilter, centred
  {
    const int N = 2 * C + 1;
    double sum = 0.0;
    for (int i = 1; i <= N; ++i) {
      lc[i - 1] = 0.5 - 0.5 * std::cos(i * 2.0 * kPi / (N + 1));
      sum += lc[i - 1];
    }
    for (int i = 0; i < N; ++i) lc[i] = -lc[i] / sum;
    lc[C] += 1.0;
  }
  // ---- layout -------------------------------------------------------------------------------------------------
  std::vector<long long> tab(3 * (size_t)(n_utts + 1));
  long long* x_off = tab.data();
  long long* y_off = x_off + (n_utts + 1);
  long long* f_off = y_off + (n_utts + 1);
  x_off[0] = offsets_host[0];
  y_off[0] = f_off[0] = 0;
  long long max_y = 0;
  for (int b = 0; b < n_utts; ++b) {
    const long long L = offsets_host[b + 1] - offsets_host[b];
    if (L < 2 * kNFact + 2) {
      set_error("evf_pitch_dio_stonemask: an utterance is shorter than 20 samples");
      return EVF_ERR_SHORT_INPUT;
    }
    x_off[b + 1] = offsets_host[b + 1];
    y_off[b + 1] = y_off[b] + 1 + L / P.r;
    f_off[b + 1] = f_off[b] + evf_pitch_num_frames(sample_rate, frame_period_ms, L);
    max_y = std::max(max_y, 1 + L / P.r);
  }
  const long long xs = (x_off[n_utts] - x_off[0]) + 2ll * kNFact * n_utts, ys = y_off[n_utts], fr = f_off[n_utts];
  if (evf_pitch_scratch_bytes(offsets_host, n_utts, sample_rate, frame_period_ms, speed) > scratch_bytes) {
    set_error("evf_pitch_dio_stonemask: scratch buffer smaller than evf_pitch_scratch_bytes");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  std::vector<long long> frame_utt((size_t)fr);
  for (int b = 0; b < n_utts; ++b)
    for (long long f = f_off[b]; f < f_off[b + 1]; ++f) frame_utt[(size_t)f] = b;
  // carve the scratch buffer (8-byte units)
  double* base = static_cast<double*>(scratch_dev);
  size_t at = 0;
  auto take = [&](size_t n_doubles) {
    double* p = base + at;
    at += n_doubles;
    return p;
  };
  double* d_fwd = take((size_t)xs);
  double* d_y = take((size_t)ys);
  double* d_ylc = take((size_t)(ys + 2ll * C * n_utts));
  double* d_interp = take((size_t)P.n_bands * fr * 4);
  double* d_cand = take((size_t)P.n_bands * fr);
  double* d_score = take((size_t)P.n_bands * fr);
  double* d_tmp = take((size_t)fr * 3);
  int* d_itmp = reinterpret_cast<int*>(take((size_t)fr));  // 2 ints per frame
  long long* d_frame_utt = reinterpret_cast<long long*>(take((size_t)fr));
  long long* d_tab = reinterpret_cast<long long*>(take(tab.size()));
  double* d_mean = take((size_t)n_utts);
  double* d_lc = take(lc.size());
  double* d_nut = take(nut.size());
  int* d_nut_off = reinterpret_cast<int*>(take(kMaxBands));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the x offsets are absolute into x_dev; rebase them so that x_off[0] may be non-zero
  EVF_CUDA(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, st));
  EVF_CUDA(cudaMemcpyAsync(d_frame_utt, frame_utt.data(), frame_utt.size() * 8, cudaMemcpyHostToDevice, st));
  EVF_CUDA(cudaMemcpyAsync(d_lc, lc.data(), lc.size() * 8, cudaMemcpyHostToDevice, st));
  EVF_CUDA(cudaMemcpyAsync(d_nut, nut.data(), nut.size() * 8, cudaMemcpyHostToDevice, st));
  EVF_CUDA(cudaMemcpyAsync(d_nut_off, nut_off.data(), kMaxBands * 4, cudaMemcpyHostToDevice, st));
  EVF_CUDA(cudaStreamSynchronize(st));  // the host vectors above go out of scope
  PitchLayout lay{d_tab, d_tab + (n_utts + 1), d_tab + 2 * (n_utts + 1)};
  // fwd rows are addressed as x_off[b] + 18 b: rebase so that utterance 0 starts at d_fwd
  double* fwd_base = d_fwd - x_off[0];
  const bool s16 = x_format == EVF_SAMPLES_S16;
  if (P.r > 1) {
    // how long the decimation filter remembers: impulse response below 1e-18 of its peak (measured, not assumed)
    int runin = 64;
    {
      double w0 = 0, w1 = 0, w2 = 0, peak = 0;
      int last_big = 0;
      for (int i = 0; i < 20000; ++i) {
        const double wt = (i == 0 ? 1.0 : 0.0) + P.dec_a[0] * w0 + P.dec_a[1] * w1 + P.dec_a[2] * w2;
        const double v = std::fabs(P.dec_b[0] * wt + P.dec_b[1] * w0 + P.dec_b[1] * w1 + P.dec_b[0] * w2);
        w2 = w1;
        w1 = w0;
        w0 = wt;
        peak = v > peak ? v : peak;
        if (v > 1e-18 * peak) last_big = i;
      }
      runin = last_big + 16;
    }
    long long max_total = 0;
    for (int b = 0; b < n_utts; ++b) max_total = std::max(max_total, x_off[b + 1] - x_off[b] + 2 * kNFact);
    const long long segs = (max_total + kDecSeg - 1) / kDecSeg;
    for (int y0 = 0; y0 < n_utts; y0 += 65535) {
      const int ny = n_utts - y0 < 65535 ? n_utts - y0 : 65535;
      PitchLayout l2{lay.x_off + y0, lay.y_off + y0, lay.f_off + y0};
      const dim3 grid((unsigned)((segs + 127) / 128), (unsigned)ny);
      double* fb = fwd_base + 2ll * kNFact * y0;
      if (s16) dio_decimate_fwd_kernel<short><<<grid, 128, 0, st>>>(static_cast<const short*>(x_dev), l2, P, runin, fb);
      else dio_decimate_fwd_kernel<float><<<grid, 128, 0, st>>>(static_cast<const float*>(x_dev), l2, P, runin, fb);
      EVF_CUDA(cudaGetLastError());
      dio_decimate_bwd_kernel<<<grid, 128, 0, st>>>(l2, P, runin, fb, d_y);
      EVF_CUDA(cudaGetLastError());
    }
  }
  if (s16)
    dio_mean_kernel<short><<<(n_utts + 63) / 64, 64, 0, st>>>(static_cast<const short*>(x_dev), lay, n_utts, P, d_y, d_mean);
  else
    dio_mean_kernel<float><<<(n_utts + 63) / 64, 64, 0, st>>>(static_cast<const float*>(x_dev), lay, n_utts, P, d_y, d_mean);
  EVF_CUDA(cudaGetLastError());
  {
    const size_t smem = (256 + 2 * (size_t)C) * sizeof(double);
    if (smem > 48 * 1024) EVF_CUDA(cudaFuncSetAttribute(dio_lowcut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int y0 = 0; y0 < n_utts; y0 += 65535) {
      const int ny = n_utts - y0 < 65535 ? n_utts - y0 : 65535;
      PitchLayout l2{lay.x_off + y0, lay.y_off + y0, lay.f_off + y0};
      const dim3 grid((unsigned)((max_y + 2 * C + 255) / 256), (unsigned)ny);
      // rows are addressed with the utterance index: shift the bases by the slice
      dio_lowcut_kernel<<<grid, 256, smem, st>>>(d_y, l2, C, d_lc, d_mean + y0, d_ylc + 2ll * C * y0);
      EVF_CUDA(cudaGetLastError());
    }
  }
  {
    const long long jobs = (long long)n_utts * P.n_bands;
    dio_band_kernel<<<(unsigned)((jobs + kBandWarps - 1) / kBandWarps), kBandWarps * 32, 0, st>>>(d_ylc, lay, n_utts, P, d_nut, d_nut_off,
                                                                                                  d_interp, d_cand, d_score);
    EVF_CUDA(cudaGetLastError());
  }
  dio_fix_kernel<<<(n_utts + 63) / 64, 64, 0, st>>>(lay, n_utts, P, d_cand, d_score, d_tmp, d_itmp, f0_out_dev);
  EVF_CUDA(cudaGetLastError());
  if (fr > 0) {
    const unsigned blocks = (unsigned)((fr + 3) / 4);
    if (s16)
      stonemask_kernel<short><<<blocks, 128, 0, st>>>(static_cast<const short*>(x_dev), lay, n_utts, sample_rate, frame_period_ms,
                                                      d_frame_utt, fr, f0_out_dev);
    else
      stonemask_kernel<float><<<blocks, 128, 0, st>>>(static_cast<const float*>(x_dev), lay, n_utts, sample_rate, frame_period_ms,
                                                      d_frame_utt, fr, f0_out_dev);
    EVF_CUDA(cudaGetLastError());
  }
  return EVF_OK;
}

}  // extern "C"
