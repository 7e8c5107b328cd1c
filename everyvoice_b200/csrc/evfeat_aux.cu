// Small kernels around the spectral path (sm_100a):
//   segment_mean   -- Preprocessor.average_data_by_durations, preprocessor/preprocessor.py:287-300
//   stats_partial  -- the reductions of Scaler.calculate_stats, preprocessor/helpers.py:86-106
//   normalize      -- Scaler.normalize, preprocessor/helpers.py:78-80
//   energy_from_spec -- Preprocessor.extract_energy, preprocessor/preprocessor.py:302-309
#include <math_constants.h>

#include "evfeat_internal.h"

namespace evf {

namespace {

// ---------------------------------------------------------------------------------------------
// Phone-level averaging: one warp per utterance, lane = phone.  The exclusive scan of the
// int64 durations is a warp shuffle scan with a running carry; each lane then averages its
// (Python-clipped) slice.  Slices of >= kCoopLen frames are handed back to the whole warp.
// ---------------------------------------------------------------------------------------------
constexpr int kCoopLen = 64;

__global__ void __launch_bounds__(256) segment_mean_kernel(
    const float* __restrict__ values, const long long* __restrict__ value_off,
    const long long* __restrict__ durations, const long long* __restrict__ phone_off, int n_utts,
    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int gwarp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * warps_per_block;
  for (int u = gwarp; u < n_utts; u += nwarps) {
    const long long v0 = value_off[u];
    const long long T = value_off[u + 1] - v0;
    const long long p0 = phone_off[u];
    const long long P = phone_off[u + 1] - p0;
    const float* vals = values + v0;
    long long carry = 0;  // current_frame_position at the start of this chunk of 32 phones
    for (long long base = 0; base < P; base += 32) {
      const long long idx = base + lane;
      const long long d = (idx < P) ? durations[p0 + idx] : 0;
      long long incl = d;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const long long start = carry + incl - d;
      carry += __shfl_sync(0xffffffffu, incl, 31);

      // data[start : start + d] with Python slice semantics on a length-T tensor
      long long lo = start, hi = start + d;
      if (lo < 0) { lo += T; if (lo < 0) lo = 0; } else if (lo > T) lo = T;
      if (hi < 0) { hi += T; if (hi < 0) hi = 0; } else if (hi > T) hi = T;
      const long long n = (d > 0 && hi > lo) ? hi - lo : 0;

      float result;
      if (d <= 0) {
        result = 1e-7f;
      } else if (n == 0) {
        result = CUDART_NAN_F;  // torch.mean of an empty slice
      } else if (n < kCoopLen) {
        float s = 0.f;
        for (long long i = 0; i < n; ++i) s += vals[lo + i];
        result = s / (float)n;
      } else {
        result = 0.f;  // filled in cooperatively below
      }
      unsigned big = __ballot_sync(0xffffffffu, idx < P && d > 0 && n >= kCoopLen);
      while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const long long blo = __shfl_sync(0xffffffffu, lo, src);
        const long long bn = __shfl_sync(0xffffffffu, n, src);
        float s = 0.f;
        for (long long i = lane; i < bn; i += 32) s += vals[blo + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == src) result = s / (float)bn;
      }
      if (idx < P) out[p0 + idx] = result;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// {count, sum, sumsq, min, max} over non-NaN values, float64 accumulation.
// ---------------------------------------------------------------------------------------------
__global__ void stats_init_kernel(double* out5) {
  out5[0] = 0.0;
  out5[1] = 0.0;
  out5[2] = 0.0;
  out5[3] = CUDART_INF;
  out5[4] = -CUDART_INF;
}

__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (v < __longlong_as_double((long long)old)) {
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (v > __longlong_as_double((long long)old)) {
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

__global__ void __launch_bounds__(256) stats_partial_kernel(const float* __restrict__ x,
                                                            long long n, double* out5) {
  double cnt = 0.0, sum = 0.0, sq = 0.0;
  float mn = CUDART_INF_F, mx = -CUDART_INF_F;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    if (v == v) {  // skip NaN (Scaler: non_nan_data / nanmean)
      const double dv = (double)v;
      cnt += 1.0;
      sum += dv;
      sq = fma(dv, dv, sq);
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ double s_cnt[8], s_sum[8], s_sq[8];
  __shared__ float s_mn[8], s_mx[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    s_cnt[w] = cnt; s_sum[w] = sum; s_sq[w] = sq; s_mn[w] = mn; s_mx[w] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
      cnt += s_cnt[i]; sum += s_sum[i]; sq += s_sq[i];
      mn = fminf(mn, s_mn[i]); mx = fmaxf(mx, s_mx[i]);
    }
    if (cnt > 0.0) {
      atomicAdd(out5 + 0, cnt);
      atomicAdd(out5 + 1, sum);
      atomicAdd(out5 + 2, sq);
      atomic_min_double(out5 + 3, (double)mn);
      atomic_max_double(out5 + 4, (double)mx);
    }
  }
}

__global__ void __launch_bounds__(256) normalize_kernel(float* __restrict__ x, long long n,
                                                        float mean, float std) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    x[i] = (x[i] - mean) / std;  // IEEE division, same two operations as the reference
}

// mean / std derived on the device from {count, sum, sumsq, ...} (no host round trip):
// mean = sum / n, std = sqrt((sumsq - sum^2 / n) / (n - 1)) (unbiased, like torch.std),
// both rounded to float32 before use, as the reference's float32 Scaler holds them.
__global__ void __launch_bounds__(256) normalize_by_stats_kernel(float* __restrict__ x, long long n,
                                                                 const double* __restrict__ stats5) {
  const double cnt = stats5[0], sum = stats5[1], sq = stats5[2];
  const double mean_d = sum / cnt;
  const double var_d = (sq - sum * sum / cnt) / (cnt - 1.0);
  const float mean = (float)mean_d;
  const float std = (float)sqrt(var_d > 0.0 ? var_d : 0.0);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    x[i] = (x[i] - mean) / std;
}

// Same, from the all-gathered summaries of n_parts ranks (parts[r * stride + 0..4]): every block
// merges the <= a few dozen numbers itself (SUM of count/sum/sumsq; min/max are not needed for
// the normalisation), so the all-gather's output feeds the normalisation with no kernel between.
__global__ void __launch_bounds__(256) normalize_by_gathered_kernel(float* __restrict__ x, long long n,
                                                                    const double* __restrict__ parts,
                                                                    int n_parts, int stride) {
  double cnt = 0.0, sum = 0.0, sq = 0.0;
  for (int r = 0; r < n_parts; ++r) {  // fixed order: every rank derives bit-identical mean / std
    cnt += parts[(long long)r * stride + 0];
    sum += parts[(long long)r * stride + 1];
    sq += parts[(long long)r * stride + 2];
  }
  const double mean_d = sum / cnt;
  const double var_d = (sq - sum * sum / cnt) / (cnt - 1.0);
  const float mean = (float)mean_d;
  const float std = (float)sqrt(var_d > 0.0 ? var_d : 0.0);
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
    x[i] = (x[i] - mean) / std;
}

// out5 = merge of n_parts five-number summaries (what an all-reduce with SUM/SUM/SUM/MIN/MAX gives).
__global__ void stats_merge_kernel(const double* __restrict__ parts, int n_parts, int stride,
                                   double* __restrict__ out5) {
  double cnt = 0.0, sum = 0.0, sq = 0.0, mn = CUDART_INF, mx = -CUDART_INF;
  for (int r = 0; r < n_parts; ++r) {
    const double* p = parts + (long long)r * stride;
    cnt += p[0];
    sum += p[1];
    sq += p[2];
    mn = fmin(mn, p[3]);
    mx = fmax(mx, p[4]);
  }
  out5[0] = cnt;
  out5[1] = sum;
  out5[2] = sq;
  out5[3] = mn;
  out5[4] = mx;
}

__global__ void __launch_bounds__(256) log_compress_kernel(const float* __restrict__ in,
                                                           float* __restrict__ out, long long n,
                                                           float c, float clip) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = in[i];
    v = (v < clip) ? clip : v;  // torch.clamp(min=clip): NaN propagates
    out[i] = logf(v * c);
  }
}

__global__ void __launch_bounds__(256) energy_kernel(const float* __restrict__ spec,
                                                     long long n_frames, int row,
                                                     float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long f = gwarp; f < n_frames; f += nwarps) {
    const float* r = spec + f * row;
    float acc = 0.f;
    for (int m = lane; m < row; m += 32) {
      const float v = r[m];
      acc = fmaf(v, v, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[f] = sqrtf(acc);
  }
}

int grid_for(long long work_items, int per_block, int cap) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace

int launch_energy_from_spec(const float* spec, int64_t n_frames, int row, float* out, cudaStream_t s) {
  if (n_frames == 0) return EVF_OK;
  energy_kernel<<<grid_for(n_frames, 8, 148 * 16), 256, 0, s>>>(spec, n_frames, row, out);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

int launch_segment_mean(const float* values, const int64_t* value_off, const int64_t* durations,
                        const int64_t* phone_off, int n_utts, float* out, cudaStream_t s) {
  if (n_utts == 0) return EVF_OK;
  segment_mean_kernel<<<grid_for(n_utts, 8, 148 * 8), 256, 0, s>>>(
      values, reinterpret_cast<const long long*>(value_off),
      reinterpret_cast<const long long*>(durations), reinterpret_cast<const long long*>(phone_off),
      n_utts, out);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

int launch_stats_partial(const float* values, int64_t n, double* out5, int accumulate, cudaStream_t s) {
  if (!accumulate) {
    stats_init_kernel<<<1, 1, 0, s>>>(out5);
    EVF_CUDA(cudaGetLastError());
  }
  if (n > 0) {
    stats_partial_kernel<<<grid_for(n, 256 * 8, 148 * 4), 256, 0, s>>>(values, n, out5);
    EVF_CUDA(cudaGetLastError());
  }
  return EVF_OK;
}

int launch_normalize_by_stats(float* values, int64_t n, const double* stats5, cudaStream_t s) {
  if (n == 0) return EVF_OK;
  normalize_by_stats_kernel<<<grid_for(n, 256 * 4, 148 * 8), 256, 0, s>>>(values, n, stats5);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

int launch_normalize_by_gathered(float* values, int64_t n, const double* parts, int n_parts, int stride,
                                 cudaStream_t s) {
  if (n == 0) return EVF_OK;
  normalize_by_gathered_kernel<<<grid_for(n, 256 * 4, 148 * 8), 256, 0, s>>>(values, n, parts, n_parts, stride);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

int launch_stats_merge(const double* parts, int n_parts, int stride, double* out5, cudaStream_t s) {
  stats_merge_kernel<<<1, 1, 0, s>>>(parts, n_parts, stride, out5);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

// Pitch post-processing of Preprocessor.extract_pitch (preprocessor.py:278-285): pitch[pitch == 0] = NaN, the NaNs
// are filled by np.interp over the frame index (linear between the neighbouring voiced frames, constant beyond the
// first / last voiced frame), an utterance without any voiced frame becomes all zeros; float64 in (what pyworld
// returns), float32 out (torch.tensor(pitch).float()).  One thread per utterance: a pitch track has at most a few
// thousand frames and the fill is a sequential walk over runs; np.interp's arithmetic is mirrored operation by
// operation in double precision (slope * (x - x0) + y0 without contraction).
__global__ void __launch_bounds__(64) pitch_fill_unvoiced_kernel(const double* __restrict__ pitch,
                                                                 const long long* __restrict__ off, int n_utts,
                                                                 float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_utts) return;
  const double* x = pitch + off[b];
  float* y = out + off[b];
  const long long T = off[b + 1] - off[b];
  auto voiced = [](double v) { return v != 0.0 && v == v; };  // 0 -> NaN by the reference; a NaN from the tracker is a gap too
  long long prev = -1;  // last voiced frame seen
  for (long long t = 0; t < T; ++t) {
    const double v = x[t];
    if (!voiced(v)) continue;
    if (prev < 0) {
      for (long long u = 0; u < t; ++u) y[u] = (float)v;  // np.interp: left of the first point -> its value
    } else if (t - prev > 1) {
      const double y0 = x[prev], slope = __ddiv_rn(__dsub_rn(v, y0), (double)(t - prev));
      for (long long u = prev + 1; u < t; ++u) y[u] = (float)__dadd_rn(__dmul_rn(slope, (double)(u - prev)), y0);
    }
    y[t] = (float)v;
    prev = t;
  }
  if (prev < 0) {
    for (long long u = 0; u < T; ++u) y[u] = 0.f;  // ValueError branch: pitch-less sample -> zeros
  } else {
    const float last = (float)x[prev];
    for (long long u = prev + 1; u < T; ++u) y[u] = last;  // right of the last point -> its value
  }
}

int launch_pitch_fill_unvoiced(const double* pitch, const int64_t* offsets, int n_utts, float* out, cudaStream_t s) {
  if (n_utts == 0) return EVF_OK;
  pitch_fill_unvoiced_kernel<<<(n_utts + 63) / 64, 64, 0, s>>>(pitch, reinterpret_cast<const long long*>(offsets),
                                                              n_utts, out);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

// The loudness gate consumed on the device (preprocessor.py:177-186: skipped when the loudness is NaN or < -36): the
// per-utterance values of a skipped utterance become NaN, so that the NaN-skipping statistics leave it out and the
// host can drop it after the batch; keep[b] tells which ones stay.  One warp per utterance.
__global__ void __launch_bounds__(128) gate_mask_kernel(const float* __restrict__ lkfs, float gate,
                                                        float* __restrict__ values, const long long* __restrict__ off,
                                                        int n_utts, int* __restrict__ keep_out) {
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= n_utts) return;
  const float l = lkfs[b];
  const bool keep = !(l != l) && !(l < gate);
  const int lane = threadIdx.x & 31;
  if (keep_out != nullptr && lane == 0) keep_out[b] = keep ? 1 : 0;
  if (keep || values == nullptr) return;
  const float nan = __int_as_float(0x7fc00000);
  for (long long i = off[b] + lane; i < off[b + 1]; i += 32) values[i] = nan;
}

int launch_gate_mask(const float* lkfs, float gate, float* values, const int64_t* offsets, int n_utts, int* keep_out,
                     cudaStream_t s) {
  if (n_utts == 0) return EVF_OK;
  gate_mask_kernel<<<(n_utts + 3) / 4, 128, 0, s>>>(lkfs, gate, values, reinterpret_cast<const long long*>(offsets), n_utts,
                                                   keep_out);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

int launch_log_compress(const float* in, float* out, int64_t n, float c, float clip, cudaStream_t s) {
  if (n == 0) return EVF_OK;
  log_compress_kernel<<<grid_for(n, 256 * 4, 148 * 8), 256, 0, s>>>(in, out, n, c, clip);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

int launch_normalize(float* values, int64_t n, float mean, float std, cudaStream_t s) {
  if (n == 0) return EVF_OK;
  normalize_kernel<<<grid_for(n, 256 * 4, 148 * 8), 256, 0, s>>>(values, n, mean, std);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

}  // namespace evf
