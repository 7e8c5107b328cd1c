// n_fft = R * 1024 (R = 3, 4) through the 1024-point warp kernel (sm_100a): the "output" transform of a vocoder
// configuration with a sampling-rate change, whose n_fft / win / hop are the base sizes times R
//   everyvoice/preprocessor/preprocessor.py:94-121   (16 -> 48 kHz: 3072 / 768, x4: 4096 / 1024)
//   everyvoice/utils/heavy.py:47-113                 (the transform itself)
//
// Decimation in time by R: with xpad the reflect-padded signal (n_fft / 2 on both sides, torch.stft(center=True)) and
// the phase streams y_r[i] = xpad[R i + r], frame t of the R*1024-point transform is
//   X_t[k] = sum_r W_N^(r k) Y_{r,t}[k mod 1024],   Y_{r,t} = DFT_1024( w[R m + r] * y_r[t * hop / R + m] ),
// i.e. R ordinary 1024-point short-time transforms (hop / R, no padding, window w_r[m] = w[R m + r]) of the phase
// streams, followed by a radix-R butterfly per bin.  Three steps per chunk of utterances:
//   1. deinterleave_kernel: reflect padding + the R phase streams, float32 (int16 PCM converted as s / 32768);
//   2. features_kernel<MODE_PACK2, raw> (evfeat_features.cu, unchanged) once per stream: two frames ride as real and
//      imaginary part of one complex FFT, the separated half spectra Y_r[0..512] go to a scratch;
//   3. combine_kernel (here), one warp per frame: lane k' forms B_r = W_N^(r k') Y_r[k'] and the R-point DFT D_q of
//      (B_r); real input makes Y_r[1024 - k'] = conj(Y_r[k']), so D_q also yields the mirrored bins:
//        R = 4:  X[k'] = D_0, X[1024 + k'] = D_1, X[2048 - k'] = conj(D_2), X[1024 - k'] = conj(D_3)
//        R = 3:  X[k'] = D_0, X[1024 + k'] = D_1, X[1024 - k'] = conj(D_2)
//      then |X|^2 (sqrt(. + 1e-9) for mel-librosa) -> mel projection (lane = filter walks the filter's rising and
//      falling bins) -> log(max(., clip)) -> coalesced stores + the frame's energy.
// The scratch (R * 4 KB per frame) is allocated stream-ordered per chunk by the caller (evfeat_api.cu).
#include "evfeat_fft.cuh"
#include "evfeat_internal.h"

namespace evf {

namespace {

#include "evfeat_device.cuh"

template <typename SampleT>
__global__ void __launch_bounds__(256) deinterleave_kernel(const SampleT* __restrict__ x,
                                                           const long long* __restrict__ sample_off,
                                                           const long long* __restrict__ stream_off, long long chunk_soff0,
                                                           long long plane_stride, int R, int n_fft,
                                                           float* __restrict__ planes) {
  const int b = blockIdx.y;
  const long long s0 = sample_off[b];
  const int L = (int)(sample_off[b + 1] - s0);
  const long long o0 = stream_off[b] - chunk_soff0;
  const int Ls = (int)(stream_off[b + 1] - stream_off[b]);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ls) return;
  const int half = n_fft / 2;
  for (int r = 0; r < R; ++r) {
    const int j = R * i + r - half;  // position in the utterance; outside [0, L): reflect padding
    float v = 0.f;
    if (j < L + half) v = to_float(__ldg(x + s0 + reflect_index(j, L)));
    planes[r * plane_stride + o0 + i] = v;
  }
}

struct DecParams {
  const float* raw;         // [R][n_frames][1026]: Y_r[0..512] of every frame as (re, im)
  long long plane_stride;   // floats between the R planes
  long long n_frames;
  float* spec_out;          // rows of the chunk's frames
  float* energy_out;        // or NULL
  const float2* wcomb;      // [513] W_N^k'
  const float2* melw;       // [k_used] {rising, falling} weight of every bin
  const int* kstart;        // [n_mels + 2]
  int n_mels, n_freq, k_used, row_floats, apply_log;
  float log_clip;
};

constexpr int kDecWarps = 8;

__device__ __forceinline__ float2 c_mul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 c_conj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 c_mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }

// Latency is what matters here (a frame's R half spectra are 16 KB read once from the scratch): a warp keeps the
// loads of four rows of bins (4 x R x 256 B) in flight, the mel weights sit in shared memory, and the power column
// holds only the bins that carry a mel weight, so that 40 to 64 warps fit on an SM.
template <int R, int SPEC>
__global__ void __launch_bounds__(kDecWarps * 32) combine_kernel(const DecParams p) {
  constexpr bool kMel = (SPEC == EVF_SPEC_MEL || SPEC == EVF_SPEC_MEL_LIBROSA);
  extern __shared__ __align__(16) float dec_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kcap = kMel ? p.k_used : p.n_freq;          // bins >= kcap are not consumed
  const int col = (kcap + 3) & ~3;
  float2* s_melw = reinterpret_cast<float2*>(dec_smem);  // [k_used]
  int* s_kstart = reinterpret_cast<int*>(s_melw + (kMel ? p.k_used : 0));  // [n_mels + 2]
  float* P = reinterpret_cast<float*>(s_kstart + (kMel ? ((p.n_mels + 2 + 3) & ~3) : 0)) + (size_t)warp * col;
  if constexpr (kMel) {
    for (int i = threadIdx.x; i < p.k_used; i += kDecWarps * 32) s_melw[i] = p.melw[i];
    for (int i = threadIdx.x; i < p.n_mels + 2; i += kDecWarps * 32) s_kstart[i] = p.kstart[i];
    __syncthreads();
  }
  const long long ps = p.plane_stride / 2;  // float2 elements between the planes
  for (long long f = (long long)blockIdx.x * kDecWarps + warp; f < p.n_frames; f += (long long)gridDim.x * kDecWarps) {
    float* ga = p.spec_out + f * p.row_floats;
    float esum = 0.f;
    // one bin of the output: complex (raw), log-power (linear) or power into the column (mel types)
    auto emit = [&](int k, float2 X) {
      if (k >= kcap) return;
      if constexpr (SPEC == EVF_SPEC_RAW) {
        reinterpret_cast<float2*>(ga)[k] = X;
      } else {
        float pw = fmaf(X.x, X.x, X.y * X.y);
        if constexpr (SPEC == EVF_SPEC_MEL_LIBROSA) pw = fast_sqrt(pw + 1e-9f);
        if constexpr (kMel) {
          P[k] = pw;
        } else {
          const float v = compress(pw, p.apply_log, p.log_clip);
          ga[k] = v;
          esum = fmaf(v, v, esum);
        }
      }
    };
    // bin k' of the R half spectra (A[r]) -> the output bins it determines
    auto butterfly = [&](int k, const float2 (&A)[R], float2 t) {
      const float2 B0 = A[0];
      const float2 B1 = c_mul(A[1], t);
      const float2 t2 = c_mul(t, t);
      const float2 B2 = c_mul(A[2], t2);
      const bool inner = (k > 0 && k < 512);  // the bins whose mirror images are not produced a second time
      if constexpr (R == 4) {
        const float2 B3 = c_mul(A[3], c_mul(t2, t));
        const float2 s02 = c_add(B0, B2), d02 = c_sub(B0, B2);
        const float2 s13 = c_add(B1, B3), d13 = c_mul_neg_i(c_sub(B1, B3));
        emit(k, c_add(s02, s13));                              // D_0
        emit(1024 + k, c_add(d02, d13));                       // D_1
        if (k < 512) emit(2048 - k, c_conj(c_sub(s02, s13)));  // conj(D_2)
        if (inner) emit(1024 - k, c_conj(c_sub(d02, d13)));    // conj(D_3)
      } else {
        constexpr float c = 0.86602540378443864676f;           // sin(2 pi / 3)
        const float2 s = c_add(B1, B2), d = c_sub(B1, B2);
        const float2 m = make_float2(fmaf(-0.5f, s.x, B0.x), fmaf(-0.5f, s.y, B0.y));
        const float2 n = make_float2(c * d.y, -c * d.x);       // -i * c * d
        emit(k, c_add(B0, s));                                 // D_0
        emit(1024 + k, c_add(m, n));                           // D_1
        if (inner) emit(1024 - k, c_conj(c_sub(m, n)));        // conj(D_2)
      }
    };
    const float2* src = reinterpret_cast<const float2*>(p.raw + f * 1026) + lane;
#pragma unroll 1
    for (int j0 = 0; j0 < 16; j0 += 4) {
      float2 A[4][R], t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        t[u] = __ldg(p.wcomb + lane + 32 * (j0 + u));
#pragma unroll
        for (int r = 0; r < R; ++r) A[u][r] = __ldg(src + 32 * (j0 + u) + r * ps);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) butterfly(lane + 32 * (j0 + u), A[u], t[u]);
    }
    if (lane == 0) {  // bin 512
      float2 A[R];
#pragma unroll
      for (int r = 0; r < R; ++r) A[r] = __ldg(src + 512 + r * ps);
      butterfly(512, A, __ldg(p.wcomb + 512));
    }
    if constexpr (kMel) {
      __syncwarp();
      for (int m = lane; m < p.n_mels; m += 32) {
        const int k0 = s_kstart[m], k1 = s_kstart[m + 1], k2 = s_kstart[m + 2];
        float acc = 0.f;
        for (int k = k0; k < k1; ++k) acc = fmaf(s_melw[k].x, P[k], acc);  // rising side: interval m
        for (int k = k1; k < k2; ++k) acc = fmaf(s_melw[k].y, P[k], acc);  // falling side: interval m + 1
        const float v = compress(acc, p.apply_log, p.log_clip);
        ga[m] = v;
        esum = fmaf(v, v, esum);
      }
      __syncwarp();  // the column is rewritten by the warp's next frame
    }
    if constexpr (SPEC != EVF_SPEC_RAW) {
      if (p.energy_out != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
        if (lane == 0) p.energy_out[f] = sqrtf(esum);
      }
    }
  }
}

template <int R>
int launch_combine_r(int spec, const DecParams& p, int grid, int smem, cudaStream_t st) {
#define EVF_LAUNCH_DEC(SPEC)                                                                        \
  {                                                                                                 \
    auto k = combine_kernel<R, SPEC>;                                                               \
    if (smem > 48 * 1024) EVF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    k<<<grid, kDecWarps * 32, smem, st>>>(p);                                                       \
  }
  switch (spec) {
    case EVF_SPEC_MEL: EVF_LAUNCH_DEC(EVF_SPEC_MEL) break;
    case EVF_SPEC_MEL_LIBROSA: EVF_LAUNCH_DEC(EVF_SPEC_MEL_LIBROSA) break;
    case EVF_SPEC_LINEAR: EVF_LAUNCH_DEC(EVF_SPEC_LINEAR) break;
    case EVF_SPEC_RAW: EVF_LAUNCH_DEC(EVF_SPEC_RAW) break;
    default: set_error("unknown spec_type"); return EVF_ERR_UNSUPPORTED;
  }
#undef EVF_LAUNCH_DEC
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

}  // namespace

int decimated_deinterleave(const void* samples, int sample_format, const long long* sample_off_dev,
                           const long long* stream_off_dev, int utt0, int n_utts, long long chunk_soff0,
                           long long plane_stride, long long max_stream_len, int R, int n_fft, float* planes,
                           cudaStream_t st) {
  if (n_utts <= 0 || max_stream_len <= 0) return EVF_OK;
  for (int y0 = 0; y0 < n_utts; y0 += 65535) {
    const int ny = n_utts - y0 < 65535 ? n_utts - y0 : 65535;
    const dim3 grid((unsigned)((max_stream_len + 255) / 256), (unsigned)ny);
    if (sample_format == EVF_SAMPLES_S16)
      deinterleave_kernel<short><<<grid, 256, 0, st>>>(static_cast<const short*>(samples), sample_off_dev + utt0 + y0,
                                                      stream_off_dev + utt0 + y0, chunk_soff0, plane_stride, R, n_fft,
                                                      planes);
    else
      deinterleave_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(samples), sample_off_dev + utt0 + y0,
                                                      stream_off_dev + utt0 + y0, chunk_soff0, plane_stride, R, n_fft,
                                                      planes);
    EVF_CUDA(cudaGetLastError());
  }
  return EVF_OK;
}

int decimated_combine(int R, int spec_type, const float* raw, long long plane_stride, long long n_frames,
                      float* spec_out, float* energy_out, const float2* wcomb, const float2* melw, const int* kstart,
                      int n_mels, int n_freq, int k_used, int row_floats, int apply_log, float log_clip, int num_sms,
                      cudaStream_t st) {
  if (n_frames <= 0) return EVF_OK;
  DecParams p{};
  p.raw = raw;
  p.plane_stride = plane_stride;
  p.n_frames = n_frames;
  p.spec_out = spec_out;
  p.energy_out = (spec_type == EVF_SPEC_RAW) ? nullptr : energy_out;
  p.wcomb = wcomb;
  p.melw = melw;
  p.kstart = kstart;
  p.n_mels = n_mels;
  p.n_freq = n_freq;
  p.k_used = k_used;
  p.row_floats = row_floats;
  p.apply_log = (spec_type == EVF_SPEC_RAW) ? 0 : apply_log;
  p.log_clip = log_clip;
  const bool mel = (spec_type == EVF_SPEC_MEL || spec_type == EVF_SPEC_MEL_LIBROSA);
  // mel types: weights and interval starts once per block, one power column of the weighted bins per warp
  const int smem = mel ? (2 * k_used + ((n_mels + 2 + 3) & ~3) + kDecWarps * ((k_used + 3) & ~3)) * (int)sizeof(float) : 0;
  int per_sm = (227 * 1024) / (smem + 1024);
  if (per_sm > 2048 / (kDecWarps * 32)) per_sm = 2048 / (kDecWarps * 32);
  if (per_sm < 1) per_sm = 1;
  long long grid = (n_frames + kDecWarps - 1) / kDecWarps;
  const long long cap = (long long)num_sms * per_sm;
  if (grid > cap) grid = cap;
  return R == 4 ? launch_combine_r<4>(spec_type, p, (int)grid, smem, st)
                : launch_combine_r<3>(spec_type, p, (int)grid, smem, st);
}

}  // namespace evf
