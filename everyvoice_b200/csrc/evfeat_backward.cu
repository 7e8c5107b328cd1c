// Backward of the spectral transform (sm_100a): d loss / d audio from d loss / d spectrogram, for the training-time
// use of get_spectral_transform: HiFiGAN recomputes the mel spectrogram of the
// GENERATED audio every step and back-propagates through it,
//   everyvoice/model/vocoder/HiFiGAN_iSTFT_lightning/hfgl/model.py:581-590, 719-721, 812-814
//   everyvoice/utils/heavy.py:47-113 (the transform), :39-40 (the log compression, evfeat_aux.cu)
//
// The forward is  x -> reflect pad -> frames -> window -> rFFT -> |X|^2 (-> sqrt(. + 1e-9)) -> mel basis.  Its
// transpose, per FFT job (two frames ride as real and imaginary part of one complex FFT, like the forward kernel):
//   1. recompute X_a, X_b of the two frames (same loads, window-fused butterflies, 32x32 four-step FFT, real-FFT
//      separation as evfeat_features.cu);
//   2. g_P[k] = rise[k] * g_mel[j(k)] + fall[k] * g_mel[j(k) - 1]   (each bin feeds two adjacent filters; the
//      transposed mel projection needs no reduction at all),  G[k] = dL/dX[k] = 2 g_P[k] X[k];
//   3. dL/d(windowed frame)[n] = Re sum_{k=0}^{N/2} G[k] e^{+2 pi i k n / N}: the one-sided spectra of both frames are
//      Hermitian-extended and packed as C = C_a + i C_b, whose unnormalised inverse DFT is r_a + i r_b; the inverse
//      runs through the SAME forward FFT code with real and imaginary parts swapped on the way in and out;
//   4. times the window -> one row of 1024 frame gradients per frame (scratch);
// then overlap_add_kernel gathers, for every sample, the <= n_fft / hop frame rows that cover it, plus the rows that
// cover its mirror images in the reflect padding.  No atomics; the result is deterministic.
#include "evfeat_fft.cuh"
#include "evfeat_internal.h"

namespace evf {

namespace {

#include "evfeat_device.cuh"

constexpr int kBwdWarps = 16;
constexpr int kGmStride = 128;  // n_mels <= 128 gradient values per frame in shared memory

// MODE_PACK2 (n_fft 1024): two frames per job as real / imaginary part.  MODE_HALF (n_fft 2048): one frame per job as
// 1024 complex points z[m] = v[2m] + i v[2m + 1]; the transpose of its real-FFT split is the classic half-size
// inverse: with Y = G / 2 (Y_0 = Re G_0, Y_1024 = Re G_1024),  A[k] = Y[k] + conj(Y[M - k]),
// B[k] = (Y[k] - conj(Y[M - k])) e^{+2 pi i k / 2048},  Zin[k] = A[k] + i B[k] (k < M = 1024; Zin[M - k] =
// conj(A[k]) + i conj(B[k])), and the unnormalised inverse DFT of Zin is  g_v[2m] + i g_v[2m + 1].
template <int MODE, int SPEC>
__global__ void __launch_bounds__(kBwdWarps * 32, 1) features_backward_kernel(const BwdParams p) {
  constexpr bool kHalf = (MODE == MODE_HALF);
  constexpr int NFFT = kHalf ? 2048 : 1024;
  constexpr int FPJ = kHalf ? 1 : 2;
  constexpr bool kMel = (SPEC == EVF_SPEC_MEL || SPEC == EVF_SPEC_MEL_LIBROSA);
  extern __shared__ __align__(16) float smem[];
  float* s_win = smem;                                              // forward layout of the mode, pre-scaled by 1/2
  float4* s_tw4 = reinterpret_cast<float4*>(smem + NFFT);           // [16][32]
  float2* s_wpost = reinterpret_cast<float2*>(smem + NFFT + 2 * kFftSize);   // MODE_HALF: (cos, -sin)(2 pi k / 2048)
  float* s_warp = smem + NFFT + 2 * kFftSize + (kHalf ? 1028 : 0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* scr = s_warp + warp * (32 * kScrStride + 2 * kGmStride);
  float* gm = scr + 32 * kScrStride;                                 // [2][kGmStride] mel gradients of frames a, b

  for (int i = tid; i < NFFT; i += kBwdWarps * 32) s_win[i] = p.window[i];
  for (int i = tid; i < kFftSize / 2; i += kBwdWarps * 32) s_tw4[i] = p.tw4[i];
  if constexpr (kHalf)
    for (int i = tid; i <= 512; i += kBwdWarps * 32) s_wpost[i] = p.wpost[i];
  for (int i = tid; i < kBwdWarps * (32 * kScrStride + 2 * kGmStride); i += kBwdWarps * 32) s_warp[i] = 0.f;
  __syncthreads();

  const TileDesc ti = p.tiles[blockIdx.x];
  const int fa = FPJ * warp;
  if (fa >= ti.nvalid) return;
  const bool b_valid = !kHalf && (fa + 1 < ti.nvalid);
  const int hop = p.hop;
  const float* xs = p.samples + ti.s_off;
  const long long frame_a = ti.out_frame0 + fa;

  // ---- forward recomputation: samples -> window-fused first stage -> FFT ------------------------------
  float re[32], im[32];
  if constexpr (!kHalf) {
    const int ua = ti.start + fa * hop + lane, ub = ua + hop;
    const float2* wv = reinterpret_cast<const float2*>(s_win) + lane;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float2 w = wv[32 * r];
      const int i = bitrev5(r);
      win_head(re[i], re[i + 1], __ldg(xs + reflect_index(ua + 32 * r, ti.L)), w.x,
               __ldg(xs + reflect_index(ua + 32 * (r + 16), ti.L)), w.y);
      win_head(im[i], im[i + 1], __ldg(xs + reflect_index(ub + 32 * r, ti.L)), w.x,
               __ldg(xs + reflect_index(ub + 32 * (r + 16), ti.L)), w.y);
    }
  } else {
    const int u0 = ti.start + fa * hop + 2 * lane;
    const float2* w2 = reinterpret_cast<const float2*>(s_win) + lane;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float2 wa = w2[32 * r], wb = w2[32 * (r + 16)];
      const int ia = u0 + 64 * r, ib = u0 + 64 * (r + 16);
      const int i = bitrev5(r);
      win_head(re[i], re[i + 1], __ldg(xs + reflect_index(ia, ti.L)), wa.x, __ldg(xs + reflect_index(ib, ti.L)), wb.x);
      win_head(im[i], im[i + 1], __ldg(xs + reflect_index(ia + 1, ti.L)), wa.y,
               __ldg(xs + reflect_index(ib + 1, ti.L)), wb.y);
    }
  }
  warp_fft1024_tail(re, im, s_tw4, scr, lane);

  // ---- mel gradients of the frame(s) -> shared memory ----------------------------------------------------
  const float* ga = p.grad_spec + frame_a * p.row_floats;
  if constexpr (kMel) {
    for (int m = lane; m < p.n_mels; m += 32) {
      gm[m] = __ldg(ga + m);
      gm[kGmStride + m] = b_valid ? __ldg(ga + p.row_floats + m) : 0.f;
    }
    __syncwarp();
  }
  // transposed mel projection for bin k of frame f (0 = a, 1 = b): a bin feeds two adjacent filters
  auto g_power = [&](int k, int f) -> float {
    if constexpr (kMel) {
      if (k >= p.k_used) return 0.f;
      const float2 w = __ldg(p.melw + k);   // {rising weight -> filter j(k), falling weight -> filter j(k) - 1}
      const int jj = __ldg(p.jk + k);
      const int m_r = min(jj, p.n_mels - 1), m_f = max(jj - 1, 0);
      const float wr = (jj < p.n_mels) ? w.x : 0.f, wf = (jj >= 1) ? w.y : 0.f;
      return fmaf(wr, gm[f * kGmStride + m_r], wf * gm[f * kGmStride + m_f]);
    } else {
      return (f == 0 || b_valid) ? __ldg(ga + f * p.row_floats + k) : 0.f;
    }
  };
  // mel-librosa: mel = basis @ sqrt(P + 1e-9), dM/dP = 1 / (2 sqrt(P + 1e-9))
  auto librosa = [&](float g, float xr, float xi) -> float {
    if constexpr (SPEC == EVF_SPEC_MEL_LIBROSA) return g * 0.5f * rsqrtf(fmaf(xr, xr, xi * xi) + 1e-9f);
    return g;
  };

  // ---- spectrum gradients per bin, extension to the full complex input of the inverse ------------------------
  // element e = lane + 32 q; ck: this lane's elements k = lane + 32 j (j <= 15, and k = 512 for lane 0);
  // cm: the mirror elements 1024 - k, fetched below by the lane that owns them
  float ck_r[17], ck_i[17], cm_r[17], cm_i[17];
  const int src_lane = (32 - lane) & 31;
#pragma unroll
  for (int j = 0; j <= 16; ++j) {
    const int k = lane + 32 * j;
    float zr, zi, pr, pi;
    if (j < 16) {
      zr = re[j];
      zi = im[j];
      const float sr = (lane == 0) ? re[(32 - j) & 31] : re[31 - j];
      const float si = (lane == 0) ? im[(32 - j) & 31] : im[31 - j];
      pr = __shfl_sync(0xffffffffu, sr, src_lane);
      pi = __shfl_sync(0xffffffffu, si, src_lane);
    } else {
      zr = pr = re[16];
      zi = pi = im[16];
    }
    const bool own = (j < 16) || (lane == 0);
    if constexpr (!kHalf) {
      // window was pre-scaled by 1/2: X_a = Z[k] + conj(Z[N-k]), X_b = (Z[k] - conj(Z[N-k])) / i
      const float ar = zr + pr, ai = zi - pi;
      const float br = zi + pi, bi = pr - zr;
      const float gpa = own ? librosa(g_power(k, 0), ar, ai) : 0.f;
      const float gpb = own ? librosa(g_power(k, 1), br, bi) : 0.f;
      const float ur = gpa * ar, ui = gpa * ai, vr = gpb * br, vi = gpb * bi;  // u = G_a / 2, v = G_b / 2
      if (j == 16 || (j == 0 && lane == 0)) {
        // k = 0 and k = N/2: the spectrum is real there; C = Re G_a + i Re G_b, no mirror
        ck_r[j] = 2.f * ur;
        ck_i[j] = 2.f * vr;
        cm_r[j] = ck_r[j];  // (only read for j == 16 by lane 0: element 512 is its own mirror)
        cm_i[j] = ck_i[j];
      } else {
        ck_r[j] = ur - vi;  // u + i v
        ck_i[j] = ui + vr;
        cm_r[j] = ur + vi;  // conj(u) + i conj(v)
        cm_i[j] = vr - ui;
      }
    } else {
      // forward split: X[k] = E - T, X[M - k] = conj(E + T), E = Z[k] + conj(Z[M - k]),
      // T = i w_k (Z[k] - conj(Z[M - k])), w_k = e^{-2 pi i k / 2048}
      const float er = zr + pr, ei = zi - pi;
      const float orr = zr - pr, oi = zi + pi;
      const float2 w = s_wpost[own ? k : 0];
      const float tr = -fmaf(w.x, oi, w.y * orr);
      const float tq = fmaf(w.x, orr, -w.y * oi);
      const float x0r = er - tr, x0i = ei - tq;      // bin k
      const float x1r = er + tr, x1i = -(ei + tq);   // bin 1024 - k
      const float g0 = own ? librosa(g_power(k, 0), x0r, x0i) : 0.f;
      const float g1 = own ? librosa(g_power(1024 - k, 0), x1r, x1i) : 0.f;
      float y0r = g0 * x0r, y0i = g0 * x0i;          // Y[k]     = G[k] / 2
      float y1r = g1 * x1r, y1i = g1 * x1i;          // Y[M - k] = G[M - k] / 2
      if (j == 0 && lane == 0) {                     // bins 0 and 1024: real, not halved
        y0r *= 2.f;
        y0i = 0.f;
        y1r *= 2.f;
        y1i = 0.f;
      }
      if (j == 16) {                                 // k = 512 is its own mirror
        y1r = y0r;
        y1i = y0i;
      }
      const float Ar = y0r + y1r, Ai = y0i - y1i;     // A = Y[k] + conj(Y[M - k])
      const float dr = y0r - y1r, di = y0i + y1i;     // Y[k] - conj(Y[M - k])
      const float Br = fmaf(dr, w.x, di * w.y);       // times e^{+2 pi i k / 2048} = (w.x, -w.y)
      const float Bi = fmaf(di, w.x, -dr * w.y);
      ck_r[j] = Ar - Bi;  // A + i B
      ck_i[j] = Ai + Br;
      cm_r[j] = Ar + Bi;  // conj(A) + i conj(B)
      cm_i[j] = Br - Ai;
      if (j == 16) {      // element 512 (lane 0) is read through the mirror path below
        cm_r[j] = ck_r[j];
        cm_i[j] = ck_i[j];
      }
    }
  }

  // ---- inverse DFT through the forward FFT: swap real and imaginary parts on the way in and out -------------
  // position bitrev5(q) of the DIT arrays holds element q
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    re[bitrev5(q)] = ck_i[q];
    im[bitrev5(q)] = ck_r[q];
  }
#pragma unroll
  for (int q = 16; q < 32; ++q) {
    // element lane + 32 q > 512 (or == 512 for lane 0, q == 16): the mirror of element 1024 - e, owned by src_lane
    const float sr = (lane == 0) ? cm_r[(32 - q) & 31] : cm_r[31 - q];
    const float si = (lane == 0) ? cm_i[(32 - q) & 31] : cm_i[31 - q];
    re[bitrev5(q)] = __shfl_sync(0xffffffffu, si, src_lane);
    im[bitrev5(q)] = __shfl_sync(0xffffffffu, sr, src_lane);
  }
  dft32_dit_head(re, im);
  warp_fft1024_tail(re, im, s_tw4, scr, lane);
  // FFT(swap(C)) = swap(y): real part of the inverse = im, imaginary part = re; element n = lane + 32 k2

  // ---- times the (true) window -> frame-gradient rows --------------------------------------------------------
  float* fg_a = p.frame_grad + frame_a * NFFT;
  if constexpr (!kHalf) {
    const float2* wv = reinterpret_cast<const float2*>(s_win) + lane;
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
      const float2 wp = wv[32 * (k2 & 15)];
      const float w = 2.f * ((k2 < 16) ? wp.x : wp.y);
      fg_a[lane + 32 * k2] = w * im[k2];
      if (b_valid) fg_a[NFFT + lane + 32 * k2] = w * re[k2];
    }
  } else {
    const float2* w2 = reinterpret_cast<const float2*>(s_win) + lane;
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
      const float2 w = w2[32 * k2];  // {w[2m], w[2m + 1]} / 2, m = lane + 32 k2
      reinterpret_cast<float2*>(fg_a)[lane + 32 * k2] = make_float2(2.f * w.x * im[k2], 2.f * w.y * re[k2]);
    }
  }
}

// d loss / d x[s] = sum over the padded positions that read x[s] (itself and its mirror images in the reflect
// padding, torch.stft(center=True, pad_mode="reflect")) of the frame-gradient rows covering that position.
__global__ void __launch_bounds__(256) overlap_add_kernel(const float* __restrict__ frame_grad,
                                                          const long long* __restrict__ sample_off,
                                                          const long long* __restrict__ frame_off, int n_fft, int hop,
                                                          float* __restrict__ grad_x) {
  const int b = blockIdx.y;
  const long long s0 = sample_off[b], L = sample_off[b + 1] - s0;
  const long long f0 = frame_off[b], T = frame_off[b + 1] - f0;
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= L) return;
  const long long H = n_fft / 2;
  // the padded positions that read x[s]: itself, and its mirror images in the left / right reflect margin
  const long long pos[3] = {s + H, H - s, H + 2 * (L - 1) - s};
  const bool use[3] = {true, s >= 1 && s <= H, s <= L - 2 && s >= L - 1 - H};
  float g = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (!use[i]) continue;
    const long long pp = pos[i];
    long long t_hi = pp / hop;
    if (t_hi > T - 1) t_hi = T - 1;
    long long t_lo = (pp - n_fft + hop) / hop;  // ceil((pp - n_fft + 1) / hop) for pp - n_fft + 1 > 0
    if (pp - n_fft + 1 <= 0) t_lo = 0;
    for (long long t = t_lo; t <= t_hi; ++t) g += __ldg(frame_grad + (f0 + t) * n_fft + (pp - t * hop));
  }
  grad_x[s0 + s] = g;
}

// d/dx log(clamp(x, min = clip) * C) = 1 / x where the clamp passes (x >= clip), else 0   (utils/heavy.py:39-40)
__global__ void __launch_bounds__(256) log_compress_backward_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ g, float* __restrict__ out,
                                                                    long long n, float clip) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    out[i] = (v >= clip) ? g[i] / v : ((v != v) ? v : 0.f);
  }
}

template <int MODE, int SPEC>
int launch_bwd_t(const BwdParams& p, int smem, cudaStream_t st) {
  auto k = features_backward_kernel<MODE, SPEC>;
  EVF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k<<<p.n_tiles, kBwdWarps * 32, smem, st>>>(p);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}
template <int SPEC>
int launch_bwd(const BwdParams& p, int smem, cudaStream_t st) {
  return p.n_fft == 2048 ? launch_bwd_t<MODE_HALF, SPEC>(p, smem, st) : launch_bwd_t<MODE_PACK2, SPEC>(p, smem, st);
}

}  // namespace

int features_backward_launch(const BwdParams& p, cudaStream_t st) {
  if (p.n_tiles == 0 || p.n_utts == 0) return EVF_OK;
  if (p.n_mels > kGmStride) {
    set_error("evf_features_backward: more than 128 mel filters are not supported");
    return EVF_ERR_UNSUPPORTED;
  }
  const int smem = (p.n_fft + 2 * kFftSize + (p.n_fft == 2048 ? 1028 : 0) +
                    kBwdWarps * (32 * kScrStride + 2 * kGmStride)) * (int)sizeof(float);
  int rc;
  switch (p.spec_type) {
    case EVF_SPEC_MEL: rc = launch_bwd<EVF_SPEC_MEL>(p, smem, st); break;
    case EVF_SPEC_MEL_LIBROSA: rc = launch_bwd<EVF_SPEC_MEL_LIBROSA>(p, smem, st); break;
    case EVF_SPEC_LINEAR: rc = launch_bwd<EVF_SPEC_LINEAR>(p, smem, st); break;
    default:
      set_error("evf_features_backward: spec_type has no backward (raw is complex)");
      return EVF_ERR_UNSUPPORTED;
  }
  if (rc != EVF_OK) return rc;
  return overlap_add_launch(p.frame_grad, p.sample_off, p.frame_off, p.n_utts, p.max_len, p.n_fft, p.hop, p.grad_samples, st);
}

int overlap_add_launch(const float* frame_grad, const long long* sample_off, const long long* frame_off, int n_utts,
                       long long max_len, int n_fft, int hop, float* grad_samples, cudaStream_t st) {
  if (max_len <= 0) return EVF_OK;
  for (int y0 = 0; y0 < n_utts; y0 += 65535) {  // gridDim.y limit: slices of utterances
    const dim3 grid((unsigned)((max_len + 255) / 256), (unsigned)(n_utts - y0 < 65535 ? n_utts - y0 : 65535));
    overlap_add_kernel<<<grid, 256, 0, st>>>(frame_grad, sample_off + y0, frame_off + y0, n_fft, hop, grad_samples);
    EVF_CUDA(cudaGetLastError());
  }
  return EVF_OK;
}

int launch_log_compress_backward(const float* x, const float* g, float* out, int64_t n, float clip, cudaStream_t s) {
  if (n == 0) return EVF_OK;
  long long blocks = (n + 256 * 4 - 1) / (256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  log_compress_backward_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, g, out, n, clip);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

}  // namespace evf
