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
//   4. times the window -> one row of n_fft frame gradients per frame.
// The rows of a tile (32 / 16 consecutive frames of one utterance) overlap; they are summed INSIDE the tile, in shared
// memory, in rounds of frames that do not overlap (frame t takes part in round t mod ceil(n_fft / hop); a barrier
// between rounds: no atomics, a fixed order), and only the tile's span of (frames - 1) * hop + n_fft sums goes to the
// scratch (3.7x less than the rows for hop = n_fft / 4).  tile_overlap_add_kernel then gathers, for every sample, the
// <= 2 tile sums that cover it, plus those that cover its mirror images in the reflect padding.  Deterministic.
// Hops so large or so small that the tile's span does not fit beside the FFT scratch (or that need more than 16
// rounds) keep the per-frame rows and overlap_add_kernel (kTileSum = false).
#include "evfeat_fft.cuh"
#include "evfeat_internal.h"

namespace evf {

namespace {

#include "evfeat_device.cuh"

// A CTA of 8 warps handles HALF a forward tile (16 / 8 frames); two CTAs share an SM, so that one's barrier rounds and
// copy-out overlap the other's FFTs.
constexpr int kBwdWarps = 8;
constexpr int kBwdSplit = kMaxWarps / kBwdWarps;  // backward CTAs per forward tile
static_assert(kBwdSplit >= 1 && kBwdSplit * kBwdWarps == kMaxWarps, "a forward tile is a whole number of backward CTAs");
constexpr int kGmStride = 128;  // n_mels <= 128 gradient values per frame in shared memory

// MODE_PACK2 (n_fft 1024): two frames per job as real / imaginary part.  MODE_HALF (n_fft 2048): one frame per job as
// 1024 complex points z[m] = v[2m] + i v[2m + 1]; the transpose of its real-FFT split is the classic half-size
// inverse: with Y = G / 2 (Y_0 = Re G_0, Y_1024 = Re G_1024),  A[k] = Y[k] + conj(Y[M - k]),
// B[k] = (Y[k] - conj(Y[M - k])) e^{+2 pi i k / 2048},  Zin[k] = A[k] + i B[k] (k < M = 1024; Zin[M - k] =
// conj(A[k]) + i conj(B[k])), and the unnormalised inverse DFT of Zin is  g_v[2m] + i g_v[2m + 1].
template <int MODE, int SPEC, bool kTileSum>
__global__ void __launch_bounds__(kBwdWarps * 32, 2) features_backward_kernel(const BwdParams p) {
  using MT = ModeTraits<MODE>;
  constexpr bool kHalf = (MODE == MODE_HALF);
  constexpr int NFFT = MT::kNfft;
  constexpr int FPJ = MT::kFramesPerJob;
  constexpr int J = MT::kJobs;      // n_fft 512: two packed jobs side by side in the warp's register file (evfeat_features.cu)
  constexpr int R1 = 32 / J;        // first-pass rows of a job = lanes that own a job's bins / samples after a transform
  constexpr int FPW = FPJ * J;      // frames of a warp
  constexpr int kWarpWords = 32 * kScrStride + 2 * FPW * kGmStride;  // transpose scratch + gradients / log values of the frames
  static_assert(J == 1 || J == 2, "the register re-layout between the two transforms is written for one or two jobs");
  constexpr bool kMel = (SPEC == EVF_SPEC_MEL || SPEC == EVF_SPEC_MEL_LIBROSA);
  extern __shared__ __align__(16) float smem[];
  float* s_win = smem;                                              // forward layout of the mode, pre-scaled by 1/2
  float4* s_tw4 = reinterpret_cast<float4*>(smem + NFFT);           // [16][32]
  float2* s_wpost = reinterpret_cast<float2*>(smem + NFFT + 2 * kFftSize);   // MODE_HALF: (cos, -sin)(2 pi k / 2048)
  float* s_warp = smem + NFFT + 2 * kFftSize + (kHalf ? 1028 : 0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* scr = s_warp + warp * kWarpWords;
  float* gm = scr + 32 * kScrStride;                                 // [FPW][kGmStride] mel gradients of the warp's frames (+ as many log values)
  float* s_acc = s_warp + kBwdWarps * kWarpWords;                    // kTileSum: the tile's span of summed rows

  for (int i = tid; i < NFFT; i += kBwdWarps * 32) s_win[i] = p.window[i];
  for (int i = tid; i < kFftSize / 2; i += kBwdWarps * 32) s_tw4[i] = p.tw4[i];
  if constexpr (kHalf)
    for (int i = tid; i <= 512; i += kBwdWarps * 32) s_wpost[i] = p.wpost[i];
  for (int i = tid; i < kBwdWarps * kWarpWords; i += kBwdWarps * 32) s_warp[i] = 0.f;
  TileDesc ti = p.tiles[blockIdx.x / kBwdSplit];
  const int hop = p.hop;
  {  // this CTA's part of the forward tile: frames [half * FPT, half * FPT + FPT)
    constexpr int FPT = FPW * kBwdWarps;
    const int half = blockIdx.x % kBwdSplit;
    ti.nvalid -= half * FPT;
    if (ti.nvalid <= 0) return;
    if (ti.nvalid > FPT) ti.nvalid = FPT;
    ti.start += half * FPT * hop;
    ti.out_frame0 += half * FPT;
  }
  const int span_valid = (ti.nvalid - 1) * hop + NFFT;  // padded positions the CTA's frames cover
  if constexpr (kTileSum)
    for (int i = tid; i < span_valid; i += kBwdWarps * 32) s_acc[i] = 0.f;
  __syncthreads();

  const int fw = FPW * warp;            // first frame of the warp (tile-relative)
  const bool a_valid = fw < ti.nvalid;  // warp-uniform: the warp has work
  const int k1 = lane & (R1 - 1);       // the lane inside its job's R1 lanes
  const int jw = lane / R1;             // the lane's job (J > 1)
  const int fa = fw + FPJ * jw;         // first frame of the lane's job
  const bool la_valid = (J == 1) ? a_valid : (fa < ti.nvalid);
  const GradView gv(p.gs, ti, NFFT, hop);
  if (!kTileSum && !a_valid) return;
  const bool b_valid = !kHalf && (fa + 1 < ti.nvalid);
  float re[32], im[32];
  if (a_valid) {
  // mel gradients (and, fused log, the forward's log output) of the warp's frames: copied global -> shared
  // asynchronously now, consumed after the forward FFT (the scattered [F][T] layout autograd hands back would
  // otherwise stall the warp for a DRAM round trip per filter)
  float* gy = gm + FPW * kGmStride;  // [FPW][kGmStride] raw log values
  if constexpr (kMel) {
    for (int m = lane; m < p.n_mels; m += 32) {
#pragma unroll
      for (int f = 0; f < FPW; ++f) {
        if (fw + f < ti.nvalid) {
          cp_async4(gm + f * kGmStride + m, gv.g + gv.g_index(fw + f, m));
          if (gv.y != nullptr) cp_async4(gy + f * kGmStride + m, gv.y + gv.y_index(fw + f, m));
        }
      }
    }
    cp_async_commit();
  }
  const float* xs = p.samples + ti.s_off;
  const long long frame_a = ti.out_frame0 + fa;

  // ---- forward recomputation: samples -> window-fused first stage -> FFT ------------------------------
  if constexpr (J > 1) {
    constexpr int HR = R1 / 2;
    constexpr int LB = (R1 == 16) ? 3 : 2;
    const float2* wv = reinterpret_cast<const float2*>(s_win) + lane;
#pragma unroll
    for (int jj = 0; jj < J; ++jj) {
      const int ua = ti.start + (fw + FPJ * jj) * hop + lane, ub = ua + hop;
#pragma unroll
      for (int r = 0; r < HR; ++r) {
        const float2 w = wv[32 * r];
        const int i = jj * R1 + 2 * bitrev_n(r, LB);
        win_head(re[i], re[i + 1], __ldg(xs + reflect_index(ua + 32 * r, ti.L)), w.x,
                 __ldg(xs + reflect_index(ua + 32 * (r + HR), ti.L)), w.y);
        win_head(im[i], im[i + 1], __ldg(xs + reflect_index(ub + 32 * r, ti.L)), w.x,
                 __ldg(xs + reflect_index(ub + 32 * (r + HR), ti.L)), w.y);
      }
    }
  } else if constexpr (!kHalf) {
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
  warp_fft1024_tail<J>(re, im, s_tw4, scr, lane);

  // ---- mel gradients of the frame(s) -> shared memory ----------------------------------------------------
  if constexpr (kMel) {
    cp_async_wait_all();
    for (int m = lane; m < p.n_mels; m += 32) {  // every lane finishes the words it copied itself
#pragma unroll
      for (int f = 0; f < FPW; ++f) {
        float v = 0.f;
        if (fw + f < ti.nvalid) {
          v = gm[f * kGmStride + m];
          if (gv.y != nullptr) v = gv.chain(v, gy[f * kGmStride + m]);
        }
        gm[f * kGmStride + m] = v;
      }
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
      const float* gf = gm + (FPJ * jw + f) * kGmStride;  // the lane's job, frame f of it
      return fmaf(wr, gf[m_r], wf * gf[m_f]);
    } else {
      return ((f == 0) ? la_valid : b_valid) ? gv.at(fa + f, k) : 0.f;
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
  const int src_lane = (lane & ~(R1 - 1)) | ((R1 - k1) & (R1 - 1));  // owner of the mirrored bin, inside the job's lanes
#pragma unroll
  for (int j = 0; j <= 16; ++j) {
    const int k = k1 + R1 * j;
    float zr, zi, pr, pi;
    if (j < 16) {
      zr = re[j];
      zi = im[j];
      const float sr = (k1 == 0) ? re[(32 - j) & 31] : re[31 - j];
      const float si = (k1 == 0) ? im[(32 - j) & 31] : im[31 - j];
      pr = __shfl_sync(0xffffffffu, sr, src_lane);
      pi = __shfl_sync(0xffffffffu, si, src_lane);
    } else {
      zr = pr = re[16];
      zi = pi = im[16];
    }
    const bool own = (j < 16) || (k1 == 0);
    if constexpr (!kHalf) {
      // window was pre-scaled by 1/2: X_a = Z[k] + conj(Z[N-k]), X_b = (Z[k] - conj(Z[N-k])) / i
      const float ar = zr + pr, ai = zi - pi;
      const float br = zi + pi, bi = pr - zr;
      const float gpa = own ? librosa(g_power(k, 0), ar, ai) : 0.f;
      const float gpb = own ? librosa(g_power(k, 1), br, bi) : 0.f;
      const float ur = gpa * ar, ui = gpa * ai, vr = gpb * br, vi = gpb * bi;  // u = G_a / 2, v = G_b / 2
      if (j == 16 || (j == 0 && k1 == 0)) {
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
  // J = 1: position bitrev5(q) of the DIT arrays holds element k1 + 32 q.  J = 2: the lane holds elements k1 + 16 q of
  // ITS job, but the transform wants lane n2 to hold rows n1 of BOTH jobs: element k1 + 16 (2 n1 + b) is row n1 of
  // column k1 + 16 b, so it is parked at position b * 16 + bitrev4(n1) (b as the block) and the values with b != job
  // change lane halves below.
  auto pos = [](int q) { return (J == 1) ? bitrev5(q) : ((q & 1) * 16 + bitrev_n(q >> 1, 4)); };
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    re[pos(q)] = ck_i[q];
    im[pos(q)] = ck_r[q];
  }
#pragma unroll
  for (int q = 16; q < 32; ++q) {
    // element k1 + R1 q > N / 2 (or == N / 2 for k1 == 0, q == 16): the mirror of element N - e, owned by src_lane
    const float sr = (k1 == 0) ? cm_r[(32 - q) & 31] : cm_r[31 - q];
    const float si = (k1 == 0) ? cm_i[(32 - q) & 31] : cm_i[31 - q];
    re[pos(q)] = __shfl_sync(0xffffffffu, si, src_lane);
    im[pos(q)] = __shfl_sync(0xffffffffu, sr, src_lane);
  }
  if constexpr (J == 2) {
    // lower lanes (job 0) keep their b = 0 values as block 0 and receive job 1's b = 0 values as block 1; upper lanes
    // (job 1) keep b = 1 as block 1 and receive job 0's b = 1 values as block 0
    const bool upper = lane >= 16;
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int p0 = bitrev_n(n1, 4), p1 = 16 + p0;
      const float gr = __shfl_xor_sync(0xffffffffu, upper ? re[p0] : re[p1], 16);
      const float gi = __shfl_xor_sync(0xffffffffu, upper ? im[p0] : im[p1], 16);
      if (upper) {
        re[p0] = gr;
        im[p0] = gi;
      } else {
        re[p1] = gr;
        im[p1] = gi;
      }
    }
  }
  dft32_dit_head(re, im);
  warp_fft1024_tail<J>(re, im, s_tw4, scr, lane);
  // FFT(swap(C)) = swap(y): real part of the inverse = im, imaginary part = re; the lane holds samples
  // n = k1 + R1 * k2 of its job's two frames
  // window of sample n: pair table entry [row % (R1 / 2)][column], .x / .y for the lower / upper half of the rows
  const float2* s_win2 = reinterpret_cast<const float2*>(s_win);
  auto win_of = [&](int k2) -> float {
    const int row = k2 / J, col = k1 + R1 * (k2 % J);
    const float2 wp = s_win2[32 * (row % (R1 / 2)) + col];
    return 2.f * ((row < R1 / 2) ? wp.x : wp.y);
  };

  // ---- times the (true) window -> frame-gradient rows --------------------------------------------------------
  if constexpr (!kTileSum) {
    float* fg_a = p.frame_grad + frame_a * NFFT;
    if constexpr (!kHalf) {
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) {
        const float w = win_of(k2);
        if (la_valid) fg_a[k1 + R1 * k2] = w * im[k2];
        if (b_valid) fg_a[NFFT + k1 + R1 * k2] = w * re[k2];
      }
    } else {
      const float2* w2 = reinterpret_cast<const float2*>(s_win) + lane;
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) {
        const float2 w = w2[32 * k2];  // {w[2m], w[2m + 1]} / 2, m = lane + 32 k2
        reinterpret_cast<float2*>(fg_a)[lane + 32 * k2] = make_float2(2.f * w.x * im[k2], 2.f * w.y * re[k2]);
      }
    }
  } else {
    // keep the windowed rows in the FFT registers: im = frame a (MODE_HALF: even samples), re = frame b (odd samples)
    if constexpr (!kHalf) {
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) {
        const float w = win_of(k2);
        im[k2] *= w;
        re[k2] *= w;
      }
    } else {
      const float2* w2 = reinterpret_cast<const float2*>(s_win) + lane;
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) {
        const float2 w = w2[32 * k2];
        im[k2] *= 2.f * w.x;
        re[k2] *= 2.f * w.y;
      }
    }
  }
  }  // a_valid

  if constexpr (kTileSum) {
    // ---- overlap-add inside the tile: frames t, t + n_round, ... do not overlap and add side by side --------------
    const int n_round = (NFFT + hop - 1) / hop;
    for (int r = 0; r < n_round; ++r) {
      if constexpr (!kHalf) {
        if (la_valid && fa % n_round == r) {
          float* acc = s_acc + fa * hop + k1;
#pragma unroll
          for (int k2 = 0; k2 < 32; ++k2) acc[R1 * k2] += im[k2];
        }
        if (n_round == 1) __syncwarp();  // frames a and b of a warp do not overlap either, but keep the order fixed
        if (b_valid && (fa + 1) % n_round == r) {
          float* acc = s_acc + (fa + 1) * hop + k1;
#pragma unroll
          for (int k2 = 0; k2 < 32; ++k2) acc[R1 * k2] += re[k2];
        }
      } else {
        if (a_valid && fa % n_round == r) {
          float2* acc = reinterpret_cast<float2*>(s_acc + fa * hop) + lane;  // hop is even in this mode
#pragma unroll
          for (int k2 = 0; k2 < 32; ++k2) {
            float2 v = acc[32 * k2];
            v.x += im[k2];
            v.y += re[k2];
            acc[32 * k2] = v;
          }
        }
      }
      __syncthreads();
    }
    // the tile's sums: scratch row of the tile's first frame onwards ((nvalid - 1) * hop + n_fft <= nvalid * n_fft)
    float* dst = p.frame_grad + ti.out_frame0 * NFFT;
    const int n4 = span_valid >> 2;
    for (int i = tid; i < n4; i += kBwdWarps * 32)
      reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_acc)[i];
    for (int i = 4 * n4 + tid; i < span_valid; i += kBwdWarps * 32) dst[i] = s_acc[i];
  }
}

// d loss / d x[s] from the tile sums: tile i of an utterance (frames [FR i, FR i + nv)) holds the sums of the padded
// positions [FR i hop, FR i hop + (nv - 1) hop + n_fft) at scratch row (first frame of the tile); a position lies in
// the tile that starts at or before it and, near the tile's head, in the tail of the one(s) before.
__global__ void __launch_bounds__(256) tile_overlap_add_kernel(const float* __restrict__ tile_sum,
                                                               const long long* __restrict__ sample_off,
                                                               const long long* __restrict__ frame_off, int n_fft,
                                                               int hop, int frames_per_tile,
                                                               float* __restrict__ grad_x) {
  const int b = blockIdx.y;
  const long long s0 = sample_off[b];
  const int L = (int)(sample_off[b + 1] - s0);   // utterances are shorter than 2^31 samples (evf_batch_create)
  const long long f0 = frame_off[b];
  const int T = (int)(frame_off[b + 1] - f0);
  const int sa = 4 * (blockIdx.x * blockDim.x + threadIdx.x);  // this thread's four consecutive samples
  if (sa >= L) return;
  const int H = n_fft / 2;
  const unsigned tile_hop = (unsigned)frames_per_tile * (unsigned)hop;
  const int n_tiles = (T + frames_per_tile - 1) / frames_per_tile;
  // sum of the tile sums that cover padded position pp, from the last tile that starts at or before it downwards
  auto at = [&](int pp) -> float {
    if (n_tiles == 0) return 0.f;
    int ti = (int)((unsigned)pp / tile_hop);
    if (ti > n_tiles - 1) ti = n_tiles - 1;
    long long o = (long long)pp - (long long)ti * tile_hop;
    float g = 0.f;
    for (; ti >= 0; --ti, o += tile_hop) {
      const int nv = min(frames_per_tile, T - ti * frames_per_tile);
      if (o >= (long long)(nv - 1) * hop + n_fft) break;
      g += __ldg(tile_sum + (f0 + (long long)ti * frames_per_tile) * n_fft + o);
    }
    return g;
  };
  float g[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int s = sa + j;
    g[j] = 0.f;
    if (s < L) {
      g[j] = at(s + H);
      // the mirror images of x[s] in the left / right reflect margin
      if (s >= 1 && s <= H) g[j] += at(H - s);
      if (s <= L - 2 && s >= L - 1 - H) g[j] += at(H + 2 * (L - 1) - s);
    }
  }
  float* dst = grad_x + s0 + sa;
  if (sa + 3 < L && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
    *reinterpret_cast<float4*>(dst) = make_float4(g[0], g[1], g[2], g[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (sa + j < L) dst[j] = g[j];
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

template <int MODE, int SPEC, bool kTileSum>
int launch_bwd_t(const BwdParams& p, int smem, cudaStream_t st) {
  auto k = features_backward_kernel<MODE, SPEC, kTileSum>;
  EVF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k<<<kBwdSplit * p.n_tiles, kBwdWarps * 32, smem, st>>>(p);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}
template <int SPEC>
int launch_bwd(const BwdParams& p, int smem, bool tile_sum, cudaStream_t st) {
  if (p.n_fft == 512)
    return tile_sum ? launch_bwd_t<MODE_PACK2_512, SPEC, true>(p, smem, st)
                    : launch_bwd_t<MODE_PACK2_512, SPEC, false>(p, smem, st);
  if (p.n_fft == 2048)
    return tile_sum ? launch_bwd_t<MODE_HALF, SPEC, true>(p, smem, st) : launch_bwd_t<MODE_HALF, SPEC, false>(p, smem, st);
  return tile_sum ? launch_bwd_t<MODE_PACK2, SPEC, true>(p, smem, st) : launch_bwd_t<MODE_PACK2, SPEC, false>(p, smem, st);
}

}  // namespace

int features_backward_launch(const BwdParams& p, cudaStream_t st) {
  if (p.n_tiles == 0 || p.n_utts == 0) return EVF_OK;
  if (p.n_mels > kGmStride) {
    set_error("evf_features_backward: more than 128 mel filters are not supported");
    return EVF_ERR_UNSUPPORTED;
  }
  const int fpw = (p.n_fft == 2048) ? 1 : (p.n_fft == 512 ? 4 : 2);  // frames of a warp
  int smem = (p.n_fft + 2 * kFftSize + (p.n_fft == 2048 ? 1028 : 0) +
              kBwdWarps * (32 * kScrStride + 2 * fpw * kGmStride)) * (int)sizeof(float);
  // overlap-add inside the tile when its span fits beside the FFT scratch and the rounds stay few
  const int frames_per_tile = fpw * kBwdWarps;  // of a CTA: 1 / kBwdSplit of a forward tile
  const long long span_bytes = 4ll * (((long long)(frames_per_tile - 1) * p.hop + p.n_fft + 3) & ~3ll);
  const bool tile_sum = (p.n_fft + p.hop - 1) / p.hop <= 16 && smem + span_bytes <= 227 * 1024;
  if (tile_sum) smem += (int)span_bytes;
  int rc;
  switch (p.spec_type) {
    case EVF_SPEC_MEL: rc = launch_bwd<EVF_SPEC_MEL>(p, smem, tile_sum, st); break;
    case EVF_SPEC_MEL_LIBROSA: rc = launch_bwd<EVF_SPEC_MEL_LIBROSA>(p, smem, tile_sum, st); break;
    case EVF_SPEC_LINEAR: rc = launch_bwd<EVF_SPEC_LINEAR>(p, smem, tile_sum, st); break;
    default:
      set_error("evf_features_backward: spec_type has no backward (raw is complex)");
      return EVF_ERR_UNSUPPORTED;
  }
  if (rc != EVF_OK) return rc;
  if (!tile_sum)
    return overlap_add_launch(p.frame_grad, p.sample_off, p.frame_off, p.n_utts, p.max_len, p.n_fft, p.hop,
                              p.grad_samples, st);
  if (p.max_len <= 0) return EVF_OK;
  for (int y0 = 0; y0 < p.n_utts; y0 += 65535) {  // gridDim.y limit: slices of utterances
    const dim3 grid((unsigned)((p.max_len + 1023) / 1024), (unsigned)(p.n_utts - y0 < 65535 ? p.n_utts - y0 : 65535));
    tile_overlap_add_kernel<<<grid, 256, 0, st>>>(p.frame_grad, p.sample_off + y0, p.frame_off + y0, p.n_fft, p.hop,
                                                  frames_per_tile, p.grad_samples);
    EVF_CUDA(cudaGetLastError());
  }
  return EVF_OK;
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
