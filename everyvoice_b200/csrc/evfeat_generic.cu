// Feature kernel for ANY transform size (sm_100a): framing -> window -> FFT -> |X|^2 -> mel -> log -> energy for
// every n_fft / win_length / hop_length torch.stft accepts, i.e. everything get_spectral_transform can be asked for:
//   everyvoice/utils/heavy.py:47-119            (n_fft, win, hop are unconstrained config fields,
//   everyvoice/config/preprocessing_config.py:56-70)
//   everyvoice/preprocessor/preprocessor.py:94-121   the "output" transform multiplies n_fft / win / hop by
//                                                    output_sr // input_sr (3072 / 768 for 16 -> 48 kHz, ...)
// The warp-per-FFT kernel of evfeat_features.cu covers n_fft 1024 / 2048 (the benchmarked configurations); this
// kernel covers the rest of the domain with the same epilogue arithmetic.
//
// Work decomposition
//   tile   : PAIRS frame pairs of one utterance; frames 2j, 2j + 1 ride as real / imaginary part of ONE complex FFT
//            of n_fft points (the pair trick works for every n_fft, odd ones included)
//   CTA    : 256 threads = PAIRS teams; a team owns one pair; grid-stride loop over the host-built tile list
//   FFT    : mixed-radix Stockham autosort in shared memory (two ping-pong buffers per pair).  Radix 4 / 2 / 3 / 5
//            butterflies live in registers with float32 twiddles from one fp64-built table W_N^m; any other prime
//            factor p runs as a direct p-point DFT per output with float64 twiddles and accumulation (so a prime
//            n_fft degenerates to an exact O(N^2) DFT instead of an error)
//   epilogue: pair separation X_a = Z[k] + conj(Z[N-k]), X_b = (Z[k] - conj(Z[N-k])) / i (window pre-scaled by 1/2),
//            |X|^2 (sqrt(. + 1e-9) for mel-librosa) -> shared memory -> one thread per (frame, filter) walks the
//            filter's rising and falling bins (every bin feeds two adjacent triangular filters) or, for a bank that
//            is not triangular, the dense column -> log(max(., clip)) -> coalesced stores; the per-frame energy
//            sqrt(sum log^2) is reduced in a fixed order (deterministic)
#include "evfeat_fft.cuh"
#include "evfeat_internal.h"

namespace evf {

namespace {

#include "evfeat_device.cuh"

constexpr int kGenThreads = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }  // -i * a

// A team (the threads of one frame pair: a multiple of 32) synchronises on its own named barrier; the teams of a CTA
// never wait for each other.
__device__ __forceinline__ void team_sync(int team, int ts) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(ts) : "memory");
}

template <int P>
__device__ __forceinline__ void dft_small(float2 (&v)[P]);
template <>
__device__ __forceinline__ void dft_small<2>(float2 (&v)[2]) {
  const float2 a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void dft_small<4>(float2 (&v)[4]) {
  const float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
  const float2 t2 = cadd(v[1], v[3]), t3 = mul_neg_i(csub(v[1], v[3]));
  v[0] = cadd(t0, t2);
  v[1] = cadd(t1, t3);
  v[2] = csub(t0, t2);
  v[3] = csub(t1, t3);
}
template <>
__device__ __forceinline__ void dft_small<3>(float2 (&v)[3]) {
  constexpr float c = 0.86602540378443864676f;  // sin(2 pi / 3)
  const float2 s = cadd(v[1], v[2]), d = csub(v[1], v[2]);
  const float2 m = make_float2(fmaf(-0.5f, s.x, v[0].x), fmaf(-0.5f, s.y, v[0].y));
  const float2 n = make_float2(c * d.y, -c * d.x);  // -i * c * d
  v[0] = cadd(v[0], s);
  v[1] = cadd(m, n);
  v[2] = csub(m, n);
}
template <>
__device__ __forceinline__ void dft_small<5>(float2 (&v)[5]) {
  constexpr float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;  // cos(2 pi / 5), cos(4 pi / 5)
  constexpr float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;   // sin(2 pi / 5), sin(4 pi / 5)
  const float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
  const float2 m1 = make_float2(fmaf(c1, a1.x, fmaf(c2, a2.x, v[0].x)), fmaf(c1, a1.y, fmaf(c2, a2.y, v[0].y)));
  const float2 m2 = make_float2(fmaf(c2, a1.x, fmaf(c1, a2.x, v[0].x)), fmaf(c2, a1.y, fmaf(c1, a2.y, v[0].y)));
  const float2 n1 = make_float2(fmaf(s1, b1.x, s2 * b2.x), fmaf(s1, b1.y, s2 * b2.y));
  const float2 n2 = make_float2(fmaf(s2, b1.x, -s1 * b2.x), fmaf(s2, b1.y, -s1 * b2.y));
  const float2 in1 = mul_neg_i(n1), in2 = mul_neg_i(n2);
  v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
  v[1] = cadd(m1, in1);
  v[2] = cadd(m2, in2);
  v[3] = csub(m2, in2);
  v[4] = csub(m1, in1);
}

// One Stockham stage of radix P over the N points of one pair: butterfly j reads in[j + r * N/P], applies
// W_{Ns P}^{k r} (k = j mod Ns; Ns = product of the radices already done), a P-point DFT, and writes
// out[(j - k) * P + k + r * Ns].  After the last stage the output is in natural order.
template <int P>
__device__ __forceinline__ void stage_radix(const float2* __restrict__ in, float2* __restrict__ out,
                                            const float2* __restrict__ tw, int N, int Ns, int lt, int ts) {
  const int nb = N / P;
  const int step = N / (Ns * P);
  const bool pow2 = (Ns & (Ns - 1)) == 0;  // the usual case: no integer division in the loop
  for (int j = lt; j < nb; j += ts) {
    const int k = pow2 ? (j & (Ns - 1)) : (j % Ns);
    float2 v[P];
#pragma unroll
    for (int r = 0; r < P; ++r) v[r] = in[j + r * nb];
    if (Ns > 1) {
      const int ks = k * step;  // k * r * step < N for r < P
#pragma unroll
      for (int r = 1; r < P; ++r) v[r] = cmul(v[r], __ldg(tw + ks * r));
    }
    dft_small<P>(v);
    const int j0 = (j - k) * P + k;
#pragma unroll
    for (int r = 0; r < P; ++r) out[j0 + r * Ns] = v[r];
  }
}

// Any other prime factor p: one thread per output (j, q), a direct p-point DFT with the stage twiddle folded in
// (combined exponent r * c mod N), float64 twiddles and accumulation.
__device__ __forceinline__ void stage_generic(const float2* __restrict__ in, float2* __restrict__ out,
                                              const double2* __restrict__ tw64, int N, int Ns, int p, int lt, int ts) {
  const int nb = N / p;
  const int step = N / (Ns * p);
  for (int o = lt; o < N; o += ts) {
    const int k = o % Ns;
    const int t = o / Ns;
    const int q = t % p;
    const int j = (t / p) * Ns + k;
    const int c = (int)(((long long)k * step + (long long)q * nb) % N);
    double ar = 0.0, ai = 0.0;
    int idx = 0;
    for (int r = 0; r < p; ++r) {
      const float2 a = in[j + r * nb];
      const double2 w = __ldg(tw64 + idx);
      ar = fma((double)a.x, w.x, fma(-(double)a.y, w.y, ar));
      ai = fma((double)a.x, w.y, fma((double)a.y, w.x, ai));
      idx += c;
      if (idx >= N) idx -= N;
    }
    out[o] = make_float2((float)ar, (float)ai);
  }
}

template <int SPEC, typename SampleT>
__global__ void __launch_bounds__(kGenThreads) features_generic_kernel(const GenParams p) {
  constexpr bool kMel = (SPEC == EVF_SPEC_MEL || SPEC == EVF_SPEC_MEL_LIBROSA);
  extern __shared__ __align__(16) float2 gsm[];
  const int N = p.n_fft;
  const int pairs = p.pairs;
  const int ts = kGenThreads / pairs;      // threads per team
  const int tid = threadIdx.x;
  const int team = tid / ts, lt = tid % ts;
  float2* bufA = gsm + (size_t)team * 2 * N;
  float2* bufB = bufA + N;
  float* s_red = reinterpret_cast<float*>(gsm + (size_t)pairs * 2 * N);  // [pairs][2][8] per-warp energy partials
  const SampleT* __restrict__ samples = static_cast<const SampleT*>(p.samples);
  const int hop = p.hop;
  const int n_freq = p.n_freq;

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const TileDesc ti = load_tile_desc(p.tiles, tile);
    const int fa = 2 * team;                       // tile-local frame a of this team's pair
    const bool a_valid = fa < ti.nvalid, b_valid = fa + 1 < ti.nvalid;
    const SampleT* src = samples + ti.s_off;
    // ---- framing + window: z[n] = (w[n] x_a[n], w[n] x_b[n]) ---------------------------------------------
    if (a_valid) {
      const int s0 = ti.start + fa * hop;
      for (int n = lt; n < N; n += ts) {
        const float w = __ldg(p.window + n);
        const float xa = to_float(__ldg(src + reflect_index(s0 + n, ti.L)));
        const float xb = b_valid ? to_float(__ldg(src + reflect_index(s0 + hop + n, ti.L))) : 0.f;
        bufA[n] = make_float2(w * xa, w * xb);
      }
    }
    team_sync(team, ts);
    // ---- FFT ----------------------------------------------------------------------------------------------
    float2* cur = bufA;
    float2* oth = bufB;
    int Ns = 1;
    for (int s = 0; s < p.st.n; ++s) {
      const int radix = p.st.radix[s];
      if (a_valid) {
        switch (radix) {
          case 4: stage_radix<4>(cur, oth, p.tw, N, Ns, lt, ts); break;
          case 2: stage_radix<2>(cur, oth, p.tw, N, Ns, lt, ts); break;
          case 3: stage_radix<3>(cur, oth, p.tw, N, Ns, lt, ts); break;
          case 5: stage_radix<5>(cur, oth, p.tw, N, Ns, lt, ts); break;
          default: stage_generic(cur, oth, p.tw64, N, Ns, radix, lt, ts); break;
        }
      }
      team_sync(team, ts);
      float2* t = cur;
      cur = oth;
      oth = t;
      Ns *= radix;
    }
    // ---- pair separation, power, output ---------------------------------------------------------------------
    const long long fr_a = ti.out_frame0 + fa;
    float* ga = p.spec_out + fr_a * (long long)p.row_floats;
    float* gb = ga + p.row_floats;
    float ea = 0.f, eb = 0.f;
    float* Pa = reinterpret_cast<float*>(oth);  // the buffer that does not hold Z: 2 * N floats >= 2 * n_freq
    float* Pb = Pa + n_freq;
    if (a_valid) {
      for (int k = lt; k < n_freq; k += ts) {
        const float2 z = cur[k];
        const float2 m = cur[k == 0 ? 0 : N - k];
        const float ar = z.x + m.x, ai = z.y - m.y;   // X_a = Z[k] + conj(Z[N - k])   (window carries the 1/2)
        const float br = z.y + m.y, bi = m.x - z.x;   // X_b = (Z[k] - conj(Z[N - k])) / i
        if constexpr (SPEC == EVF_SPEC_RAW) {
          reinterpret_cast<float2*>(ga)[k] = make_float2(ar, ai);
          if (b_valid) reinterpret_cast<float2*>(gb)[k] = make_float2(br, bi);
        } else {
          float pa = fmaf(ar, ar, ai * ai), pb = fmaf(br, br, bi * bi);
          if constexpr (SPEC == EVF_SPEC_MEL_LIBROSA) {
            pa = sqrtf(pa + 1e-9f);
            pb = sqrtf(pb + 1e-9f);
          }
          if constexpr (kMel) {
            Pa[k] = pa;
            Pb[k] = pb;
          } else {
            const float va = compress(pa, p.apply_log, p.log_clip);
            ga[k] = va;
            ea = fmaf(va, va, ea);
            if (b_valid) {
              const float vb = compress(pb, p.apply_log, p.log_clip);
              gb[k] = vb;
              eb = fmaf(vb, vb, eb);
            }
          }
        }
      }
    }
    if constexpr (kMel) {
      team_sync(team, ts);
      if (a_valid) {
        const int n_mels = p.n_mels;
        for (int o = lt; o < 2 * n_mels; o += ts) {
          const int f = o / n_mels, m = o - f * n_mels;  // f: 0 = frame a, 1 = frame b
          if (f == 1 && !b_valid) break;
          const float* P = f ? Pb : Pa;
          float acc = 0.f;
          if (p.fb_dense != nullptr) {
            for (int k = 0; k < n_freq; ++k) acc = fmaf(__ldg(p.fb_dense + (size_t)k * n_mels + m), P[k], acc);
          } else {
            const int k0 = __ldg(p.kstart + m), k1 = __ldg(p.kstart + m + 1), k2 = __ldg(p.kstart + m + 2);
            for (int k = k0; k < k1; ++k) acc = fmaf(__ldg(&p.melw[k].x), P[k], acc);  // rising side: interval m
            for (int k = k1; k < k2; ++k) acc = fmaf(__ldg(&p.melw[k].y), P[k], acc);  // falling side: interval m + 1
          }
          const float v = compress(acc, p.apply_log, p.log_clip);
          (f ? gb : ga)[m] = v;
          if (f) eb = fmaf(v, v, eb); else ea = fmaf(v, v, ea);
        }
      }
    }
    // ---- energy: fixed-order reduction (lanes by shuffle, warps of the team through shared memory) --------
    if constexpr (SPEC != EVF_SPEC_RAW) {
      if (p.energy_out != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ea += __shfl_xor_sync(0xffffffffu, ea, o);
          eb += __shfl_xor_sync(0xffffffffu, eb, o);
        }
        const int warp = tid >> 5, lane = tid & 31;
        if (lane == 0) {
          s_red[2 * warp] = ea;
          s_red[2 * warp + 1] = eb;
        }
        team_sync(team, ts);
        if (lt == 0 && a_valid) {
          const int w0 = (team * ts) >> 5, nw = ts >> 5;  // the team's warps (ts is a multiple of 32)
          float sa = 0.f, sb = 0.f;
          for (int w = 0; w < nw; ++w) {
            sa += s_red[2 * (w0 + w)];
            sb += s_red[2 * (w0 + w) + 1];
          }
          p.energy_out[fr_a] = sqrtf(sa);
          if (b_valid) p.energy_out[fr_a + 1] = sqrtf(sb);
        }
      }
    }
    team_sync(team, ts);  // the team's buffers are reused by its next tile
  }
}

// ---- backward of the any-size transform ------------------------------------------------
// d loss / d (windowed frame) for every frame pair of a tile, written as one row of n_fft floats per frame into the
// scratch that overlap_add_kernel (evfeat_backward.cu) folds back onto the samples.  Same steps as the warp kernel's
// backward: recompute Z of the pair, separate X_a / X_b, transposed mel projection (a bin feeds two adjacent filters:
// no reduction; a bank that is not triangular takes the dense row), G = 2 g_P X (x 1 / (2 sqrt(P + 1e-9)) for
// mel-librosa), Hermitian extension packed as C = C_a + i C_b, inverse DFT through the FORWARD Stockham stages with
// real and imaginary parts swapped on the way in and out, times the window.
template <int SPEC>
__global__ void __launch_bounds__(kGenThreads) features_generic_backward_kernel(const GenParams p, const GradSrc gsrc,
                                                                                float* __restrict__ frame_grad,
                                                                                const int* __restrict__ jk, int k_used) {
  constexpr bool kMel = (SPEC == EVF_SPEC_MEL || SPEC == EVF_SPEC_MEL_LIBROSA);
  extern __shared__ __align__(16) float2 gsm[];
  const int N = p.n_fft;
  const int pairs = p.pairs;
  const int ts = kGenThreads / pairs;
  const int tid = threadIdx.x;
  const int team = tid / ts, lt = tid % ts;
  float2* bufA = gsm + (size_t)team * 2 * N;
  float2* bufB = bufA + N;
  float* s_gm = reinterpret_cast<float*>(gsm + (size_t)pairs * 2 * N) + (size_t)team * 2 * p.n_mels;  // [2][n_mels]
  const float* __restrict__ samples = static_cast<const float*>(p.samples);
  const int hop = p.hop, n_freq = p.n_freq, n_mels = p.n_mels;

  auto fft = [&](float2*& cur, float2*& oth) {  // forward transform of `cur`, in natural order, result in `cur`
    int Ns = 1;
    for (int s = 0; s < p.st.n; ++s) {
      const int radix = p.st.radix[s];
      switch (radix) {
        case 4: stage_radix<4>(cur, oth, p.tw, N, Ns, lt, ts); break;
        case 2: stage_radix<2>(cur, oth, p.tw, N, Ns, lt, ts); break;
        case 3: stage_radix<3>(cur, oth, p.tw, N, Ns, lt, ts); break;
        case 5: stage_radix<5>(cur, oth, p.tw, N, Ns, lt, ts); break;
        default: stage_generic(cur, oth, p.tw64, N, Ns, radix, lt, ts); break;
      }
      team_sync(team, ts);
      float2* t = cur;
      cur = oth;
      oth = t;
      Ns *= radix;
    }
  };

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const TileDesc ti = load_tile_desc(p.tiles, tile);
    const int fa = 2 * team;
    const bool a_valid = fa < ti.nvalid, b_valid = fa + 1 < ti.nvalid;
    const float* src = samples + ti.s_off;
    const long long fr_a = ti.out_frame0 + fa;
    const GradView gv(gsrc, ti, N, hop);
    if (a_valid) {
      const int s0 = ti.start + fa * hop;
      for (int n = lt; n < N; n += ts) {
        const float w = __ldg(p.window + n);
        const float xa = __ldg(src + reflect_index(s0 + n, ti.L));
        const float xb = b_valid ? __ldg(src + reflect_index(s0 + hop + n, ti.L)) : 0.f;
        bufA[n] = make_float2(w * xa, w * xb);
      }
      if constexpr (kMel) {
        for (int m = lt; m < n_mels; m += ts) {
          s_gm[m] = gv.at(fa, m);
          s_gm[n_mels + m] = b_valid ? gv.at(fa + 1, m) : 0.f;
        }
      }
    }
    team_sync(team, ts);
    float2* cur = bufA;
    float2* oth = bufB;
    fft(cur, oth);  // every thread takes part in the barriers; idle teams transform stale data that is never used
    // ---- spectrum gradients -> packed, Hermitian-extended input of the inverse (re / im swapped) in `oth` --------
    if (a_valid) {
      auto g_power = [&](int k, int f) -> float {
        if constexpr (kMel) {
          const float* gm = s_gm + f * n_mels;
          if (p.fb_dense != nullptr) {
            float acc = 0.f;
            for (int m = 0; m < n_mels; ++m) acc = fmaf(__ldg(p.fb_dense + (size_t)k * n_mels + m), gm[m], acc);
            return acc;
          }
          if (k >= k_used) return 0.f;
          const float2 w = __ldg(p.melw + k);
          const int jj = __ldg(jk + k);
          const float wr = (jj < n_mels) ? w.x : 0.f, wf = (jj >= 1) ? w.y : 0.f;
          return fmaf(wr, gm[min(jj, n_mels - 1)], wf * gm[max(jj - 1, 0)]);
        } else {
          return (f == 0 || b_valid) ? gv.at(fa + f, k) : 0.f;
        }
      };
      for (int k = lt; k < n_freq; k += ts) {
        const int km = (k == 0) ? 0 : N - k;
        const float2 z = cur[k], m = cur[km];
        const float ar = z.x + m.x, ai = z.y - m.y;
        const float br = z.y + m.y, bi = m.x - z.x;
        float gpa = g_power(k, 0), gpb = g_power(k, 1);
        if constexpr (SPEC == EVF_SPEC_MEL_LIBROSA) {
          gpa *= 0.5f * rsqrtf(fmaf(ar, ar, ai * ai) + 1e-9f);
          gpb *= 0.5f * rsqrtf(fmaf(br, br, bi * bi) + 1e-9f);
        }
        const float ur = gpa * ar, ui = gpa * ai, vr = gpb * br, vi = gpb * bi;  // u = G_a / 2, v = G_b / 2
        if (km == k) {  // k = 0, and k = N / 2 for even N: the spectrum is real there; C = Re G_a + i Re G_b
          oth[k] = make_float2(2.f * vr, 2.f * ur);              // swapped: (im, re)
        } else {
          oth[k] = make_float2(ui + vr, ur - vi);                // u + i v, swapped
          oth[km] = make_float2(vr - ui, ur + vi);               // conj(u) + i conj(v), swapped
        }
      }
    }
    team_sync(team, ts);
    float2* c2 = oth;
    float2* o2 = cur;
    fft(c2, o2);
    // FFT(swap(C)) = swap(r_a + i r_b): r_a = imaginary part, r_b = real part; times the true window (2 x stored)
    if (a_valid) {
      float* fg = frame_grad + fr_a * (long long)N;
      for (int n = lt; n < N; n += ts) {
        const float w = 2.f * __ldg(p.window + n);
        const float2 y = c2[n];
        fg[n] = w * y.y;
        if (b_valid) fg[N + n] = w * y.x;
      }
    }
    team_sync(team, ts);
  }
}

template <int SPEC>
int launch_generic_s(int fmt, const GenParams& p, int grid, int smem, cudaStream_t st, bool cfg) {
  if (fmt == EVF_SAMPLES_S16) {
    auto k = features_generic_kernel<SPEC, short>;
    if (cfg) {
      EVF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      return EVF_OK;
    }
    k<<<grid, kGenThreads, smem, st>>>(p);
  } else {
    auto k = features_generic_kernel<SPEC, float>;
    if (cfg) {
      EVF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      return EVF_OK;
    }
    k<<<grid, kGenThreads, smem, st>>>(p);
  }
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

int dispatch_generic(int spec, int fmt, const GenParams& p, int grid, int smem, cudaStream_t st, bool cfg) {
  switch (spec) {
    case EVF_SPEC_MEL: return launch_generic_s<EVF_SPEC_MEL>(fmt, p, grid, smem, st, cfg);
    case EVF_SPEC_MEL_LIBROSA: return launch_generic_s<EVF_SPEC_MEL_LIBROSA>(fmt, p, grid, smem, st, cfg);
    case EVF_SPEC_LINEAR: return launch_generic_s<EVF_SPEC_LINEAR>(fmt, p, grid, smem, st, cfg);
    case EVF_SPEC_RAW: return launch_generic_s<EVF_SPEC_RAW>(fmt, p, grid, smem, st, cfg);
  }
  set_error("unknown spec_type");
  return EVF_ERR_UNSUPPORTED;
}

}  // namespace

// Radix sequence of the Stockham FFT: 4s, then 2, 3s, 5s, then whatever primes remain (direct DFT stages).
int generic_factorize(int n_fft, GenStages* st, bool* needs_tw64) {
  st->n = 0;
  *needs_tw64 = false;
  int n = n_fft;
  auto push = [&](int r) {
    if (st->n >= kGenMaxStages) return false;
    st->radix[st->n++] = r;
    return true;
  };
  while (n % 4 == 0) { if (!push(4)) return EVF_ERR_UNSUPPORTED; n /= 4; }
  while (n % 2 == 0) { if (!push(2)) return EVF_ERR_UNSUPPORTED; n /= 2; }
  while (n % 3 == 0) { if (!push(3)) return EVF_ERR_UNSUPPORTED; n /= 3; }
  while (n % 5 == 0) { if (!push(5)) return EVF_ERR_UNSUPPORTED; n /= 5; }
  for (int q = 7; n > 1; q += 2) {
    if ((long long)q * q > n) q = n;  // what is left is prime
    while (n % q == 0) {
      if (!push(q)) return EVF_ERR_UNSUPPORTED;
      *needs_tw64 = true;
      n /= q;
    }
  }
  return EVF_OK;
}

// Frame pairs per CTA iteration and dynamic shared memory; -1 if n_fft does not fit.
int generic_smem_bytes(int n_fft, int* pairs_out) {
  int pairs = 1;
  while (pairs < 8 && (long long)(2 * pairs) * n_fft <= 2048) pairs *= 2;  // teams of >= 32 threads with >= 4 points per thread
  const long long bytes = (long long)pairs * 2 * n_fft * (long long)sizeof(float2) + 2 * (kGenThreads / 32) * sizeof(float);
  *pairs_out = pairs;
  return bytes <= 227 * 1024 ? (int)bytes : -1;
}

int generic_configure(int spec_type, int sample_format, int smem_bytes) {
  GenParams dummy{};
  return dispatch_generic(spec_type, sample_format, dummy, 1, smem_bytes, nullptr, true);
}

int generic_launch(int spec_type, int sample_format, const GenParams& p, int grid, int smem_bytes, cudaStream_t stream) {
  return dispatch_generic(spec_type, sample_format, p, grid, smem_bytes, stream, false);
}

int generic_backward_launch(int spec_type, const GenParams& p, const GradSrc& grad_spec, float* frame_grad, const int* jk,
                            int k_used, int grid, int smem_bytes, cudaStream_t st) {
  // + [pairs][2][n_mels] mel gradients behind the FFT buffers
  const int smem = smem_bytes + p.pairs * 2 * (p.n_mels > 0 ? p.n_mels : 0) * (int)sizeof(float);
  if (smem > 227 * 1024) {
    set_error("evf_features_backward: n_fft / n_mels too large for the backward's shared memory");
    return EVF_ERR_UNSUPPORTED;
  }
#define EVF_LAUNCH_GBWD(SPEC)                                                                              \
  {                                                                                                        \
    auto k = features_generic_backward_kernel<SPEC>;                                                       \
    EVF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                  \
    k<<<grid, kGenThreads, smem, st>>>(p, grad_spec, frame_grad, jk, k_used);                              \
  }
  switch (spec_type) {
    case EVF_SPEC_MEL: EVF_LAUNCH_GBWD(EVF_SPEC_MEL) break;
    case EVF_SPEC_MEL_LIBROSA: EVF_LAUNCH_GBWD(EVF_SPEC_MEL_LIBROSA) break;
    case EVF_SPEC_LINEAR: EVF_LAUNCH_GBWD(EVF_SPEC_LINEAR) break;
    default:
      set_error("evf_features_backward: spec_type has no backward (raw is complex)");
      return EVF_ERR_UNSUPPORTED;
  }
#undef EVF_LAUNCH_GBWD
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

}  // namespace evf
