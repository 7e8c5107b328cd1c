// Fused framing -> window -> FFT -> |X|^2 -> mel -> log -> energy kernel (sm_100a).
//
// Replaces, for a ragged batch of utterances in one launch, what the reference does per
// utterance through torchaudio on the CPU:
//   everyvoice/utils/heavy.py:47-113            get_spectral_transform (mel / mel-librosa / linear / raw)
//   everyvoice/utils/heavy.py:39-40             dynamic_range_compression_torch
//   everyvoice/preprocessor/preprocessor.py:220-233, 921-927   extract_spectral_features + [:, :L//hop]
//   everyvoice/preprocessor/preprocessor.py:302-309            extract_energy
//
// Work decomposition
//   grid   : persistent, one CTA of 16 warps per SM, static round-robin over frame tiles
//   tile   : 16 FFT jobs of one utterance = 32 consecutive frames (n_fft 1024: two real
//            frames ride as re/im of one complex FFT) or 16 frames (n_fft 2048: one frame
//            = 1024 complex points + real-FFT split)
//   phase 0: the tile's sample span ((FR-1)*hop + n_fft samples; every sample is read from
//            HBM once, the 4x frame overlap is served from shared memory) is staged with
//            reflect padding resolved by index mirroring at the utterance edges
//   phase A: one warp per FFT job: window * samples -> registers, 32x32 four-step FFT
//            (in-register radix-2 DFT-32, transpose + twiddle through shared memory),
//            real-FFT separation with the mirrored bin fetched by warp shuffle, |X|^2
//            (or sqrt(|X|^2+1e-9) for mel-librosa) -> shared P[bin][frame];
//            linear / raw write straight to HBM (lanes = consecutive bins, coalesced)
//   phase B: mel projection with lane = frame: all lanes walk the same bins, weights are
//            warp-uniform; each bin lies between two adjacent filter centres so it feeds
//            exactly two filters: two running FMAs per bin, no atomics, no divergence
//   phase C: log-mel tile -> HBM (coalesced rows) + per-frame energy = sqrt(sum log^2)
#include "evfeat_fft.cuh"
#include "evfeat_internal.h"

namespace evf {

namespace {

#include "evfeat_device.cuh"

using TileInfo = TileDesc;  // host-built, 32 bytes, one per tile (evfeat_internal.h)

// The manual (non-bulk) part of a tile: tile words [0, a_lo) and [a_hi, span).
struct ManualRange {
  int a_lo, a_hi, total;
};

template <int MODE, int SPEC, typename SampleT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 16 / WARPS) features_kernel(const FeatParams p) {
  using MT = ModeTraits<MODE>;
  constexpr int kThreads = WARPS * 32;
  constexpr int NFFT = MT::kNfft;
  constexpr int FPJ = MT::kFramesPerJob;
  constexpr int FR = WARPS * FPJ;         // frames per tile
  constexpr int FS = FR + 1;              // padded frame stride of the P tile
  constexpr int PARTS = 32 / FR;          // projection workers per warp
  constexpr bool kMel = (SPEC == EVF_SPEC_MEL || SPEC == EVF_SPEC_MEL_LIBROSA);
  constexpr bool kBulk = sizeof(SampleT) == 4;  // float samples can be bulk-copied as they are
  constexpr bool kHalf = (MODE != MODE_PACK2);          // n_fft 2048: one frame per FFT + real-FFT split
  constexpr bool kStageWpost = (MODE == MODE_HALF);     // MODE_HALF_L1: post-twiddles read through L1

  extern __shared__ __align__(16) float smem[];
  float* s_win = smem + p.off_win;
  float4* s_tw4 = reinterpret_cast<float4*>(smem + p.off_tw);
  const float2* s_wpost = reinterpret_cast<const float2*>(smem + p.off_wpost);
  float4* s_mw4 = reinterpret_cast<float4*>(smem + p.off_melw);
  int* s_vwk = reinterpret_cast<int*>(smem + p.off_vwk);
  float* s_p = smem + p.off_p;
  float* s_sa = smem + p.off_sa;
  float* s_sb = smem + p.off_sb;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  float* scr = smem + p.off_scr + warp * (32 * kScrStride);
  const int nbuf = p.nbuf;

  // ---- one-time setup: mbarriers + table copy --------------------------------------------
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < NFFT; i += kThreads) s_win[i] = p.window[i];
  for (int i = tid; i < kFftSize / 2; i += kThreads) s_tw4[i] = p.tw4[i];
  if constexpr (kStageWpost) {
    for (int i = tid; i <= 512; i += kThreads) reinterpret_cast<float2*>(smem + p.off_wpost)[i] = p.wpost[i];
  }
  if constexpr (kMel) {
    for (int i = tid; i < p.k_used; i += kThreads) s_mw4[i] = p.melw4[i];
    for (int i = tid; i < WARPS * PARTS + 1; i += kThreads) s_vwk[i] = p.vw_k[i];
    for (int i = tid; i < 2 * (p.n_mels + 1) * FS; i += kThreads) s_sa[i] = 0.f;  // SA and SB are adjacent
  }
  __syncthreads();

  const int hop = p.hop;
  const SampleT* __restrict__ samples = static_cast<const SampleT*>(p.samples);

  // Tile descriptors are host-built (no dependent loads) and fetched one iteration ahead.
  auto tile_info = [&](int tile) {
    const int4* q = reinterpret_cast<const int4*>(p.tiles + tile);
    const int4 a = __ldg(q), b = __ldg(q + 1);
    TileInfo ti;
    ti.s_off = ((long long)(unsigned)a.x) | ((long long)a.y << 32);
    ti.out_frame0 = ((long long)(unsigned)a.z) | ((long long)a.w << 32);
    ti.L = b.x;
    ti.start = b.y;
    ti.nvalid = b.z;
    ti.span = b.w;
    return ti;
  };

  // Issue the staging of a tile into `buf`: one bulk copy for the 16-byte aligned in-range part
  // (thread 0), the rest (reflected margins, unaligned / int16 input) is left to the caller.
  auto stage_issue = [&](const TileInfo& ti, float* buf, uint64_t* bar) {
    ManualRange mr;
    const int lo = max(0, -ti.start);                 // first tile word inside the utterance
    const int hi = min(ti.span, ti.L - ti.start);     // one past the last
    mr.a_lo = lo;
    mr.a_hi = lo;
    if constexpr (kBulk) {
      // tile word i <-> packed sample s_off + start + i ; both sides must be 16-byte aligned
      if ((((ti.s_off + ti.start + lo) | lo) & 3) == 0 && hi > lo) mr.a_hi = lo + ((hi - lo) & ~3);
    }
    mr.total = mr.a_lo + (ti.span - mr.a_hi);
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)(mr.a_hi - mr.a_lo) * 4u;
      mbar_arrive_expect_tx(bar, bytes);
      if (bytes) bulk_g2s(buf + mr.a_lo, reinterpret_cast<const float*>(samples) + ti.s_off + ti.start + mr.a_lo, bytes, bar);
    }
    return mr;
  };
  auto manual_word = [&](const ManualRange& mr, int e) { return (e < mr.a_lo) ? e : mr.a_hi + (e - mr.a_lo); };
  auto manual_load = [&](const TileInfo& ti, int word) {
    return load_sample(samples + ti.s_off, (long long)reflect_index(ti.start + word, ti.L));
  };
  auto manual_fill_now = [&](const TileInfo& ti, const ManualRange& mr, float* buf) {
    for (int e = tid; e < mr.total; e += kThreads) {
      const int w = manual_word(mr, e);
      buf[w] = manual_load(ti, w);
    }
  };

  int tile = blockIdx.x;
  if (tile >= p.n_tiles) return;
#ifdef EVF_EXP_SKEW
  {
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)(warp >> 2) * EVF_EXP_SKEW) {
    }
  }
#endif
  TileInfo cur = tile_info(tile);
  TileInfo nxt = (tile + (int)gridDim.x < p.n_tiles) ? tile_info(tile + gridDim.x) : cur;
  {
    const ManualRange mr = stage_issue(cur, smem + p.off_in, &s_bar[0]);
    manual_fill_now(cur, mr, smem + p.off_in);
  }
  __syncthreads();

  for (int it = 0; tile < p.n_tiles; ++it, tile += gridDim.x) {
    const int b = (nbuf == 2) ? (it & 1) : 0;
    float* s_in = smem + (b ? p.off_in2 : p.off_in);
    const uint32_t parity = (nbuf == 2) ? ((it >> 1) & 1) : (it & 1);
    const int next_tile = tile + gridDim.x;
    const bool has_next = next_tile < p.n_tiles;
    const TileInfo fut = (next_tile + (int)gridDim.x < p.n_tiles) ? tile_info(next_tile + gridDim.x) : nxt;
    ManualRange nmr{0, 0, 0};
    float* s_next = smem + ((nbuf == 2 && !b) ? p.off_in2 : p.off_in);
    uint64_t* bar_next = &s_bar[(nbuf == 2) ? (b ^ 1) : 0];
    float mv0 = 0.f, mv1 = 0.f;
    bool deferred = false;
    if (has_next && nbuf == 2) {
      // prefetch the next tile into the other buffer; its few manual words ride in registers
      // across phase A (loads issued now, stores after the FFT)
      nmr = stage_issue(nxt, s_next, bar_next);
      if (nmr.total <= 2 * kThreads) {
        deferred = true;
        if (tid < nmr.total) mv0 = manual_load(nxt, manual_word(nmr, tid));
        if (tid + kThreads < nmr.total) mv1 = manual_load(nxt, manual_word(nmr, tid + kThreads));
      } else {
        manual_fill_now(nxt, nmr, s_next);
      }
    }
    const int nvalid = cur.nvalid;
    const long long out_frame0 = cur.out_frame0;

    mbar_wait(&s_bar[b], parity);

    // ---- phase A: one FFT job per warp ------------------------------------------------
    if (warp * FPJ < nvalid) {
      float re[32], im[32];
      if constexpr (MODE == MODE_PACK2) {
        // window pairs: s_win holds {w[32*r + lane], w[32*(r + 16) + lane]} at [r][lane] (16 LDS.64): rows r and
        // r + 16 are the two inputs of one first-stage butterfly, which absorbs the window multiplication
        const float* xa = s_in + (2 * warp) * hop + lane;
        const float2* wv = reinterpret_cast<const float2*>(s_win) + lane;
        if (hop == 256) {
          // the two frames of the job overlap by 768 samples: sample rows 8..31 of frame a ARE rows
          // 0..23 of frame b, so 40 row loads feed both frames (instead of 64)
          float v[40];
#pragma unroll
          for (int r = 0; r < 40; ++r) v[r] = xa[32 * r];
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const float2 w = wv[32 * r];
            const int i = bitrev5(r);
            win_head(re[i], re[i + 1], v[r], w.x, v[r + 16], w.y);
            win_head(im[i], im[i + 1], v[r + 8], w.x, v[r + 24], w.y);
          }
        } else {
          const float* xb = xa + hop;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const float2 w = wv[32 * r];
            const int i = bitrev5(r);
            win_head(re[i], re[i + 1], xa[32 * r], w.x, xa[32 * (r + 16)], w.y);
            win_head(im[i], im[i + 1], xb[32 * r], w.x, xb[32 * (r + 16)], w.y);
          }
        }
      } else {
        const float2* x2 = reinterpret_cast<const float2*>(s_in + warp * hop) + lane;
        const float2* w2 = reinterpret_cast<const float2*>(s_win) + lane;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float2 wa = w2[32 * r], wb = w2[32 * (r + 16)];
          const float2 xa = x2[32 * r], xb = x2[32 * (r + 16)];
          const int i = bitrev5(r);
          win_head(re[i], re[i + 1], xa.x, wa.x, xb.x, wb.x);
          win_head(im[i], im[i + 1], xa.y, wa.y, xb.y, wb.y);
        }
      }
      warp_fft1024_tail(re, im, s_tw4, scr, lane);

      // ---- real-FFT separation; the mirrored bin lives in lane (32 - lane) % 32 -------
      const int src_lane = (32 - lane) & 31;
      const int fa = warp * FPJ;  // tile-local frame of this job (PACK2: fa and fa + 1)
      float esum_a = 0.f, esum_b = 0.f;
      float* ga = p.spec_out + (out_frame0 + fa) * (long long)p.row_floats;
      float* gb = ga + p.row_floats;
      const bool b_valid = (MODE == MODE_PACK2) && (fa + 1 < nvalid);
      const int kcap = kMel ? p.k_used : (NFFT / 2 + 1);

#pragma unroll
      for (int j = 0; j <= 16; ++j) {
        const int k = lane + 32 * j;
        // warp-uniform: does any lane of this row own a bin that is consumed?
        bool need = 32 * j < kcap;
        if constexpr (kHalf) need = need || (1024 - 32 * j - 31 < kcap);
        if (need) {
          float zr, zi, pr, pi;
          if (j < 16) {
            zr = re[j];
            zi = im[j];
            // lane 0 is its own partner, with a different register (bin 32*(32-j) instead of 32*(31-j)+32-lane)
            const float sr = (lane == 0) ? re[(32 - j) & 31] : re[31 - j];
            const float si = (lane == 0) ? im[(32 - j) & 31] : im[31 - j];
            pr = __shfl_sync(0xffffffffu, sr, src_lane);
            pi = __shfl_sync(0xffffffffu, si, src_lane);
          } else {  // bin 512 (lane 0 only): its own mirror
            zr = pr = re[16];
            zi = pi = im[16];
          }
          if (j < 16 || lane == 0) {
            if constexpr (MODE == MODE_PACK2) {
              // window was pre-scaled by 1/2: X_a = Z[k] + conj(Z[N-k]), X_b = (Z[k] - conj(Z[N-k])) / i
              const float ar = zr + pr, ai = zi - pi;
              const float br = zi + pi, bi = pr - zr;
              if constexpr (SPEC == EVF_SPEC_RAW) {
                reinterpret_cast<float2*>(ga)[k] = make_float2(ar, ai);
                if (b_valid) reinterpret_cast<float2*>(gb)[k] = make_float2(br, bi);
              } else {
                float pa = fmaf(ar, ar, ai * ai);
                float pb = fmaf(br, br, bi * bi);
                if constexpr (SPEC == EVF_SPEC_MEL_LIBROSA) {
                  pa = fast_sqrt(pa + 1e-9f);
                  pb = fast_sqrt(pb + 1e-9f);
                }
                if constexpr (kMel) {
                  if (k < kcap) {
                    s_p[k * FS + fa] = pa;
                    s_p[k * FS + fa + 1] = pb;
                  }
                } else {
                  const float va = compress(pa, p.apply_log, p.log_clip);
                  const float vb = compress(pb, p.apply_log, p.log_clip);
                  ga[k] = va;
                  esum_a = fmaf(va, va, esum_a);
                  if (b_valid) {
                    gb[k] = vb;
                    esum_b = fmaf(vb, vb, esum_b);
                  }
                }
              }
            } else {
              // X[k] = E - T, X[M-k] = conj(E + T), E = Z[k] + conj(Z[M-k]), T = i * w_k * (Z[k] - conj(Z[M-k]))
              const float er = zr + pr, ei = zi - pi;
              const float orr = zr - pr, oi = zi + pi;
              // (cos, -sin)(2 pi k / 2048): from shared memory, or through L1 when the plan has no room to stage it
              const float2 w = kStageWpost ? s_wpost[k] : __ldg(p.wpost + k);
              const float tr = -fmaf(w.x, oi, w.y * orr);
              const float ti = fmaf(w.x, orr, -w.y * oi);
              const float x0r = er - tr, x0i = ei - ti;      // bin k
              const float x1r = er + tr, x1i = -(ei + ti);   // bin 1024 - k
              const int km = 1024 - k;
              constexpr bool kHasMirrorRow = true;
              const bool has_mirror = kHasMirrorRow && (j < 16);  // k == 512 is its own mirror
              if constexpr (SPEC == EVF_SPEC_RAW) {
                reinterpret_cast<float2*>(ga)[k] = make_float2(x0r, x0i);
                if (has_mirror) reinterpret_cast<float2*>(ga)[km] = make_float2(x1r, x1i);
              } else {
                float p0 = fmaf(x0r, x0r, x0i * x0i);
                float p1 = fmaf(x1r, x1r, x1i * x1i);
                if constexpr (SPEC == EVF_SPEC_MEL_LIBROSA) {
                  p0 = fast_sqrt(p0 + 1e-9f);
                  p1 = fast_sqrt(p1 + 1e-9f);
                }
                if constexpr (kMel) {
                  if (k < kcap) s_p[k * FS + fa] = p0;
                  if (has_mirror && km < kcap) s_p[km * FS + fa] = p1;
                } else {
                  const float v0 = compress(p0, p.apply_log, p.log_clip);
                  ga[k] = v0;
                  esum_a = fmaf(v0, v0, esum_a);
                  if (has_mirror) {
                    const float v1 = compress(p1, p.apply_log, p.log_clip);
                    ga[km] = v1;
                    esum_a = fmaf(v1, v1, esum_a);
                  }
                }
              }
            }
          }
        }
      }
      if constexpr (SPEC == EVF_SPEC_LINEAR) {
        if (p.energy_out != nullptr) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            esum_a += __shfl_xor_sync(0xffffffffu, esum_a, o);
            esum_b += __shfl_xor_sync(0xffffffffu, esum_b, o);
          }
          if (lane == 0) {
            p.energy_out[out_frame0 + fa] = sqrtf(esum_a);
            if (b_valid) p.energy_out[out_frame0 + fa + 1] = sqrtf(esum_b);
          }
        }
      }
    }

    // the next tile's manual words (reflected margins) were loaded before the FFT: store them now
    if (deferred) {
      if (tid < nmr.total) s_next[manual_word(nmr, tid)] = mv0;
      if (tid + kThreads < nmr.total) s_next[manual_word(nmr, tid + kThreads)] = mv1;
    }
#ifndef EVF_EXP_NOSYNC1
    __syncthreads();  // (1) input tile consumed, P tile complete, next tile's manual words visible
#endif
    if (has_next && nbuf == 1) {
      // single input buffer (large-footprint plans): restage now, overlapping phases B and C
      const ManualRange mr = stage_issue(nxt, s_in, &s_bar[0]);
      manual_fill_now(nxt, mr, s_in);
    }

#ifdef EVF_EXP_SKIP_BC
    if constexpr (false) {
#else
    if constexpr (kMel) {
#endif
      // ---- phase B: mel projection, lane = frame ---------------------------------------
      // Each worker (a warp, or a half / quarter warp when the tile has fewer than 32 frames)
      // walks a contiguous run of bins that starts and ends on interval boundaries.  Two running
      // FMAs per bin; when the interval index changes the two sums are flushed (fire-and-forget
      // stores, linear domain): SA[j] = sum of rising-slope weights * P over interval j,
      // SB[j] = the same for the falling slopes.  mel[m] = SA[m] + SB[m + 1] (phase C).
      {
        const float* __restrict__ sp = s_p;
        float* __restrict__ sa_out = s_sa;
        float* __restrict__ sb_out = s_sb;
        const float4* __restrict__ mw = s_mw4;  // per bin: {rising w, falling w, interval, interval of next bin}
        if constexpr (PARTS == 1) {
          // lane = frame, warp = worker: everything but the P value is warp-uniform
          const int wu = __shfl_sync(0xffffffffu, warp, 0);
          int k = s_vwk[wu];
          const int k1 = s_vwk[wu + 1];
          const float* pp = sp + k * FS + lane;
          const float4* wp = mw + k;
          float sa = 0.f, sb = 0.f;
          for (; k + 4 <= k1; k += 4, pp += 4 * FS, wp += 4) {
            const float p0 = pp[0], p1 = pp[FS], p2 = pp[2 * FS], p3 = pp[3 * FS];
            const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
#define EVF_BIN_STEP(W, P)                                                     \
  sa = fmaf(W.x, P, sa);                                                       \
  sb = fmaf(W.y, P, sb);                                                       \
  if (__float_as_int(W.z) != __float_as_int(W.w)) {                            \
    sa_out[__float_as_int(W.z) * FS + lane] = sa;                              \
    sb_out[__float_as_int(W.z) * FS + lane] = sb;                              \
    sa = 0.f;                                                                  \
    sb = 0.f;                                                                  \
  }
            EVF_BIN_STEP(w0, p0)
            EVF_BIN_STEP(w1, p1)
            EVF_BIN_STEP(w2, p2)
            EVF_BIN_STEP(w3, p3)
          }
          for (; k < k1; ++k, pp += FS, ++wp) {
            const float p0 = pp[0];
            const float4 w0 = wp[0];
            EVF_BIN_STEP(w0, p0)
          }
        } else {
          // several workers per warp (tiles of 16 or 8 frames): same walk, per-lane bin ranges
          const int fr = lane % FR;
          const int vw = warp * PARTS + lane / FR;
          int k = s_vwk[vw];
          const int k1 = s_vwk[vw + 1];
          float sa = 0.f, sb = 0.f;
          for (; k < k1; ++k) {
            const float pv = sp[k * FS + fr];
            const float4 w = mw[k];
            sa = fmaf(w.x, pv, sa);
            sb = fmaf(w.y, pv, sb);
            if (__float_as_int(w.z) != __float_as_int(w.w)) {
              sa_out[__float_as_int(w.z) * FS + fr] = sa;
              sb_out[__float_as_int(w.z) * FS + fr] = sb;
              sa = 0.f;
              sb = 0.f;
            }
          }
        }
#undef EVF_BIN_STEP
      }
#ifndef EVF_EXP_NOSYNC2
      __syncthreads();  // (2)
#endif
      // ---- phase C: combine, log, coalesced store of the log-mel rows + per-frame energy ---
      // (32-bit offsets from one row pointer per frame; rows of intervals without bins are never
      // flushed and stay zero, cleared at start)
      const int n_mels = p.n_mels;
      const int apply_log = p.apply_log;
      const float clip = p.log_clip;
#pragma unroll
      for (int q = 0; q < FPJ; ++q) {
        const int f = warp * FPJ + q;
        if (f < nvalid) {
          float* __restrict__ dst = p.spec_out + (out_frame0 + f) * (long long)p.row_floats;
          const float* sa_f = s_sa + f + lane * FS;
          const float* sb_f = s_sb + FS + f + lane * FS;
          float acc = 0.f;
          int m = lane;
#pragma unroll 1
          for (; m + 32 < n_mels; m += 64, sa_f += 64 * FS, sb_f += 64 * FS) {  // two rows of 32 per trip
            const float v0 = compress(sa_f[0] + sb_f[0], apply_log, clip);
            const float v1 = compress(sa_f[32 * FS] + sb_f[32 * FS], apply_log, clip);
            dst[m] = v0;
            dst[m + 32] = v1;
            acc = fmaf(v0, v0, acc);
            acc = fmaf(v1, v1, acc);
          }
          if (m < n_mels) {
            const float v0 = compress(sa_f[0] + sb_f[0], apply_log, clip);
            dst[m] = v0;
            acc = fmaf(v0, v0, acc);
          }
          if (p.energy_out != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) p.energy_out[out_frame0 + f] = sqrtf(acc);
          }
        }
      }
    } else {
      if (nbuf == 1) __syncthreads();  // the restaged manual words must be visible to the next FFT
    }
    cur = nxt;
    nxt = fut;
  }
}

template <int MODE, int SPEC, typename SampleT, int WARPS>
int launch_w(const FeatParams& p, int grid, int smem, cudaStream_t stream, bool configure_only) {
  auto kern = features_kernel<MODE, SPEC, SampleT, WARPS>;
  if (configure_only) {
    EVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return EVF_OK;
  }
  kern<<<grid, WARPS * 32, smem, stream>>>(p);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

template <int MODE, int SPEC, typename SampleT>
int launch_t(int warps, const FeatParams& p, int grid, int smem, cudaStream_t stream, bool cfg) {
  if (warps == 8) return launch_w<MODE, SPEC, SampleT, 8>(p, grid, smem, stream, cfg);
  return launch_w<MODE, SPEC, SampleT, 16>(p, grid, smem, stream, cfg);
}

template <int MODE, int SPEC>
int launch_s(int fmt, int warps, const FeatParams& p, int grid, int smem, cudaStream_t st, bool cfg) {
  if (fmt == EVF_SAMPLES_S16) return launch_t<MODE, SPEC, short>(warps, p, grid, smem, st, cfg);
  return launch_t<MODE, SPEC, float>(warps, p, grid, smem, st, cfg);
}

template <int MODE>
int launch_m(int spec, int fmt, int warps, const FeatParams& p, int grid, int smem, cudaStream_t st, bool cfg) {
  switch (spec) {
    case EVF_SPEC_MEL: return launch_s<MODE, EVF_SPEC_MEL>(fmt, warps, p, grid, smem, st, cfg);
    case EVF_SPEC_MEL_LIBROSA: return launch_s<MODE, EVF_SPEC_MEL_LIBROSA>(fmt, warps, p, grid, smem, st, cfg);
    case EVF_SPEC_LINEAR: return launch_s<MODE, EVF_SPEC_LINEAR>(fmt, warps, p, grid, smem, st, cfg);
    case EVF_SPEC_RAW: return launch_s<MODE, EVF_SPEC_RAW>(fmt, warps, p, grid, smem, st, cfg);
  }
  set_error("unknown spec_type");
  return EVF_ERR_UNSUPPORTED;
}

int dispatch(int mode, int spec, int fmt, int warps, const FeatParams& p, int grid, int smem, cudaStream_t st,
             bool cfg) {
  if (mode == MODE_PACK2) return launch_m<MODE_PACK2>(spec, fmt, warps, p, grid, smem, st, cfg);
  if (mode == MODE_HALF) return launch_m<MODE_HALF>(spec, fmt, warps, p, grid, smem, st, cfg);
  if (mode == MODE_HALF_L1) return launch_m<MODE_HALF_L1>(spec, fmt, warps, p, grid, smem, st, cfg);
  set_error("unknown FFT mode");
  return EVF_ERR_UNSUPPORTED;
}

}  // namespace

// Computes the shared-memory carve-up for a plan; returns bytes (or -1 if it cannot fit).
int features_smem_bytes(int mode, int spec_type, int warps, int hop, int n_fft, int n_mels, int k_used,
                        FeatParams* c) {
  const bool mel = (spec_type == EVF_SPEC_MEL || spec_type == EVF_SPEC_MEL_LIBROSA);
  const int fpj = (mode == MODE_PACK2) ? 2 : 1;
  const int fr = warps * fpj;
  const int parts = 32 / fr;
  // 16 warps: one CTA per SM (227 KB); 8 warps: two CTAs per SM (228 KB - 2 x 1 KB reserved, halved)
  const long long limit = (warps == 16) ? 227 * 1024 : 113 * 1024;
  auto up4 = [](int w) { return (w + 3) & ~3; };
  // Prefer two input buffers (the next tile's bulk copy overlaps this tile's FFTs); fall back
  // to one when the plan's tables do not leave room for it.  (The caller retries with
  // MODE_HALF_L1 -- post-twiddles read through L1 instead of staged -- when MODE_HALF cannot fit.)
  const int stage_wpost = (mode == MODE_HALF) ? 1 : 0;
  for (int nbuf = 2; nbuf >= 1; --nbuf) {
    int w = 0;
    c->nbuf = nbuf;
    c->off_bar = w;
    w += 4;  // two 8-byte mbarriers
    c->off_in = w;
    c->in_words = up4((fr - 1) * hop + n_fft);
    w += c->in_words;
    c->off_in2 = c->off_in;
    if (nbuf == 2) {
      c->off_in2 = w;
      w += c->in_words;
    }
    c->off_win = w;
    w += up4(n_fft);
    c->off_tw = w;
    w += 2 * kFftSize;
    c->off_wpost = w;
    if (stage_wpost) w += up4(2 * 513);
    c->off_melw = w;
    c->off_vwk = w;
    c->off_p = w;
    c->off_sa = w;
    c->off_sb = w;
    if (mel) {
      w += 4 * k_used;
      c->off_vwk = w;
      w += up4(warps * parts + 1);
      c->off_p = w;
      w += up4(k_used * (fr + 1));
      c->off_sa = w;
      w += (n_mels + 1) * (fr + 1);
      c->off_sb = w;  // directly after SA (cleared together)
      w += (n_mels + 1) * (fr + 1);
      w = up4(w);
    }
    c->off_scr = w;
    w += warps * 32 * kScrStride;
    const long long bytes = 4ll * w;
    if (bytes <= limit) return (int)bytes;
  }
  return -1;
}

int features_configure(int mode, int spec_type, int sample_format, int warps, int smem_bytes) {
  FeatParams dummy{};
  return dispatch(mode, spec_type, sample_format, warps, dummy, 1, smem_bytes, nullptr, true);
}

int features_launch(int mode, int spec_type, int sample_format, int warps, const FeatParams& p, int grid,
                    int smem_bytes, cudaStream_t stream) {
  return dispatch(mode, spec_type, sample_format, warps, p, grid, smem_bytes, stream, false);
}

}  // namespace evf
