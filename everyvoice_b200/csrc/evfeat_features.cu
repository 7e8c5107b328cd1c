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
//   grid   : persistent, two CTAs of 8 warps per SM (the register file allows 16 warps; two small CTAs
//            instead of one: shorter tiles leave fewer idle job slots at utterance ends and the two
//            rings drift independently), static round-robin over frame tiles
//   tile   : 8 FFT jobs of one utterance = 16 consecutive frames (n_fft 1024: two real
//            frames ride as re/im of one complex FFT) or 8 frames (n_fft 2048: one frame
//            = 1024 complex points + real-FFT split); warp w owns job slot (w + iteration) % 8
//   ring   : the tile's sample span ((FR-1)*hop + n_fft samples; every sample leaves HBM once,
//            the 4x frame overlap is served from shared memory) is fetched with cp.async.bulk
//            (TMA 1-D) into a ring of input buffers; "full" is an mbarrier per buffer, "empty"
//            is a counter: a warp releases the buffer as soon as its samples sit in registers
//            (before the FFT), and the LAST warp to release refills it with the tile after
//            next (reflect padding at the utterance edges = index mirroring on the few margin
//            words, by that warp).  There is NO block-wide barrier in the steady state: the
//            warps drift apart, so one warp's shared-memory phases (sample loads, the
//            transpose, the mel walk) overlap the FMA-bound butterflies of the others.
//   job    : window * samples -> registers, 32x32 four-step FFT (in-register radix-2 DFT-32,
//            transpose + twiddle through the warp's private scratch), real-FFT separation with
//            the mirrored bin fetched by warp shuffle, |X|^2 (or sqrt(|X|^2+1e-9) for
//            mel-librosa) -> the warp's private P column (aliases the transpose scratch);
//            linear / raw write straight to HBM (lanes = consecutive bins, coalesced)
//   mel    : per warp, no cross-warp traffic.  Every bin lies between two adjacent filter
//            centres, so it feeds exactly two filters.  Step 1 (lane = bin chunk): lane l walks
//            bins [n*l, n*l + n), n odd so that the 32 lanes read 32 distinct banks, two running
//            FMAs per bin and frame; when the interval between two centres ends the sums are
//            flushed to a slot (the lane that starts an interval owns its slot, a lane whose
//            chunk begins inside an interval flushes that first partial to its head slot).
//            Step 2 (lane = filter): mel[m] = rising partials of interval m + falling partials
//            of interval m + 1, gathered in a fixed order through a host-built slot list, then
//            log(max(., clip)), coalesced row stores and the per-frame energy sqrt(sum log^2).
#include "evfeat_fft.cuh"
#include "evfeat_internal.h"

namespace evf {

namespace {

#include "evfeat_device.cuh"

using TileInfo = TileDesc;  // host-built, 32 bytes, one per tile (evfeat_internal.h)

// compile-time loop indices for the fully unrolled register-array code
template <int V>
struct IntC {
  static constexpr int value = V;
};
template <int B, int E, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (B < E) {
    f(IntC<B>{});
    static_for<B + 1, E>(f);
  }
}

template <int MODE, int SPEC, typename SampleT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 16 / WARPS) features_kernel(const FeatParams p) {
  using MT = ModeTraits<MODE>;
  constexpr int kThreads = WARPS * 32;
  constexpr int NFFT = MT::kNfft;
  constexpr int FPJ = MT::kFramesPerJob;
  constexpr int J = MT::kJobs;        // packed jobs side by side in the warp's register file (n_fft 512: 2, 256: 4)
  constexpr int R1 = 32 / J;          // first-pass rows of a job = lanes that own a job's bins after the second pass
  constexpr int FPW = FPJ * J;        // frames a warp owns per tile
  constexpr bool kPack = (MODE != MODE_HALF);
  constexpr bool kMel = (SPEC == EVF_SPEC_MEL || SPEC == EVF_SPEC_MEL_LIBROSA);
  constexpr bool kHalf = (MODE == MODE_HALF);           // n_fft 2048: one frame per FFT + real-FFT split
  using SlotT = typename MT::SlotT;                      // {rise a, rise b, fall a, fall b} or {rise, fall}

  extern __shared__ __align__(16) float smem[];
  float* s_win = smem + p.off_win;
  float4* s_tw4 = reinterpret_cast<float4*>(smem + p.off_tw);
  const float2* s_wpost = reinterpret_cast<const float2*>(smem + p.off_wpost);
  const float2* s_wtab = reinterpret_cast<const float2*>(smem + p.off_wtab);
  const uint2* s_gtab = reinterpret_cast<const uint2*>(smem + p.off_gtab);
  const unsigned* s_ltab = reinterpret_cast<const unsigned*>(smem + p.off_ltab);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);  // full[0], full[1]
  int* s_cnt = reinterpret_cast<int*>(smem + p.off_bar + 4);         // released-by counters

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  float* scr = smem + p.off_warp + warp * p.warp_words;  // transpose scratch; the P column aliases it
  // the P columns of J > 1 jobs (each padded to the walk's n_chunk * 32 bins) may outgrow the transpose scratch
  const int scr_words = (J == 1) ? 32 * kScrStride : p.scr_words;
  SlotT* slots = reinterpret_cast<SlotT*>(scr + scr_words);
  const int nbuf = p.nbuf;

  // ---- one-time setup: mbarriers, tables, zeroed scratch / slots ---------------------------
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    s_cnt[0] = 0;
    s_cnt[1] = 0;
    fence_mbar_init();
  }
  for (int i = tid; i < NFFT; i += kThreads) s_win[i] = p.window[i];
  for (int i = tid; i < kFftSize / 2; i += kThreads) s_tw4[i] = p.tw4[i];
  if constexpr (kHalf) {
    for (int i = tid; i <= 512; i += kThreads) reinterpret_cast<float2*>(smem + p.off_wpost)[i] = p.wpost[i];
  }
  if constexpr (kMel) {
    for (int i = tid; i < p.n_chunk * 32; i += kThreads) reinterpret_cast<float2*>(smem + p.off_wtab)[i] = p.wtab[i];
    for (int i = tid; i < (p.n_heads + 1) * p.m_pad; i += kThreads)
      reinterpret_cast<uint2*>(smem + p.off_gtab)[i] = p.gtab[i];
    if (tid < 32) reinterpret_cast<unsigned*>(smem + p.off_ltab)[tid] = p.ltab[tid];
  }
  // pad words of the scratch and never-flushed slots (empty intervals, the zero slot) stay zero
  for (int i = tid; i < WARPS * p.warp_words; i += kThreads) smem[p.off_warp + i] = 0.f;

  const int hop = p.hop;
  const SampleT* __restrict__ samples = static_cast<const SampleT*>(p.samples);
  const int G = gridDim.x;

    auto tile_info = [&](int tile) { return load_tile_desc(p.tiles, tile); };
  auto stage_manual = [&](const TileInfo& ti, SampleT* buf, int t, int nt, int& a_lo, int& a_hi) {
    evf_stage_manual(samples, ti, buf, t, nt, a_lo, a_hi);
  };
  auto stage_bulk = [&](const TileInfo& ti, SampleT* buf, uint64_t* bar, int a_lo, int a_hi) {
    evf_stage_bulk(samples, ti, buf, bar, a_lo, a_hi);
  };
  auto in_buf = [&](int b) { return reinterpret_cast<SampleT*>(smem + (b ? p.off_in2 : p.off_in)); };

  int tile = blockIdx.x;  // grid <= n_tiles
  TileInfo cur = tile_info(tile);
  TileInfo nxt = (tile + G < p.n_tiles) ? tile_info(tile + G) : cur;
  __syncthreads();  // barriers initialised, tables in place
  {
    int a_lo, a_hi;
    stage_manual(cur, in_buf(0), tid, kThreads, a_lo, a_hi);
    int b_lo = 0, b_hi = 0;
    const bool second = (nbuf == 2) && (tile + G < p.n_tiles);
    if (second) stage_manual(nxt, in_buf(1), tid, kThreads, b_lo, b_hi);
    __syncthreads();
    if (tid == 0) {
      stage_bulk(cur, in_buf(0), &s_bar[0], a_lo, a_hi);
      if (second) stage_bulk(nxt, in_buf(1), &s_bar[1], b_lo, b_hi);
    }
  }

  // per-lane constants of the mel walk
  unsigned lt = 0;
  if constexpr (kMel) lt = s_ltab[lane];

  for (int it = 0; tile < p.n_tiles; ++it, tile += G) {
    const int b = (nbuf == 2) ? (it & 1) : 0;
    const SampleT* s_in = in_buf(b);
    const uint32_t parity = (nbuf == 2) ? ((it >> 1) & 1) : (it & 1);
    // clamped index instead of a select on the loaded values: nothing consumes the descriptor before the refill
    // (or the end of the iteration), so the load latency stays off the critical path; past the end it is unused
    const TileInfo fut = tile_info(min(tile + 2 * G, p.n_tiles - 1));
    const int nvalid = cur.nvalid;
    const long long out_frame0 = cur.out_frame0;
    // the warp's job slot rotates with the iteration: the idle slots of partial tiles (utterance ends) are spread over
    // the warps instead of always hitting the high-numbered ones
    const int slot = (warp + it) & (WARPS - 1);
    const bool active = slot * FPW < nvalid;

    mbar_wait(&s_bar[b], parity);

    float re[32], im[32];
    if (active) {
      // ---- samples -> registers, window fused into the first butterfly stage -------------
      if constexpr (J > 1) {
        // J packed jobs of R1 rows: register block jj holds job jj (frames FPW * warp + 2 * jj, + 1 as re / im);
        // window pairs {w[32 * r + lane], w[32 * (r + R1 / 2) + lane]} at [r][lane], r < R1 / 2
        constexpr int HR = R1 / 2;
        constexpr int LB = (R1 == 16) ? 3 : 2;  // log2(HR)
        const SampleT* x0 = s_in + (FPW * slot) * hop + lane;
        const float2* wv = reinterpret_cast<const float2*>(s_win) + lane;
        if (hop * 4 == NFFT) {
          // hop = n_fft / 4 = R1 / 4 sample rows: the warp's FPW frames share rows, (FPW - 1) * R1 / 4 + R1 row loads
          // feed all of them (n_fft 512: 28 instead of 64)
          constexpr int RH = R1 / 4;
          constexpr int NV = (FPW - 1) * RH + R1;
          float v[NV];
#pragma unroll
          for (int r = 0; r < NV; ++r) v[r] = to_float(x0[32 * r]);
#pragma unroll
          for (int jj = 0; jj < J; ++jj) {
#pragma unroll
            for (int r = 0; r < HR; ++r) {
              const float2 w = wv[32 * r];
              const int i = jj * R1 + 2 * bitrev_n(r, LB);
              win_head(re[i], re[i + 1], v[(2 * jj) * RH + r], w.x, v[(2 * jj) * RH + r + HR], w.y);
              win_head(im[i], im[i + 1], v[(2 * jj + 1) * RH + r], w.x, v[(2 * jj + 1) * RH + r + HR], w.y);
            }
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < J; ++jj) {
            const SampleT* xa = x0 + (2 * jj) * hop;
            const SampleT* xb = xa + hop;
#pragma unroll
            for (int r = 0; r < HR; ++r) {
              const float2 w = wv[32 * r];
              const int i = jj * R1 + 2 * bitrev_n(r, LB);
              win_head(re[i], re[i + 1], to_float(xa[32 * r]), w.x, to_float(xa[32 * (r + HR)]), w.y);
              win_head(im[i], im[i + 1], to_float(xb[32 * r]), w.x, to_float(xb[32 * (r + HR)]), w.y);
            }
          }
        }
      } else if constexpr (MODE == MODE_PACK2) {
        // window pairs: s_win holds {w[32*r + lane], w[32*(r + 16) + lane]} at [r][lane] (16 LDS.64): rows r and
        // r + 16 are the two inputs of one first-stage butterfly, which absorbs the window multiplication
        const SampleT* xa = s_in + (2 * slot) * hop + lane;
        const float2* wv = reinterpret_cast<const float2*>(s_win) + lane;
        if (hop == 256) {
          // the two frames of the job overlap by 768 samples: sample rows 8..31 of frame a ARE rows
          // 0..23 of frame b, so 40 row loads feed both frames (instead of 64)
          float v[40];
#pragma unroll
          for (int r = 0; r < 40; ++r) v[r] = to_float(xa[32 * r]);
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const float2 w = wv[32 * r];
            const int i = bitrev5(r);
            win_head(re[i], re[i + 1], v[r], w.x, v[r + 16], w.y);
            win_head(im[i], im[i + 1], v[r + 8], w.x, v[r + 24], w.y);
          }
        } else {
          const SampleT* xb = xa + hop;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const float2 w = wv[32 * r];
            const int i = bitrev5(r);
            win_head(re[i], re[i + 1], to_float(xa[32 * r]), w.x, to_float(xa[32 * (r + 16)]), w.y);
            win_head(im[i], im[i + 1], to_float(xb[32 * r]), w.x, to_float(xb[32 * (r + 16)]), w.y);
          }
        }
      } else {
        const SampleT* x2 = s_in + slot * hop + 2 * lane;
        const float2* w2 = reinterpret_cast<const float2*>(s_win) + lane;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float2 wa = w2[32 * r], wb = w2[32 * (r + 16)];
          const float2 xa = load_pair(x2 + 64 * r), xb = load_pair(x2 + 64 * (r + 16));
          const int i = bitrev5(r);
          win_head(re[i], re[i + 1], xa.x, wa.x, xb.x, wb.x);
          win_head(im[i], im[i + 1], xa.y, wa.y, xb.y, wb.y);
        }
      }
    }

    // ---- release the input buffer; the last warp to do so refills it ---------------------
    {
      const int refill_tile = tile + nbuf * G;
      int last = 0;
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();  // this warp's reads of the buffer are done before the release is visible
        last = (atomicAdd(&s_cnt[b], 1) == WARPS - 1);
        __threadfence_block();  // ... and the other warps' releases are visible before the refill
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        if (lane == 0) s_cnt[b] = 0;
        if (refill_tile < p.n_tiles) {
          const TileInfo& rt = (nbuf == 2) ? fut : nxt;
          SampleT* dst = in_buf(b);
          int a_lo, a_hi;
          stage_manual(rt, dst, lane, 32, a_lo, a_hi);
          __syncwarp();
          if (lane == 0) stage_bulk(rt, dst, &s_bar[b], a_lo, a_hi);
        }
      }
    }

    if (active) {
      warp_fft1024_tail<J>(re, im, s_tw4, scr, lane);

      // ---- real-FFT separation; the mirrored bin lives in lane (32 - lane) % 32 (of the job's R1 lanes) -------
      const int k1 = lane & (R1 - 1);  // bin k = k1 + R1 * j of the lane's job
      const int jw = lane / R1;        // the lane's job (J > 1)
      const int src_lane = (lane & ~(R1 - 1)) | ((R1 - k1) & (R1 - 1));
      const int fa = slot * FPW + FPJ * jw;  // tile-local frame of this lane's job (packed: fa and fa + 1)
      float esum_a = 0.f, esum_b = 0.f;
      float* ga = p.spec_out + (out_frame0 + fa) * (long long)p.row_floats;
      float* gb = ga + p.row_floats;
      const bool a_valid = (J == 1) || (fa < nvalid);
      const bool b_valid = kPack && (fa + 1 < nvalid);
      // mel: every bin the walk below reads is rewritten by this job (zero-weight bins past k_used
      // included), so nothing stale -- a NaN of an earlier job -- can leak into this one
      const int kcap = kMel ? min(p.n_chunk * 32, NFFT / 2 + 1) : (NFFT / 2 + 1);
      // the P column of this job: PACK2 float2 {frame a, frame b} per bin, HALF one float per bin
      const int pstride = kMel ? p.n_chunk * 32 : 0;  // bins between the P columns of a warp's jobs (J > 1)
      float2* P2 = reinterpret_cast<float2*>(scr) + ((J > 1) ? jw * pstride : 0);
      float* P1 = scr;
      if constexpr (kMel && J > 1) {
        // the walk reads n_chunk * 32 bins per job; those past the last bin carry zero weights but must not hold a
        // NaN of another job's transposed data
        for (int k = kcap + lane; k < pstride; k += 32) {
#pragma unroll
          for (int jj = 0; jj < J; ++jj) reinterpret_cast<float2*>(scr)[jj * pstride + k] = make_float2(0.f, 0.f);
        }
      }

      // one row of the separation: bins k = k1 + R1 * j of the warp's jobs (j is a compile-time register index)
      auto sep_row = [&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const int k = k1 + R1 * j;
        float zr, zi, pr, pi;
        if (j < 16) {
          zr = re[j];
          zi = im[j];
          // lane 0 is its own partner, with a different register (bin 32*(32-j) instead of 32*(31-j)+32-lane)
          const float sr = (k1 == 0) ? re[(32 - j) & 31] : re[31 - j];
          const float si = (k1 == 0) ? im[(32 - j) & 31] : im[31 - j];
          pr = __shfl_sync(0xffffffffu, sr, src_lane);
          pi = __shfl_sync(0xffffffffu, si, src_lane);
        } else {  // bin 512 (lane 0 only): its own mirror
          zr = pr = re[16];
          zi = pi = im[16];
        }
        if (j < 16 || k1 == 0) {
          if constexpr (kPack) {
            // window was pre-scaled by 1/2: X_a = Z[k] + conj(Z[N-k]), X_b = (Z[k] - conj(Z[N-k])) / i
            const float ar = zr + pr, ai = zi - pi;
            const float br = zi + pi, bi = pr - zr;
            if constexpr (SPEC == EVF_SPEC_RAW) {
              if (a_valid) reinterpret_cast<float2*>(ga)[k] = make_float2(ar, ai);
              if (b_valid) reinterpret_cast<float2*>(gb)[k] = make_float2(br, bi);
            } else {
              float pa = fmaf(ar, ar, ai * ai);
              float pb = fmaf(br, br, bi * bi);
              if constexpr (SPEC == EVF_SPEC_MEL_LIBROSA) {
                pa = fast_sqrt(pa + 1e-9f);
                pb = fast_sqrt(pb + 1e-9f);
              }
              if constexpr (kMel) {
                P2[k] = make_float2(pa, pb);
              } else {
                const float va = compress(pa, p.apply_log, p.log_clip);
                const float vb = compress(pb, p.apply_log, p.log_clip);
                if (a_valid) ga[k] = va;
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
            const float2 w = s_wpost[k];  // (cos, -sin)(2 pi k / 2048)
            const float tr = -fmaf(w.x, oi, w.y * orr);
            const float ti = fmaf(w.x, orr, -w.y * oi);
            const float x0r = er - tr, x0i = ei - ti;      // bin k
            const float x1r = er + tr, x1i = -(ei + ti);   // bin 1024 - k
            const int km = 1024 - k;
            const bool has_mirror = (j < 16);  // k == 512 is its own mirror
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
                if (k < kcap) P1[k] = p0;
                if (has_mirror && km < kcap) P1[km] = p1;
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
      };
      if constexpr (kMel && !kHalf) {
        // rows j < ceil(kcap / R1) are consumed (warp-uniform).  One computed jump into the unrolled rows, highest row
        // first, instead of a test per row (the rows only store to the P column: their order is free)
        switch ((kcap + R1 - 1) / R1) {
          default: sep_row(IntC<16>{}); [[fallthrough]];
          case 16: sep_row(IntC<15>{}); [[fallthrough]];
          case 15: sep_row(IntC<14>{}); [[fallthrough]];
          case 14: sep_row(IntC<13>{}); [[fallthrough]];
          case 13: sep_row(IntC<12>{}); [[fallthrough]];
          case 12: sep_row(IntC<11>{}); [[fallthrough]];
          case 11: sep_row(IntC<10>{}); [[fallthrough]];
          case 10: sep_row(IntC<9>{}); [[fallthrough]];
          case 9: sep_row(IntC<8>{}); [[fallthrough]];
          case 8: sep_row(IntC<7>{}); [[fallthrough]];
          case 7: sep_row(IntC<6>{}); [[fallthrough]];
          case 6: sep_row(IntC<5>{}); [[fallthrough]];
          case 5: sep_row(IntC<4>{}); [[fallthrough]];
          case 4: sep_row(IntC<3>{}); [[fallthrough]];
          case 3: sep_row(IntC<2>{}); [[fallthrough]];
          case 2: sep_row(IntC<1>{}); [[fallthrough]];
          case 1: sep_row(IntC<0>{});
        }
      } else {
        static_for<0, 17>([&](auto jc) {
          constexpr int j = decltype(jc)::value;
          // warp-uniform: does any lane of this row own a bin that is consumed?
          bool need = R1 * j < kcap;
          if constexpr (kHalf) need = need || (1024 - 32 * j - 31 < kcap);
          if (need) sep_row(jc);
        });
      }

      if constexpr (kMel) {
        // ---- mel step 1: lane = chunk of n bins; two running FMAs per bin and frame ---------
        // (J > 1: the warp's jobs take turns; the slots are reused, the P columns lie pstride bins apart)
        __syncwarp();
        const int n = p.n_chunk;
#pragma unroll 1
        for (int jj = 0; jj < J; ++jj) {
          const int fj = (J == 1) ? fa : slot * FPW + FPJ * jj;  // first frame of job jj (warp-uniform)
          if (J > 1 && fj >= nvalid) break;
          float* gja = (J == 1) ? ga : p.spec_out + (out_frame0 + fj) * (long long)p.row_floats;
          float* gjb = gja + p.row_floats;
          const bool jb_valid = (J == 1) ? b_valid : (kPack && fj + 1 < nvalid);
          float ea = 0.f, eb = 0.f;
        {
          const float2* wp = s_wtab + lane;
          SlotT* dst = slots + (lt & 0xffffu);
          SlotT* dst_next = slots + (lt >> 16);
          if constexpr (kPack) {
            const float2* pp = ((J == 1) ? P2 : reinterpret_cast<const float2*>(scr) + jj * pstride) + n * lane;
            float ra = 0.f, rb = 0.f, fa_ = 0.f, fb = 0.f;
#pragma unroll 4
            for (int i = 0; i < n; ++i, ++pp, wp += 32) {
              const float2 pv = *pp;
              const float2 w = *wp;
              const float wr = fabsf(w.x);
              ra = fmaf(wr, pv.x, ra);
              rb = fmaf(wr, pv.y, rb);
              fa_ = fmaf(w.y, pv.x, fa_);
              fb = fmaf(w.y, pv.y, fb);
              if (__float_as_int(w.x) < 0) {  // the interval (or the chunk) ends with this bin
                *dst = make_float4(ra, rb, fa_, fb);
                dst = dst_next;
                ++dst_next;
                ra = rb = fa_ = fb = 0.f;
              }
            }
          } else {
            const float* pp = P1 + n * lane;
            float ra = 0.f, fa_ = 0.f;
#pragma unroll 2  // measured: 2 beats 4 and 8 here (n_fft 2048: 1.273 -> 1.245 ms), 4 is best for the packed loop above
            for (int i = 0; i < n; ++i, ++pp, wp += 32) {
              const float pv = *pp;
              const float2 w = *wp;
              ra = fmaf(fabsf(w.x), pv, ra);
              fa_ = fmaf(w.y, pv, fa_);
              if (__float_as_int(w.x) < 0) {
                *dst = make_float2(ra, fa_);
                dst = dst_next;
                ++dst_next;
                ra = fa_ = 0.f;
              }
            }
          }
        }
        __syncwarp();
        // ---- mel step 2: lane = filter; gather, log, coalesced row stores, energy -----------
        const int n_mels = p.n_mels;
        const int apply_log = p.apply_log;
        const float clip = p.log_clip;
        const int nh = p.n_heads;
        const int m_pad = p.m_pad;
        // the number of partials per interval (nh + 1) is a property of the plan: the common small counts get a fully
        // unrolled row body (no inner loop control), anything else the run-time loop.  Fixed summation order either way.
        auto gather_rows = [&](auto cc) {
          constexpr int C = decltype(cc)::value;  // partials per interval, 0 = run-time count
          const char* sb = reinterpret_cast<const char*>(slots);
          const uint2* gp0 = s_gtab + lane;
          float* oa = gja + lane;
          float* ob = gjb + lane;
          for (int m = lane; m < n_mels; m += 32, gp0 += 32, oa += 32, ob += 32) {
            const uint2* gp = gp0;
            if constexpr (kPack) {
              float va = 0.f, vb = 0.f;
              auto add = [&](const uint2 g) {  // byte offsets of {rise a, rise b} and of {fall a, fall b}
                const float2 r = *reinterpret_cast<const float2*>(sb + g.x);
                const float2 f = *reinterpret_cast<const float2*>(sb + g.y);
                va += r.x;
                vb += r.y;
                va += f.x;
                vb += f.y;
              };
              if constexpr (C > 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) add(gp[c * m_pad]);
              } else {
                for (int c = 0; c <= nh; ++c, gp += m_pad) add(*gp);
              }
              va = compress(va, apply_log, clip);
              vb = compress(vb, apply_log, clip);
              *oa = va;
              ea = fmaf(va, va, ea);
              if (jb_valid) {
                *ob = vb;
                eb = fmaf(vb, vb, eb);
              }
            } else {
              float va = 0.f;
              auto add = [&](const uint2 g) {
                va += *reinterpret_cast<const float*>(sb + g.x);
                va += *reinterpret_cast<const float*>(sb + g.y);
              };
              if constexpr (C > 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) add(gp[c * m_pad]);
              } else {
                for (int c = 0; c <= nh; ++c, gp += m_pad) add(*gp);
              }
              va = compress(va, apply_log, clip);
              *oa = va;
              ea = fmaf(va, va, ea);
            }
          }
        };
        switch (nh) {
          case 0: gather_rows(IntC<1>{}); break;
          case 1: gather_rows(IntC<2>{}); break;
          case 2: gather_rows(IntC<3>{}); break;
          default: gather_rows(IntC<0>{}); break;
        }
          if constexpr (J == 1) {
            esum_a = ea;
            esum_b = eb;
          } else {
            // every lane holds filters of job jj: reduce over the warp, then the next job reuses the slots
            if (p.energy_out != nullptr) {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                ea += __shfl_xor_sync(0xffffffffu, ea, o);
                eb += __shfl_xor_sync(0xffffffffu, eb, o);
              }
              if (lane == 0) {
                p.energy_out[out_frame0 + fj] = sqrtf(ea);
                if (jb_valid) p.energy_out[out_frame0 + fj + 1] = sqrtf(eb);
              }
            }
            __syncwarp();
          }
        }
      }

      if constexpr (SPEC != EVF_SPEC_RAW && !(kMel && J > 1)) {
        if (p.energy_out != nullptr) {
#pragma unroll
          for (int o = R1 / 2; o > 0; o >>= 1) {
            esum_a += __shfl_xor_sync(0xffffffffu, esum_a, o);
            if constexpr (kPack) esum_b += __shfl_xor_sync(0xffffffffu, esum_b, o);
          }
          if (k1 == 0) {
            if (a_valid) p.energy_out[out_frame0 + fa] = sqrtf(esum_a);
            if (b_valid) p.energy_out[out_frame0 + fa + 1] = sqrtf(esum_b);
          }
        }
      }
    }
    cur = nxt;
    nxt = fut;
  }
}

template <int MODE, int SPEC, typename SampleT>
int launch_t(const FeatParams& p, int grid, int smem, cudaStream_t stream, bool configure_only) {
  auto kern = features_kernel<MODE, SPEC, SampleT, kMaxWarps>;
  if (configure_only) {
    // the attribute belongs to the kernel, not to a plan: plans of the same instantiation with different carve-ups
    // (hops) coexist, so it is set to the opt-in maximum once instead of to this plan's size
    (void)smem;
    EVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return EVF_OK;
  }
  kern<<<grid, kMaxWarps * 32, smem, stream>>>(p);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

template <int MODE, int SPEC>
int launch_s(int fmt, const FeatParams& p, int grid, int smem, cudaStream_t st, bool cfg) {
  if (fmt == EVF_SAMPLES_S16) return launch_t<MODE, SPEC, short>(p, grid, smem, st, cfg);
  return launch_t<MODE, SPEC, float>(p, grid, smem, st, cfg);
}

template <int MODE>
int launch_m(int spec, int fmt, const FeatParams& p, int grid, int smem, cudaStream_t st, bool cfg) {
  switch (spec) {
    case EVF_SPEC_MEL: return launch_s<MODE, EVF_SPEC_MEL>(fmt, p, grid, smem, st, cfg);
    case EVF_SPEC_MEL_LIBROSA: return launch_s<MODE, EVF_SPEC_MEL_LIBROSA>(fmt, p, grid, smem, st, cfg);
    case EVF_SPEC_LINEAR: return launch_s<MODE, EVF_SPEC_LINEAR>(fmt, p, grid, smem, st, cfg);
    case EVF_SPEC_RAW: return launch_s<MODE, EVF_SPEC_RAW>(fmt, p, grid, smem, st, cfg);
  }
  set_error("unknown spec_type");
  return EVF_ERR_UNSUPPORTED;
}

int dispatch(int mode, int spec, int fmt, const FeatParams& p, int grid, int smem, cudaStream_t st, bool cfg) {
  if (mode == MODE_PACK2) return launch_m<MODE_PACK2>(spec, fmt, p, grid, smem, st, cfg);
  if (mode == MODE_HALF) return launch_m<MODE_HALF>(spec, fmt, p, grid, smem, st, cfg);
  if (mode == MODE_PACK2_512) return launch_m<MODE_PACK2_512>(spec, fmt, p, grid, smem, st, cfg);
  if (mode == MODE_PACK2_256) return launch_m<MODE_PACK2_256>(spec, fmt, p, grid, smem, st, cfg);
  set_error("unknown FFT mode");
  return EVF_ERR_UNSUPPORTED;
}

}  // namespace

// Computes the shared-memory carve-up for a plan; returns bytes (or -1 if it cannot fit).
int features_smem_bytes(int mode, int spec_type, int warps, int hop, int n_fft, const PlanTables& t,
                        FeatParams* c) {
  const bool mel = (spec_type == EVF_SPEC_MEL || spec_type == EVF_SPEC_MEL_LIBROSA);
  const int fpj = (mode == MODE_HALF) ? 1 : 2;  // frames per job
  const int jobs = mode_jobs_per_warp(mode);
  const int fr = warps * mode_frames_per_warp(mode);
  const long long limit = (228 * 1024) / kCtasPerSm - 1024;  // kCtasPerSm resident CTAs, 1 KB reserved for each
  auto up4 = [](int w) { return (w + 3) & ~3; };
  c->n_chunk = t.n_chunk;
  c->n_heads = t.n_heads;
  c->m_pad = t.m_pad;
  c->n_slots = t.n_slots;
  // the P column must fit into the transpose scratch it aliases (one job per warp); the padded P columns of several
  // jobs extend the scratch instead
  int scr_words = 32 * kScrStride;
  if (mel) {
    const int need = jobs * t.n_chunk * 32 * fpj;
    if (jobs == 1 && need > scr_words) return -1;
    if (need > scr_words) scr_words = up4(need);
  }
  c->scr_words = scr_words;
  // Prefer a ring of two input buffers (the refill of one overlaps the FFTs on the other); fall back
  // to one when the hop is so large that two do not fit.
  for (int nbuf = 2; nbuf >= 1; --nbuf) {
    int w = 0;
    c->nbuf = nbuf;
    c->off_bar = w;
    w += 8;  // two 8-byte mbarriers + two release counters (+ pad)
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
    if (mode == MODE_HALF) w += up4(2 * 513);
    c->off_wtab = c->off_gtab = c->off_ltab = w;
    if (mel) {
      w += 2 * 32 * t.n_chunk;
      c->off_gtab = w;
      w += up4(2 * (t.n_heads + 1) * t.m_pad);
      c->off_ltab = w;
      w += 32;
    }
    c->off_warp = w;
    c->warp_words = scr_words + (mel ? up4(t.n_slots * 2 * fpj) : 0);
    w += warps * c->warp_words;
    const long long bytes = 4ll * w;
    if (bytes <= limit) return (int)bytes;
  }
  return -1;
}

int features_configure(int mode, int spec_type, int sample_format, int smem_bytes) {
  FeatParams dummy{};
  return dispatch(mode, spec_type, sample_format, dummy, 1, smem_bytes, nullptr, true);
}

int features_launch(int mode, int spec_type, int sample_format, const FeatParams& p, int grid, int smem_bytes,
                    cudaStream_t stream) {
  return dispatch(mode, spec_type, sample_format, p, grid, smem_bytes, stream, false);
}

}  // namespace evf
