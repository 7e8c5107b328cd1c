// Packed two-FFTs-per-warp form of the fused feature kernel for n_fft == 1024 (sm_100a).
//
// Same operator contract as evfeat_features.cu (framing -> window -> FFT -> |X|^2 -> mel -> log ->
// energy; reference: everyvoice/utils/heavy.py:39-113, preprocessor/preprocessor.py:220-233,
// 302-309, 921-927), same host-built tile list, same ring of bulk-copied input buffers, same mel
// walk -- but every register of the FFT and of the epilogue holds the SAME element of TWO
// independent jobs (job w and job w + 8 of the 16 jobs of a tile) as an f32x2 pair, so each
// butterfly, window product, twiddle product, power and mel FMA is ONE packed instruction
// (FFMA2 / FADD2 / FMUL2) for both jobs.  The packed forms have the scalar lane throughput but take
// half the issue slots, and the scalar kernel is issue bound: 64.7 % issue
// utilisation against 36.9 % FMA-pipe utilisation.  Shared-memory instructions halve as well (one
// LDS.128 / STS.64 moves both jobs; window, twiddle and mel-weight loads are shared by the pair).
//
//   grid   : persistent, one CTA of 8 warps per SM (<= 255 registers per thread)
//   tile   : 16 jobs = 32 consecutive frames of one utterance; warp w owns jobs w ("A", frames
//            2w, 2w + 1) and w + 8 ("B", frames 2w + 16, 2w + 17); A and B read disjoint sample
//            rows, so a packed register is filled by two scalar loads into adjacent registers
//   output : bit-identical to the scalar kernel (each half of a packed instruction is the same
//            IEEE operation in the same order)
#include "evfeat_fft.cuh"
#include "evfeat_internal.h"

namespace evf {

namespace {

#include "evfeat_device.cuh"

using TileInfo = TileDesc;

struct alignas(16) Slot8 {
  ulonglong2 r;  // rising partial sums  {frame a (A, B), frame b (A, B)}
  ulonglong2 f;  // falling partial sums {frame a (A, B), frame b (A, B)}
};
static_assert(sizeof(Slot8) == 32, "slot layout");

__device__ __forceinline__ f32x2 shfl2(f32x2 a, int src_lane) {
  float lo, hi;
  v_unpack(a, lo, hi);
  lo = __shfl_sync(0xffffffffu, lo, src_lane);
  hi = __shfl_sync(0xffffffffu, hi, src_lane);
  return v_pack(lo, hi);
}

// Two 1024-point complex FFTs (packed) of the 2 x 32x32 values held by one warp; see warp_fft1024_tail.
// The transpose scratch holds 8-byte elements: rows of 34 elements keep the 16-byte row reads aligned and
// conflict-free (a quarter-warp reads 8 rows whose starts are 68 words = 4 banks apart).
__device__ __forceinline__ void warp_fft1024_tail_x2(f32x2 (&re)[32], f32x2 (&im)[32],
                                                     const float4* __restrict__ s_tw4,
                                                     unsigned long long* __restrict__ scr, int lane) {
  dft32_dit_tail(re, im);
  f32x2 tr[32], ti[32];
  {
    const ulonglong2* row = reinterpret_cast<const ulonglong2*>(scr + lane * kScrStride);
#pragma unroll
    for (int p = 0; p < 32; ++p) scr[p * kScrStride + lane] = re[p].v;
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const ulonglong2 v = row[m];
      tr[2 * m].v = v.x;
      tr[2 * m + 1].v = v.y;
    }
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 32; ++p) scr[p * kScrStride + lane] = im[p].v;
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const ulonglong2 v = row[m];
      ti[2 * m].v = v.x;
      ti[2 * m + 1].v = v.y;
    }
    __syncwarp();
  }
  {
    const float4 t = s_tw4[lane];
    tw_head<true>(re[0], im[0], re[1], im[1], tr[0], ti[0], t.x, t.y, tr[16], ti[16], t.z, t.w);
  }
#pragma unroll
  for (int n = 1; n < 16; ++n) {
    const float4 t = s_tw4[n * 32 + lane];
    const int i = bitrev5(n);
    tw_head<false>(re[i], im[i], re[i + 1], im[i + 1], tr[n], ti[n], t.x, t.y, tr[n + 16], ti[n + 16], t.z, t.w);
  }
  dft32_dit_tail(re, im);
}

template <int SPEC, typename SampleT>
__global__ void __launch_bounds__(kX2Warps * 32, 1) features_kernel_x2(const FeatParams p) {
  constexpr int WARPS = kX2Warps;
  constexpr int kThreads = WARPS * 32;
  constexpr int NFFT = 1024;
  constexpr bool kMel = (SPEC == EVF_SPEC_MEL || SPEC == EVF_SPEC_MEL_LIBROSA);

  extern __shared__ __align__(16) float smem[];
  float* s_win = smem + p.off_win;
  float4* s_tw4 = reinterpret_cast<float4*>(smem + p.off_tw);
  const float2* s_wtab = reinterpret_cast<const float2*>(smem + p.off_wtab);
  const unsigned* s_gtab = reinterpret_cast<const unsigned*>(smem + p.off_gtab);
  const unsigned* s_ltab = reinterpret_cast<const unsigned*>(smem + p.off_ltab);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  int* s_cnt = reinterpret_cast<int*>(smem + p.off_bar + 4);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  unsigned long long* scr = reinterpret_cast<unsigned long long*>(smem + p.off_warp + warp * p.warp_words);
  Slot8* slots = reinterpret_cast<Slot8*>(scr + 32 * kScrStride);
  const int nbuf = p.nbuf;

  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    s_cnt[0] = 0;
    s_cnt[1] = 0;
    fence_mbar_init();
  }
  for (int i = tid; i < NFFT; i += kThreads) s_win[i] = p.window[i];
  for (int i = tid; i < kFftSize / 2; i += kThreads) s_tw4[i] = p.tw4[i];
  if constexpr (kMel) {
    for (int i = tid; i < p.n_chunk * 32; i += kThreads) reinterpret_cast<float2*>(smem + p.off_wtab)[i] = p.wtab[i];
    for (int i = tid; i < (p.n_heads + 1) * p.m_pad; i += kThreads)
      reinterpret_cast<unsigned*>(smem + p.off_gtab)[i] = p.gtab[i];
    if (tid < 32) reinterpret_cast<unsigned*>(smem + p.off_ltab)[tid] = p.ltab[tid];
  }
  // never-flushed slots (empty intervals, the zero slot) stay zero; so do both input buffers until their
  // first fill, so that an idle half of a packed register never carries NaN bit patterns around
  for (int i = tid; i < WARPS * p.warp_words; i += kThreads) smem[p.off_warp + i] = 0.f;
  for (int i = tid; i < p.in_words * nbuf; i += kThreads) smem[p.off_in + i] = 0.f;

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

  int tile = blockIdx.x;
  TileInfo cur = tile_info(tile);
  TileInfo nxt = (tile + G < p.n_tiles) ? tile_info(tile + G) : cur;
  __syncthreads();
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
    const int fA = 2 * warp;       // tile-local frames of job A: fA, fA + 1
    const int fB = fA + 2 * WARPS;  // ... of job B
    const bool active = fA < nvalid;

    mbar_wait(&s_bar[b], parity);

    f32x2 re[32], im[32];
    if (active) {
      // ---- samples -> packed registers {job A, job B}; window fused into the first butterfly stage
      const SampleT* xa = s_in + fA * hop + lane;
      const SampleT* xb = s_in + fB * hop + lane;
      const float2* wv = reinterpret_cast<const float2*>(s_win) + lane;
      if (hop == 256) {
        // the two frames of a job overlap by 768 samples: 40 row loads feed both (rows 8..31 of frame a are
        // rows 0..23 of frame b)
        f32x2 v[40];
#pragma unroll
        for (int r = 0; r < 40; ++r) v[r] = v_pack(to_float(xa[32 * r]), to_float(xb[32 * r]));
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float2 w = wv[32 * r];
          const int i = bitrev5(r);
          win_head(re[i], re[i + 1], v[r], w.x, v[r + 16], w.y);
          win_head(im[i], im[i + 1], v[r + 8], w.x, v[r + 24], w.y);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float2 w = wv[32 * r];
          const int i = bitrev5(r);
          win_head(re[i], re[i + 1], v_pack(to_float(xa[32 * r]), to_float(xb[32 * r])), w.x,
                   v_pack(to_float(xa[32 * (r + 16)]), to_float(xb[32 * (r + 16)])), w.y);
          win_head(im[i], im[i + 1], v_pack(to_float(xa[hop + 32 * r]), to_float(xb[hop + 32 * r])), w.x,
                   v_pack(to_float(xa[hop + 32 * (r + 16)]), to_float(xb[hop + 32 * (r + 16)])), w.y);
        }
      }
    }

    // ---- release the input buffer; the last warp to do so refills it ---------------------
    {
      const int refill_tile = tile + nbuf * G;
      int last = 0;
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        last = (atomicAdd(&s_cnt[b], 1) == WARPS - 1);
        __threadfence_block();
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
      warp_fft1024_tail_x2(re, im, s_tw4, scr, lane);

      const int src_lane = (32 - lane) & 31;
      // validity of the four frames of this warp: A.a (always), A.b, B.a, B.b
      const bool vAb = fA + 1 < nvalid, vBa = fB < nvalid, vBb = fB + 1 < nvalid;
      float* gAa = p.spec_out + (out_frame0 + fA) * (long long)p.row_floats;
      float* gAb = gAa + p.row_floats;
      float* gBa = gAa + (long long)(2 * WARPS) * p.row_floats;
      float* gBb = gBa + p.row_floats;
      float eAa = 0.f, eAb = 0.f, eBa = 0.f, eBb = 0.f;
      const int kcap = kMel ? min(p.n_chunk * 32, NFFT / 2 + 1) : (NFFT / 2 + 1);
      ulonglong2* P4 = reinterpret_cast<ulonglong2*>(scr);  // per bin {frame a (A, B), frame b (A, B)}

#pragma unroll
      for (int j = 0; j <= 16; ++j) {
        const int k = lane + 32 * j;
        if (32 * j < kcap) {  // warp-uniform
          f32x2 zr, zi, pr, pi;
          if (j < 16) {
            zr = re[j];
            zi = im[j];
            f32x2 sr, si;
            sr.v = (lane == 0) ? re[(32 - j) & 31].v : re[31 - j].v;
            si.v = (lane == 0) ? im[(32 - j) & 31].v : im[31 - j].v;
            pr = shfl2(sr, src_lane);
            pi = shfl2(si, src_lane);
          } else {
            zr = pr = re[16];
            zi = pi = im[16];
          }
          if (j < 16 || lane == 0) {
            // window was pre-scaled by 1/2: X_a = Z[k] + conj(Z[N-k]), X_b = (Z[k] - conj(Z[N-k])) / i
            const f32x2 ar = v_add(zr, pr), ai = v_sub(zi, pi);
            const f32x2 br = v_add(zi, pi), bi = v_sub(pr, zr);
            if constexpr (SPEC == EVF_SPEC_RAW) {
              float arA, arB, aiA, aiB, brA, brB, biA, biB;
              v_unpack(ar, arA, arB);
              v_unpack(ai, aiA, aiB);
              v_unpack(br, brA, brB);
              v_unpack(bi, biA, biB);
              reinterpret_cast<float2*>(gAa)[k] = make_float2(arA, aiA);
              if (vAb) reinterpret_cast<float2*>(gAb)[k] = make_float2(brA, biA);
              if (vBa) reinterpret_cast<float2*>(gBa)[k] = make_float2(arB, aiB);
              if (vBb) reinterpret_cast<float2*>(gBb)[k] = make_float2(brB, biB);
            } else {
              f32x2 pa = v_fma2(ar, ar, v_mul2(ai, ai));
              f32x2 pb = v_fma2(br, br, v_mul2(bi, bi));
              if constexpr (SPEC == EVF_SPEC_MEL_LIBROSA) {
                float a0, a1, b0, b1;
                v_unpack(pa, a0, a1);
                v_unpack(pb, b0, b1);
                pa = v_pack(fast_sqrt(a0 + 1e-9f), fast_sqrt(a1 + 1e-9f));
                pb = v_pack(fast_sqrt(b0 + 1e-9f), fast_sqrt(b1 + 1e-9f));
              }
              if constexpr (kMel) {
                P4[k] = make_ulonglong2(pa.v, pb.v);
              } else {
                float a0, a1, b0, b1;
                v_unpack(pa, a0, a1);
                v_unpack(pb, b0, b1);
                const float vAa_ = compress(a0, p.apply_log, p.log_clip);
                gAa[k] = vAa_;
                eAa = fmaf(vAa_, vAa_, eAa);
                if (vAb) {
                  const float v = compress(b0, p.apply_log, p.log_clip);
                  gAb[k] = v;
                  eAb = fmaf(v, v, eAb);
                }
                if (vBa) {
                  const float v = compress(a1, p.apply_log, p.log_clip);
                  gBa[k] = v;
                  eBa = fmaf(v, v, eBa);
                }
                if (vBb) {
                  const float v = compress(b1, p.apply_log, p.log_clip);
                  gBb[k] = v;
                  eBb = fmaf(v, v, eBb);
                }
              }
            }
          }
        }
      }

      if constexpr (kMel) {
        // ---- mel step 1: lane = chunk of n bins; four packed running FMAs per bin ----------
        __syncwarp();
        const int n = p.n_chunk;
        {
          const float2* wp = s_wtab + lane;
          Slot8* dst = slots + (lt & 0xffffu);
          Slot8* dst_next = slots + (lt >> 16);
          const ulonglong2* pp = P4 + n * lane;
          f32x2 ra, rb, fa_, fb;
          ra.v = rb.v = fa_.v = fb.v = 0ull;
#pragma unroll 4
          for (int i = 0; i < n; ++i, ++pp, wp += 32) {
            const ulonglong2 pv = *pp;
            const float2 w = *wp;
            const float wr = fabsf(w.x);
            f32x2 pa, pb;
            pa.v = pv.x;
            pb.v = pv.y;
            ra = v_fma(pa, wr, ra);
            rb = v_fma(pb, wr, rb);
            fa_ = v_fma(pa, w.y, fa_);
            fb = v_fma(pb, w.y, fb);
            if (__float_as_int(w.x) < 0) {  // the interval (or the chunk) ends with this bin
              dst->r = make_ulonglong2(ra.v, rb.v);
              dst->f = make_ulonglong2(fa_.v, fb.v);
              dst = dst_next;
              ++dst_next;
              ra.v = rb.v = fa_.v = fb.v = 0ull;
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
        for (int m = lane; m < n_mels; m += 32) {
          const unsigned* gp = s_gtab + m;
          f32x2 va, vb;
          va.v = vb.v = 0ull;
          for (int c = 0; c <= nh; ++c, gp += m_pad) {
            const unsigned g = *gp;
            const ulonglong2 r = slots[g & 0xffffu].r;
            const ulonglong2 f = slots[g >> 16].f;
            f32x2 t;
            t.v = r.x;
            va = v_add(va, t);
            t.v = r.y;
            vb = v_add(vb, t);
            t.v = f.x;
            va = v_add(va, t);
            t.v = f.y;
            vb = v_add(vb, t);
          }
          float a0, a1, b0, b1;
          v_unpack(va, a0, a1);
          v_unpack(vb, b0, b1);
          a0 = compress(a0, apply_log, clip);
          gAa[m] = a0;
          eAa = fmaf(a0, a0, eAa);
          if (vAb) {
            b0 = compress(b0, apply_log, clip);
            gAb[m] = b0;
            eAb = fmaf(b0, b0, eAb);
          }
          if (vBa) {
            a1 = compress(a1, apply_log, clip);
            gBa[m] = a1;
            eBa = fmaf(a1, a1, eBa);
          }
          if (vBb) {
            b1 = compress(b1, apply_log, clip);
            gBb[m] = b1;
            eBb = fmaf(b1, b1, eBb);
          }
        }
      }

      if constexpr (SPEC != EVF_SPEC_RAW) {
        if (p.energy_out != nullptr) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            eAa += __shfl_xor_sync(0xffffffffu, eAa, o);
            eAb += __shfl_xor_sync(0xffffffffu, eAb, o);
            eBa += __shfl_xor_sync(0xffffffffu, eBa, o);
            eBb += __shfl_xor_sync(0xffffffffu, eBb, o);
          }
          if (lane == 0) {
            float* e = p.energy_out + out_frame0 + fA;
            e[0] = sqrtf(eAa);
            if (vAb) e[1] = sqrtf(eAb);
            if (vBa) e[2 * WARPS] = sqrtf(eBa);
            if (vBb) e[2 * WARPS + 1] = sqrtf(eBb);
          }
        }
      }
    }
    cur = nxt;
    nxt = fut;
  }
}

template <int SPEC, typename SampleT>
int launch_t(const FeatParams& p, int grid, int smem, cudaStream_t stream, bool configure_only) {
  auto kern = features_kernel_x2<SPEC, SampleT>;
  if (configure_only) {
    EVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return EVF_OK;
  }
  kern<<<grid, kX2Warps * 32, smem, stream>>>(p);
  EVF_CUDA(cudaGetLastError());
  return EVF_OK;
}

template <int SPEC>
int launch_s(int fmt, const FeatParams& p, int grid, int smem, cudaStream_t st, bool cfg) {
  if (fmt == EVF_SAMPLES_S16) return launch_t<SPEC, short>(p, grid, smem, st, cfg);
  return launch_t<SPEC, float>(p, grid, smem, st, cfg);
}

}  // namespace

int features_x2_dispatch(int spec, int fmt, const FeatParams& p, int grid, int smem, cudaStream_t st, bool cfg) {
  switch (spec) {
    case EVF_SPEC_MEL: return launch_s<EVF_SPEC_MEL>(fmt, p, grid, smem, st, cfg);
    case EVF_SPEC_MEL_LIBROSA: return launch_s<EVF_SPEC_MEL_LIBROSA>(fmt, p, grid, smem, st, cfg);
    case EVF_SPEC_LINEAR: return launch_s<EVF_SPEC_LINEAR>(fmt, p, grid, smem, st, cfg);
    case EVF_SPEC_RAW: return launch_s<EVF_SPEC_RAW>(fmt, p, grid, smem, st, cfg);
  }
  set_error("unknown spec_type");
  return EVF_ERR_UNSUPPORTED;
}

}  // namespace evf
