// Device helpers shared by the feature kernels (included inside `namespace evf { namespace {`).
//   ModeTraits      -- FFT packing per n_fft
//   compress        -- log(max(v, clip)), utils/heavy.py:39-40
//   warp_fft1024    -- one warp, one 1024-point complex FFT (32 x 32 four-step)
//   mbarrier / cp.async.bulk primitives (TMA 1-D staging)
#pragma once

template <int MODE>
struct ModeTraits;
template <>
struct ModeTraits<MODE_PACK2> {
  static constexpr int kNfft = 1024;
  static constexpr int kFramesPerJob = 2;
  static constexpr int kJobs = 1;        // packed jobs a warp runs side by side
  using SlotT = float4;  // partial sums {rising a, rising b, falling a, falling b} of the two frames of a job
};
template <>
struct ModeTraits<MODE_HALF> {
  static constexpr int kNfft = 2048;
  static constexpr int kFramesPerJob = 1;
  static constexpr int kJobs = 1;
  using SlotT = float2;  // {rising, falling}
};
// n_fft 512 / 256: the warp's 32 x 32 register file holds 2 / 4 packed jobs of 16 x 32 / 8 x 32 points: lane n2 keeps
// rows n1 < 32 / J of every job in register block j in the first pass, lane (j, k1) owns Z_j[k1 + (32 / J) * k2] after
// the second (warp_fft1024_tail<J>)
template <>
struct ModeTraits<MODE_PACK2_512> {
  static constexpr int kNfft = 512;
  static constexpr int kFramesPerJob = 2;
  static constexpr int kJobs = 2;
  using SlotT = float4;
};
template <>
struct ModeTraits<MODE_PACK2_256> {
  static constexpr int kNfft = 256;
  static constexpr int kFramesPerJob = 2;
  static constexpr int kJobs = 4;
  using SlotT = float4;
};

// Staged samples are kept in their input format (the bulk copy cannot convert); int16 PCM becomes
// s / 32768 when it is loaded into registers, exactly what torchaudio.load yields for a PCM16 wav.
__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(short v) { return (float)v * (1.0f / 32768.0f); }
__device__ __forceinline__ float2 load_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 load_pair(const short* p) {
  const short2 v = *reinterpret_cast<const short2*>(p);
  return make_float2(to_float(v.x), to_float(v.y));
}

// log(max(v, clip)) as one MUFU.LG2 + one FMUL.  The plan guarantees clip >= FLT_MIN whenever
// apply_log is set (evf_plan_create), so the clamped argument is a normal number and the
// denormal pre-scaling of __logf is dead weight; a NaN input still propagates (torch.clamp).
__device__ __forceinline__ float compress(float v, int apply_log, float clip) {
  if (apply_log) {
    v = (v < clip) ? clip : v;
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(v));
    v = l * 0.69314718055994530942f;
  }
  return v;
}

// 1024-point complex FFT of the 32x32 values held by one warp, after the caller has run the
// (window-fused) first butterfly stage of the first pass.
// In : lane n2; index i (even) / i + 1 hold the span-1 butterfly outputs of rows n1 = bitrev5(i), n1 + 16
//      of z[32*n1 + n2].
// Out: lane k1 holds Z[k1 + 32*k2] at index k2.
// Shared-memory instruction diet: the four-step twiddles come as 16 LDS.128 (two per load), the
// transposed reads as 2 x 16 LDS.64 (row stride 34 words keeps them 8-byte aligned and
// conflict-free: half-warp lanes hit banks 2*lane, 2*lane + 1).
// J = 2 / 4: J transforms of 1024 / J points side by side.  In: lane n2, register block j holds transform j's rows
// (32 / J of them) after its span-1 stage.  Out: lane j * (32 / J) + k1 holds Z_j[k1 + (32 / J) * k2] at index k2; the
// twiddle table is W_(1024 / J)^(n2 * k1) for the lane's k1.
template <int J = 1>
__device__ __forceinline__ void warp_fft1024_tail(float (&re)[32], float (&im)[32],
                                                  const float4* __restrict__ s_tw4,
                                                  float* __restrict__ scr, int lane) {
  dft32_dit_tail_jobs<J>(re, im);  // index k1 (+ block j) holds Y_j[k1][n2 = lane]
  float tr[32], ti[32];
  {
    const float2* row = reinterpret_cast<const float2*>(scr + lane * kScrStride);
#pragma unroll
    for (int p = 0; p < 32; ++p) scr[p * kScrStride + lane] = re[p];
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const float2 v = row[m];
      tr[2 * m] = v.x;
      tr[2 * m + 1] = v.y;
    }
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 32; ++p) scr[p * kScrStride + lane] = im[p];
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const float2 v = row[m];
      ti[2 * m] = v.x;
      ti[2 * m + 1] = v.y;
    }
    __syncwarp();
  }
  // lane = k1 now; element n2 needs W_1024^(n2*k1): table entry [n][lane] = {t[n], t[n + 16]}.
  // The multiplication rides in the first butterfly stage of the second pass.
  {
    const float4 t = s_tw4[lane];  // n = 0: t[0] = 1
    tw_head<true>(re[0], im[0], re[1], im[1], tr[0], ti[0], t.x, t.y, tr[16], ti[16], t.z, t.w);
  }
#pragma unroll
  for (int n = 1; n < 16; ++n) {
    const float4 t = s_tw4[n * 32 + lane];
    const int i = bitrev5(n);
    tw_head<false>(re[i], im[i], re[i + 1], im[i + 1], tr[n], ti[n], t.x, t.y, tr[n + 16], ti[n + 16], t.z, t.w);
  }
  dft32_dit_tail(re, im);  // index k2 holds Z[k1 + 32*k2]
}

// sqrt for the mel-librosa magnitude: one MUFU (relative error <= 2^-22, far inside the 1e-3
// log-domain budget) instead of the ~8-instruction correctly rounded sequence.
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// d loss / d spectrogram of one tile for the backward kernels: frame-major rows or the per-utterance [row][T_b] layout
// autograd hands back, with the log compression's backward folded in when the forward's log output is given.
struct GradView {
  const float* g;
  const float* y;
  long long fm_base, bm_base;  // element index of (tile frame 0, bin 0) frame-major / of the utterance bin-major
  int Tb, t0, row;
  float thr;                   // log(clip) exactly as the forward computes it: clamped outputs compare equal
  bool bin_major;
  __device__ __forceinline__ GradView(const GradSrc& s, const TileDesc& ti, int n_fft, int hop) {
    g = s.grad;
    y = s.log_spec;
    row = s.row;
    bin_major = s.bin_major != 0;
    const int half = n_fft / 2;         // torch.stft pads n_fft / 2 on both sides (odd n_fft included)
    t0 = (ti.start + half) / hop;       // utterance-relative index of the tile's first frame
    Tb = s.keep_last ? (int)(((long long)ti.L + 2 * half - n_fft) / hop) + 1 : (int)(ti.L / hop);
    fm_base = ti.out_frame0 * (long long)row;
    bm_base = (ti.out_frame0 - t0) * (long long)row;
    thr = compress(s.log_clip, 1, s.log_clip);
  }
  // element indices of (tile frame f, bin m) in the gradient / in the log output
  __device__ __forceinline__ long long g_index(int f, int m) const {
    return bin_major ? bm_base + (long long)m * Tb + (t0 + f) : fm_base + (long long)f * row + m;
  }
  __device__ __forceinline__ long long y_index(int f, int m) const { return fm_base + (long long)f * row + m; }
  // gradient w.r.t. the linear-domain value from the raw gradient and (fused log) the forward's log output
  __device__ __forceinline__ float chain(float v, float yy) const {
    return (yy > thr) ? v * __expf(-yy) : ((yy != yy) ? yy : 0.f);
  }
  // gradient w.r.t. the LINEAR-domain value of bin / filter m of tile frame f
  __device__ __forceinline__ float at(int f, int m) const {
    float v = bin_major ? __ldg(g + bm_base + (long long)m * Tb + (t0 + f)) : __ldg(g + fm_base + (long long)f * row + m);
    if (y != nullptr) {
      const float yy = __ldg(y + fm_base + (long long)f * row + m);
      v = (yy > thr) ? v * __expf(-yy) : ((yy != yy) ? yy : 0.f);
    }
    return v;
  }
};

// ---- mbarrier / bulk-copy (TMA 1-D) primitives ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// Orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy
// (bulk copy) accesses to the same locations.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// global -> shared bulk copy (SASS: UBLKCP); completion is signalled on the mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 4-byte asynchronous global -> shared copies (SASS: LDGSTS)
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ int reflect_index(int j, int L) {
  j = (j < 0) ? -j : j;
  j = (j >= L) ? 2 * (L - 1) - j : j;
  return (j < 0) ? 0 : j;  // only reachable for frames beyond the last valid one
}

// ---- tile descriptors and input staging shared by the feature kernels -----------------------------------------
// Tile descriptors are host-built (no dependent loads) and fetched two tiles ahead, as two 16-byte words.
__device__ __forceinline__ TileDesc load_tile_desc(const TileDesc* tiles, int tile) {
  const int4* q = reinterpret_cast<const int4*>(tiles + tile);
  const int4 a = __ldg(q), b = __ldg(q + 1);
  TileDesc ti;
  ti.s_off = ((long long)(unsigned)a.x) | ((long long)a.y << 32);
  ti.out_frame0 = ((long long)(unsigned)a.z) | ((long long)a.w << 32);
  ti.L = b.x;
  ti.start = b.y;
  ti.nvalid = b.z;
  ti.span = b.w;
  return ti;
}

// Stage one tile into `buf`.  NT threads (a warp in the steady state, the CTA in the prologue) store the words the
// bulk copy cannot take -- reflected margins, a 16-byte-misaligned span -- then a leader arms the mbarrier and issues
// one bulk copy for the aligned in-range part [a_lo, a_hi) (evf_stage_bulk).  The caller orders the manual stores before
// the leader's arrive (__syncwarp / __syncthreads).
template <typename SampleT>
__device__ __forceinline__ void evf_stage_manual(const SampleT* __restrict__ samples, const TileDesc& ti, SampleT* buf,
                                             int t, int nt, int& a_lo, int& a_hi) {
  constexpr int kAlign = 16 / (int)sizeof(SampleT);  // samples per 16 bytes (bulk-copy granularity)
  const int lo = max(0, -ti.start);                   // first tile word inside the utterance
  const int hi = min(ti.span, ti.L - ti.start);       // one past the last
  a_lo = lo;
  a_hi = lo;
  // tile word i <-> packed sample s_off + start + i ; both sides must be 16-byte aligned.  The packed buffer itself
  // may start anywhere (a caller's view such as wav[1:]): its misalignment counts like an offset of the utterance
  const long long base_mis = (long long)((reinterpret_cast<uintptr_t>(samples) & 15u) / sizeof(SampleT));
  if ((((base_mis + ti.s_off + ti.start + lo) | lo) & (kAlign - 1)) == 0 && hi > lo)
    a_hi = lo + ((hi - lo) & ~(kAlign - 1));
  const int total = a_lo + (ti.span - a_hi);
  const SampleT* src = samples + ti.s_off;
  for (int e = t; e < total; e += nt) {
    const int w = (e < a_lo) ? e : a_hi + (e - a_lo);
    buf[w] = __ldg(src + reflect_index(ti.start + w, ti.L));
  }
}
template <typename SampleT>
__device__ __forceinline__ void evf_stage_bulk(const SampleT* __restrict__ samples, const TileDesc& ti, SampleT* buf,
                                           uint64_t* bar, int a_lo, int a_hi) {
  const uint32_t bytes = (uint32_t)(a_hi - a_lo) * (uint32_t)sizeof(SampleT);
  fence_proxy_async();  // generic-proxy accesses to the buffer (reads of the previous tile) before the async write
  mbar_arrive_expect_tx(bar, bytes);
  if (bytes) bulk_g2s(buf + a_lo, samples + ti.s_off + ti.start + a_lo, bytes, bar);
}
