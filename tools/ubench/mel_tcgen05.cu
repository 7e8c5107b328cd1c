// Micro-benchmark: the mel projection (power spectrum -> mel filterbank, the one contraction on the path) on the
// 5th-generation tensor cores, in the best case for them -- operands already resident in shared memory in the UMMA
// canonical layout.  It answers "what would tcgen05.mma cost for this step" with a measured number.
//
//   problem  : configs[1] of BASELINE.json: 470 190 frames x 373 bins (f_max 8 kHz at 22.05 kHz / n_fft 1024) -> 80 mels
//   tile     : M = 64 frames (one CTA per SM; M = 128 would need 385 KB for the split power tile alone)
//   A        : power tile [64 x 376] as TF32 hi / lo parts (the 1e-3 log-domain tolerance needs a 3-pass operand split:
//              hi*hi + lo*hi + hi*lo), K-major, no swizzle: 2 x 96 KB of shared memory
//   B        : the BANDED filterbank: every 8-bin K-chunk touches at most 16 adjacent filters, so a chunk is one
//              tcgen05.mma.kind::tf32 of M = 64, N = 16, K = 8 accumulating into a 16-column window of the TMEM tile
//              (47 chunks x 3 passes = 141 MMAs per tile, plus one zero-initialising MMA over all 80 columns)
//   epilogue : 4 warps read the accumulators with tcgen05.ld, log(max(., 1e-5)), store [frame][80] rows
//   pipeline : the MMA thread runs one tile ahead of the epilogue warps (two TMEM buffers, full / empty mbarriers)
//
// Build : nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/ubench/mel_tcgen05.bin tools/ubench/mel_tcgen05.cu
// Run   : tools/ubench/mel_tcgen05.bin          (prints ms per 470 190 frames, a numerics check, and the share of the MMAs)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int kFrames = 470190;
constexpr int kBins = 376;        // 373 used bins padded to a multiple of 8
constexpr int kChunks = kBins / 8;
constexpr int kMels = 80;
constexpr int kM = 64;            // frames per tile
constexpr int kWin = 16;          // filters per chunk window (N of a chunk's MMA)
constexpr int kTmemCols = 512;    // two buffers of 256 columns (each: two accumulators of 128 columns for the split-accumulator variant)
constexpr int kABytes = kM * kBins * 4;              // one part (hi or lo) of the power tile
constexpr int kBChunkBytes = kWin * 8 * 4;           // 16 filters x 8 bins
constexpr int kBBytes = kChunks * kBChunkBytes;
constexpr int kZeroBBytes = kMels * 8 * 4;           // an all-zero [80 x 8] B operand for the initialising MMA
constexpr int kSmemBytes = 2 * kABytes + kBBytes + kZeroBBytes + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous; LBO = next 16-byte column, SBO = next 8 rows
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version of sm_100
  return d;                // layout type 0 = no swizzle, base offset 0
}
// c = f32 (1 << 4), a = b = tf32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct Params {
  const float* a_hi;      // [kM x kBins] canonical layout image (the same synthetic tile for every iteration)
  const float* a_lo;
  const float* b;         // [kChunks][16 x 8] canonical layout images (hi parts)
  const int* win0;        // first filter of each chunk's window
  float* out;             // [tiles * kM][kMels]
  int tiles;
  int mma_passes;         // 3 = operand split; 0 = no MMAs at all (epilogue only: the rest of the tile loop)
  int split_acc;          // 1 = even / odd chunks accumulate into two TMEM accumulators (shorter dependency chains), summed in the epilogue
};

__global__ void __launch_bounds__(160, 1) mel_tcgen05_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_ahi = smem;
  uint8_t* s_alo = smem + kABytes;
  uint8_t* s_b = smem + 2 * kABytes;
  uint8_t* s_zero = s_b + kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_zero + kZeroBBytes);  // full[2], empty[2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 4);
  int* s_win0 = reinterpret_cast<int*>(s_tmem + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < kABytes / 16; i += blockDim.x) {
    reinterpret_cast<float4*>(s_ahi)[i] = reinterpret_cast<const float4*>(p.a_hi)[i];
    reinterpret_cast<float4*>(s_alo)[i] = reinterpret_cast<const float4*>(p.a_lo)[i];
  }
  for (int i = tid; i < kBBytes / 16; i += blockDim.x) reinterpret_cast<float4*>(s_b)[i] = reinterpret_cast<const float4*>(p.b)[i];
  for (int i = tid; i < kZeroBBytes / 16; i += blockDim.x) reinterpret_cast<float4*>(s_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < kChunks; i += blockDim.x) s_win0[i] = p.win0[i];
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 128);
    mbar_init(&bars[3], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // generic-proxy writes of the operands must be visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t id_win = make_idesc(kM, kWin), id_all = make_idesc(kM, kMels);
      for (int t = 0; t < p.tiles; ++t) {
        const int buf = t & 1;
        if (t >= 2) {
          mbar_wait(&bars[2 + buf], ((t >> 1) - 1) & 1);  // the epilogue of tile t - 2 has drained this buffer
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t d0 = tmem + buf * 256;
        if (p.mma_passes > 0) {
          // zero-initialise all 80 columns: D = A_chunk0 * 0
          umma_tf32(d0, make_desc(smem_u32(s_ahi), kM * 16, 128), make_desc(smem_u32(s_zero), kMels * 16, 128), id_all, 0);
          if (p.split_acc) umma_tf32(d0 + 128, make_desc(smem_u32(s_ahi), kM * 16, 128), make_desc(smem_u32(s_zero), kMels * 16, 128), id_all, 0);
          for (int pass = 0; pass < p.mma_passes; ++pass) {
            const uint8_t* a = (pass == 1) ? s_alo : s_ahi;  // hi*hi, lo*hi, hi*lo (B's lo part: the hi image again, same cost)
            for (int c = 0; c < kChunks; ++c) {
              const uint64_t ad = make_desc(smem_u32(a + (size_t)(2 * c) * kM * 16), kM * 16, 128);
              const uint64_t bd = make_desc(smem_u32(s_b + (size_t)c * kBChunkBytes), kWin * 16, 128);
              umma_tf32(d0 + s_win0[c] + ((p.split_acc && (c & 1)) ? 128 : 0), ad, bd, id_win, 1);
            }
          }
        }
        umma_commit(&bars[buf]);  // arrives on full[buf] when every MMA above has completed
      }
    }
  } else {
    // epilogue warps: warp w owns TMEM lanes 32 w .. 32 w + 31; with M = 64 the accumulator rows sit in lanes 0..15 of
    // every 32-lane quadrant (row = 16 * quadrant + lane)
    for (int t = 0; t < p.tiles; ++t) {
      const int buf = t & 1;
      mbar_wait(&bars[buf], (t >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = tmem + buf * 256 + ((uint32_t)(32 * warp) << 16);
      float* row = p.out + ((size_t)blockIdx.x * p.tiles + t) * kM * kMels + (size_t)(16 * warp + (lane & 15)) * kMels;
#pragma unroll
      for (int c0 = 0; c0 < kMels; c0 += 16) {
        float v[16];
        tmem_ld16(ta + c0, v);
        if (p.split_acc) {
          float u[16];
          tmem_ld16(ta + 128 + c0, u);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += u[i];
        }
        if (lane < 16) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 o;
            o.x = __logf(fmaxf(v[i], 1e-5f));
            o.y = __logf(fmaxf(v[i + 1], 1e-5f));
            o.z = __logf(fmaxf(v[i + 2], 1e-5f));
            o.w = __logf(fmaxf(v[i + 3], 1e-5f));
            *reinterpret_cast<float4*>(row + c0 + i) = o;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&bars[2 + buf]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

static float tf32_hi(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;  // the tensor core reads the top 19 bits
  float y;
  memcpy(&y, &u, 4);
  return y;
}

int main() {
  // HTK mel filterbank of torchaudio.functional.melscale_fbanks(513, 0, 8000, 80, 22050): triangles over the bin frequencies
  const int n_freq = 513, sr = 22050;
  std::vector<double> fpts(kMels + 2);
  auto hz2mel = [](double f) { return 2595.0 * std::log10(1.0 + f / 700.0); };
  auto mel2hz = [](double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); };
  for (int i = 0; i < kMels + 2; ++i) fpts[i] = mel2hz(hz2mel(0.0) + (hz2mel(8000.0) - hz2mel(0.0)) * i / (kMels + 1));
  std::vector<float> W((size_t)kBins * kMels, 0.f);  // [bin][mel]
  for (int k = 0; k < kBins && k < n_freq; ++k) {
    const double f = (double)k * (sr / 2.0) / (n_freq - 1);
    for (int m = 0; m < kMels; ++m) {
      const double up = (f - fpts[m]) / (fpts[m + 1] - fpts[m]), down = (fpts[m + 2] - f) / (fpts[m + 2] - fpts[m + 1]);
      const double w = std::max(0.0, std::min(up, down));
      W[(size_t)k * kMels + m] = (float)w;
    }
  }
  // chunk windows
  std::vector<int> win0(kChunks);
  int worst = 0;
  for (int c = 0; c < kChunks; ++c) {
    int lo = kMels, hi = -1;
    for (int k = 8 * c; k < 8 * c + 8; ++k)
      for (int m = 0; m < kMels; ++m)
        if (W[(size_t)k * kMels + m] != 0.f) { lo = std::min(lo, m); hi = std::max(hi, m); }
    if (hi < 0) { lo = 0; hi = 0; }
    int w0 = std::min(lo & ~3, kMels - kWin);  // 4-column aligned window start
    if (hi - w0 >= kWin) { printf("chunk %d spans filters %d..%d: wider than the window\n", c, lo, hi); return 1; }
    worst = std::max(worst, hi - lo + 1);
    win0[c] = w0;
  }
  // synthetic power tile: P[f][k] = 1 + 0.001 * ((f * 7 + k * 3) % 11), split into TF32 hi / lo
  std::vector<float> P((size_t)kM * kBins), a_hi((size_t)kM * kBins), a_lo((size_t)kM * kBins);
  for (int f = 0; f < kM; ++f)
    for (int k = 0; k < kBins; ++k) P[(size_t)f * kBins + k] = (k < 373) ? 1.0f + 0.001f * ((f * 7 + k * 3) % 11) + 1e-4f * (k % 13) : 0.f;
  // canonical K-major no-swizzle image: 16-byte column q (4 bins), row r at (q * kM + r) * 16 bytes
  for (int f = 0; f < kM; ++f)
    for (int k = 0; k < kBins; ++k) {
      const size_t at = ((size_t)(k / 4) * kM + f) * 4 + (k % 4);
      const float h = tf32_hi(P[(size_t)f * kBins + k]);
      a_hi[at] = h;
      a_lo[at] = P[(size_t)f * kBins + k] - h;
    }
  std::vector<float> b((size_t)kChunks * kWin * 8);
  for (int c = 0; c < kChunks; ++c)
    for (int n = 0; n < kWin; ++n)
      for (int kk = 0; kk < 8; ++kk)
        b[(size_t)c * kWin * 8 + ((size_t)(kk / 4) * kWin + n) * 4 + (kk % 4)] = tf32_hi(W[(size_t)(8 * c + kk) * kMels + win0[c] + n]);

  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int tiles_total = (kFrames + kM - 1) / kM;
  const int tiles = (tiles_total + sms - 1) / sms;  // per CTA
  float *d_ahi, *d_alo, *d_b, *d_out;
  int* d_win0;
  CK(cudaMalloc(&d_ahi, a_hi.size() * 4));
  CK(cudaMalloc(&d_alo, a_lo.size() * 4));
  CK(cudaMalloc(&d_b, b.size() * 4));
  CK(cudaMalloc(&d_win0, win0.size() * 4));
  CK(cudaMalloc(&d_out, (size_t)sms * tiles * kM * kMels * 4));
  CK(cudaMemcpy(d_ahi, a_hi.data(), a_hi.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_alo, a_lo.data(), a_lo.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_win0, win0.data(), win0.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(mel_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  printf("banded mel projection on tcgen05: %d SMs x %d tiles of %d frames (%d frames), %d chunks x 3 passes of M=%d N=%d K=8 tf32, "
         "widest chunk touches %d filters, %d bytes of shared memory per CTA\n",
         sms, tiles, kM, sms * tiles * kM, kChunks, kM, kWin, worst, kSmemBytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int variant = 0; variant < 4; ++variant) {
    const int passes = (variant == 0 || variant == 1) ? 3 : (variant == 2 ? 1 : 0), split = (variant == 1);
    Params prm{d_ahi, d_alo, d_b, d_win0, d_out, tiles, passes, split};
    for (int i = 0; i < 3; ++i) mel_tcgen05_kernel<<<sms, 160, kSmemBytes>>>(prm);
    CK(cudaDeviceSynchronize());
    float best = 1e9f, sum = 0.f;
    const int reps = 20;
    for (int i = 0; i < reps; ++i) {
      cudaEventRecord(e0);
      mel_tcgen05_kernel<<<sms, 160, kSmemBytes>>>(prm);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, ms);
      sum += ms;
    }
    const double scale = (double)kFrames / ((double)sms * tiles * kM);
    printf("passes=%d split_acc=%d: %.4f ms mean, %.4f ms min per launch -> %.4f ms per 470 190 frames (%s)\n", passes, split, sum / reps, best,
           sum / reps * scale,
           passes == 3 ? "3-pass TF32 split: what parity needs" : passes == 1 ? "single TF32 pass: misses the tolerance" : "no MMAs: TMEM loads + log + stores only");
    if (passes == 3) {
      std::vector<float> out((size_t)kM * kMels);
      CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
      // reference of what the kernel computes: hi*hi + lo*hi + hi*hi(B's lo image is the hi image) = (2 hi + lo) * hi_w
      double worst_rel = 0.0;
      for (int f = 0; f < kM; ++f)
        for (int m = 0; m < kMels; ++m) {
          double acc = 0.0;
          for (int k = 0; k < kBins; ++k) {
            const float pv = P[(size_t)f * kBins + k], h = tf32_hi(pv);
            acc += (2.0 * h + (double)(pv - h)) * (double)tf32_hi(W[(size_t)k * kMels + m]);
          }
          const double want = std::log(std::max(acc, 1e-5)), got = out[(size_t)f * kMels + m];
          worst_rel = std::max(worst_rel, std::fabs(want - got));
        }
      printf("numerics of tile 0 (all 64 rows x 80 filters against a float64 sum of the same split products): max |d log| = %.3e\n", worst_rel);
    }
  }
  return 0;
}
