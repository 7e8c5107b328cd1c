// C ABI of libevfeat.so (see include/evfeat.h): plan / batch management and entry points.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "evfeat_fft.cuh"
#include "evfeat_internal.h"

struct evf_plan {
  evf_config cfg;
  int device = 0;
  int mode = 0;
  int n_freq = 0;
  int row_floats = 0;
  int frames_per_tile = 0;
  int warps = 16;
  int num_sms = 0;
  int smem_bytes = 0;
  int k_used = 0;
  evf::FeatParams carve{};
  // device tables
  float* d_window = nullptr;
  float4* d_tw4 = nullptr;
  float2* d_wpost = nullptr;
  float2* d_wtab = nullptr;
  uint2* d_gtab = nullptr;
  unsigned* d_ltab = nullptr;
  float2* d_melw = nullptr;  // per-bin {rising, falling} weights and interval index: the backward's transposed mel
  int* d_jk = nullptr;
  // MODE_GENERIC (evfeat_generic.cu); the n_fft 512 / 256 warp modes keep these too: their backward runs there
  evf::GenParams gen{};
  bool has_gen = false;
  int gen_grid_per_sm = 1;
  int gen_smem_bytes = 0;
  int gen_frames_per_tile = 0;
  float* d_gen_window = nullptr;  // natural order, pre-scaled by 0.5
  float2* d_tw = nullptr;
  double2* d_tw64 = nullptr;
  int* d_kstart = nullptr;
  float* d_fb_dense = nullptr;
  // MODE_DECIMATED (evfeat_decimated.cu): n_fft = dec_r * 1024 as dec_r phase-stream transforms + a combine
  int dec_r = 0;
  int dec_hop = 0;                // hop / dec_r: the hop of the phase streams
  int dec_smem_bytes = 0;
  evf::FeatParams dec_carve{};    // shared-memory carve-up of features_kernel<MODE_PACK2, raw> at dec_hop
  float* d_dec_windows = nullptr; // [dec_r][1024] pair layout of w_r[m] = w[dec_r * m + r] / 2
  float2* d_dec_wcomb = nullptr;  // [513] W_N^k
};

struct evf_batch {
  int device = 0;
  int n_utts = 0;
  int n_tiles = 0;
  int64_t total_frames = 0;
  int64_t max_len = 0;             // longest utterance in samples
  std::vector<int64_t> frame_off;  // host copy
  std::vector<int> tile_start;     // [n_utts + 1] first tile of every utterance
  long long* d_sample_off = nullptr;
  long long* d_frame_off = nullptr;
  evf::TileDesc* d_tiles = nullptr;
  // tiles of the any-size kernels when the plan's forward uses other tiles (n_fft 512 / 256: backward only)
  int n_gen_tiles = 0;
  evf::TileDesc* d_gen_tiles = nullptr;
  // MODE_DECIMATED: offsets of the utterances' phase streams inside one plane (padded to 4 words), host and device
  std::vector<long long> dec_stream_off;
  long long* d_dec_stream_off = nullptr;
};

namespace evf {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  g_last_error = buf;
  return (e == cudaErrorMemoryAllocation) ? EVF_ERR_OUT_OF_MEMORY : EVF_ERR_CUDA;
}

namespace {

template <typename T>
int upload(const std::vector<T>& h, T** d) {
  *d = nullptr;
  if (h.empty()) return EVF_OK;
  EVF_CUDA(cudaMalloc(reinterpret_cast<void**>(d), h.size() * sizeof(T)));
  EVF_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return EVF_OK;
}

// Compress a dense [n_freq][n_mels] triangular filterbank: every frequency bin lies in
// exactly one interval between adjacent filter centres, so it feeds at most two adjacent
// filters.  j(k) = interval index (monotone in k); a[k] = fb[k][j], b[k] = fb[k][j-1].
int compress_filterbank(const float* fb, int n_freq, int n_mels, PlanTables* t) {
  std::vector<int> jk(n_freq, 0);
  std::vector<float2> w(n_freq, make_float2(0.f, 0.f));
  int j_prev = 0, k_used = 0;
  for (int k = 0; k < n_freq; ++k) {
    const float* row = fb + (size_t)k * n_mels;
    int nz[3], cnt = 0;
    for (int m = 0; m < n_mels; ++m) {
      if (row[m] != 0.0f) {
        if (!(row[m] == row[m]) || std::isinf(row[m]) || row[m] < 0.0f) {
          set_error("mel filterbank contains NaN/Inf or a negative weight");
          return EVF_ERR_FILTERBANK;
        }
        if (cnt < 3) nz[cnt] = m;
        ++cnt;
      }
    }
    int j = j_prev;
    float a = 0.f, b = 0.f;
    if (cnt == 1) {
      const int m = nz[0];
      if (j_prev <= m) { j = m; a = row[m]; }
      else if (j_prev == m + 1) { j = m + 1; b = row[m]; }
      else cnt = 99;
    } else if (cnt == 2) {
      if (nz[1] == nz[0] + 1 && nz[1] >= j_prev) { j = nz[1]; b = row[nz[0]]; a = row[nz[1]]; }
      else cnt = 99;
    }
    if (cnt > 2) {
      char buf[256];
      snprintf(buf, sizeof(buf),
               "mel filterbank row %d is not covered by two adjacent, monotonically ordered "
               "triangular filters; only torchaudio/librosa style banks are supported", k);
      set_error(buf);
      return EVF_ERR_FILTERBANK;
    }
    jk[k] = j;
    w[k] = make_float2(a, b);
    if (cnt > 0) k_used = k + 1;
    j_prev = j;
  }
  if (k_used == 0) k_used = 1;  // all-zero bank: keep one (zero-weight) bin so tables are non-empty
  t->k_used = k_used;
  t->melw.assign(w.begin(), w.begin() + k_used);
  t->kstart.assign(n_mels + 2, k_used);
  {
    int k = 0;
    for (int j = 0; j <= n_mels; ++j) {
      while (k < k_used && jk[k] < j) ++k;
      t->kstart[j] = k;
    }
    t->kstart[n_mels + 1] = k_used;
  }
  t->jk.assign(jk.begin(), jk.begin() + k_used);
  t->jk.push_back(-1);  // sentinel: the last bin always ends its interval
  return EVF_OK;
}

// Tables of the per-warp mel walk (evfeat_features.cu, "mel").  Interval j (between filter centres
// j - 1 and j) owns bins [kstart[j], kstart[j + 1]); its rising sums feed filter j, its falling
// sums filter j - 1.  Lane l walks bins [n * l, n * l + n) with n odd.
//   slots    : one per non-empty interval (dense ordinal q[j]), then 32 head slots (one per lane,
//              for the partial of an interval that an earlier lane started), then one zero slot
//   wtab     : per (i, lane) the two weights of bin n * lane + i; the sign bit of the rising weight
//              says "flush after this bin" (last bin of an interval, of the chunk, or of all bins)
//   ltab     : per lane the slot of its first flush and of its second (later ones are consecutive)
//   gtab     : per filter m and c = 0 .. n_heads, ready to use as byte offsets into the warp's slot area (no decoding
//              in the gather): .x = c-th partial of interval m (its rising sums, the slot's first half), .y = c-th
//              partial of interval m + 1 (its falling sums, the slot's second half); the zero slot where there is none
int build_walk_tables(int n_mels, int slot_bytes, PlanTables* t) {
  const int k_used = t->k_used;
  int n = (k_used + 31) / 32;
  if (n < 1) n = 1;
  if ((n & 1) == 0) ++n;
  t->n_chunk = n;
  const int n_int = n_mels + 1;
  std::vector<int> q(n_int, -1);
  int Q = 0;
  for (int j = 0; j < n_int; ++j)
    if (t->kstart[j] < t->kstart[j + 1]) q[j] = Q++;
  const int head0 = Q, zero = Q + 32;
  t->n_slots = Q + 33;
  if (t->n_slots > 0xffff) {
    set_error("mel filterbank has too many filters");
    return EVF_ERR_FILTERBANK;
  }
  t->wtab.assign((size_t)n * 32, make_float2(0.f, 0.f));
  for (int l = 0; l < 32; ++l)
    for (int i = 0; i < n; ++i) {
      const int k = n * l + i;
      if (k >= k_used) continue;
      float2 w = t->melw[k];
      const bool flush = (i == n - 1) || (k == k_used - 1) || (t->jk[k + 1] != t->jk[k]);
      if (flush) w.x = -w.x;  // weights are >= 0 (checked above); -0.0f carries the flag for a zero weight
      t->wtab[(size_t)i * 32 + l] = w;
    }
  t->ltab.assign(32, (unsigned)zero | ((unsigned)zero << 16));
  std::vector<std::vector<int>> parts(n_int);  // slots that sum to interval j, in lane order
  for (int j = 0; j < n_int; ++j)
    if (q[j] >= 0) parts[j].push_back(q[j]);
  for (int l = 0; l < 32; ++l) {
    const int k0 = n * l;
    if (k0 >= k_used) break;
    const int j0 = t->jk[k0];
    const bool cont = k0 > t->kstart[j0];
    const int d0 = cont ? head0 + l : q[j0];
    t->ltab[l] = (unsigned)d0 | ((unsigned)(q[j0] + 1) << 16);
    if (cont) parts[j0].push_back(head0 + l);
  }
  size_t most = 1;
  for (int j = 0; j < n_int; ++j) most = parts[j].size() > most ? parts[j].size() : most;
  t->n_heads = (int)most - 1;
  t->m_pad = (n_mels + 31) & ~31;
  t->gtab.assign((size_t)(t->n_heads + 1) * t->m_pad, make_uint2((unsigned)(zero * slot_bytes), (unsigned)(zero * slot_bytes + slot_bytes / 2)));
  for (int m = 0; m < n_mels; ++m)
    for (int c = 0; c <= t->n_heads; ++c) {
      const unsigned r = c < (int)parts[m].size() ? parts[m][c] : zero;
      const unsigned f = c < (int)parts[m + 1].size() ? parts[m + 1][c] : zero;
      t->gtab[(size_t)c * t->m_pad + m] = make_uint2(r * slot_bytes, f * slot_bytes + slot_bytes / 2);
    }
  return EVF_OK;
}

void free_plan_tables(evf_plan* p) {
  cudaFree(p->d_window);
  cudaFree(p->d_tw4);
  cudaFree(p->d_wpost);
  cudaFree(p->d_wtab);
  cudaFree(p->d_gtab);
  cudaFree(p->d_ltab);
  cudaFree(p->d_melw);
  cudaFree(p->d_jk);
  cudaFree(p->d_tw);
  cudaFree(p->d_tw64);
  cudaFree(p->d_kstart);
  cudaFree(p->d_fb_dense);
  cudaFree(p->d_gen_window);
  cudaFree(p->d_dec_windows);
  cudaFree(p->d_dec_wcomb);
}

// frames of an utterance of L samples: torch.stft(center=True) yields 1 + (L + 2 * (n_fft / 2) - n_fft) / hop
// (= 1 + L / hop for even n_fft, 1 + (L - 1) / hop for odd n_fft); process_spec keeps the first L / hop of them
long long frames_of(const evf_config& c, long long L) {
  if (c.keep_last_frame) return 1 + (L + 2 * (long long)(c.n_fft / 2) - c.n_fft) / c.hop_length;
  return L / c.hop_length;
}

// Tables and launch geometry of the any-size kernel (evfeat_generic.cu).
int create_generic_plan(evf_plan* p, const float* window_host, const float* fb_dense_host, const PlanTables& t) {
  const evf_config& c = p->cfg;
  const int N = c.n_fft;
  const bool mel = (c.spec_type == EVF_SPEC_MEL || c.spec_type == EVF_SPEC_MEL_LIBROSA);
  GenParams& g = p->gen;
  bool needs64 = false;
  if (generic_factorize(N, &g.st, &needs64) != EVF_OK) {
    set_error("evf_plan_create: n_fft has more prime factors than the stage list holds");
    return EVF_ERR_UNSUPPORTED;
  }
  p->gen_smem_bytes = generic_smem_bytes(N, &g.pairs);
  if (p->gen_smem_bytes < 0) {
    set_error("evf_plan_create: n_fft is too large for the shared-memory FFT (two buffers of n_fft complex points "
              "must fit into 227 KB: n_fft <= 14 000)");
    return EVF_ERR_UNSUPPORTED;
  }
  p->gen_frames_per_tile = 2 * g.pairs;
  g.n_fft = N;
  g.hop = c.hop_length;
  g.n_freq = p->n_freq;
  g.n_mels = c.n_mels;
  g.row_floats = p->row_floats;
  g.apply_log = (c.spec_type == EVF_SPEC_RAW) ? 0 : c.apply_log;
  g.log_clip = c.log_clip;
  std::vector<float> win(N);
  for (int i = 0; i < N; ++i) win[i] = 0.5f * window_host[i];  // exact scaling: the pair separation omits its 1/2
  std::vector<float2> tw(N);
  std::vector<double2> tw64(needs64 ? N : 0);
  const double two_pi = 6.283185307179586476925286766559;
  for (int m = 0; m < N; ++m) {
    const double ang = -two_pi * (double)m / (double)N;
    tw[m] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    if (needs64) tw64[m] = make_double2(std::cos(ang), std::sin(ang));
  }
  int rc = generic_configure(c.spec_type, c.sample_format, p->gen_smem_bytes);
  if (rc == EVF_OK) rc = upload(win, &p->d_gen_window);
  if (rc == EVF_OK) rc = upload(tw, &p->d_tw);
  if (rc == EVF_OK) rc = upload(tw64, &p->d_tw64);
  if (rc == EVF_OK && mel) {
    if (fb_dense_host != nullptr) {
      std::vector<float> fb(fb_dense_host, fb_dense_host + (size_t)p->n_freq * c.n_mels);
      rc = upload(fb, &p->d_fb_dense);
    } else {
      if (!p->d_melw) rc = upload(t.melw, &p->d_melw);
      if (rc == EVF_OK) rc = upload(t.kstart, &p->d_kstart);
      std::vector<int> jk(t.jk.begin(), t.jk.begin() + t.k_used);
      if (rc == EVF_OK && !p->d_jk) rc = upload(jk, &p->d_jk);
    }
  }
  if (rc != EVF_OK) return rc;
  g.window = p->d_gen_window;
  g.tw = p->d_tw;
  g.tw64 = p->d_tw64;
  g.melw = p->d_melw;
  g.kstart = p->d_kstart;
  g.fb_dense = p->d_fb_dense;
  // resident CTAs per SM: shared memory (the opt-in 227 KB) and 2048 threads
  int per_sm = (227 * 1024) / (p->gen_smem_bytes + 1024);
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  p->gen_grid_per_sm = per_sm;
  p->has_gen = true;
  return EVF_OK;
}

}  // namespace
}  // namespace evf

using namespace evf;

extern "C" {

int evf_abi_version(void) { return EVF_ABI_VERSION; }

const char* evf_last_error(void) { return g_last_error.c_str(); }

int evf_plan_create(const evf_config* cfg, const float* window_host, const float* mel_fb_host,
                    int device, evf_plan** plan_out) {
  if (plan_out) *plan_out = nullptr;
  if (!cfg || !window_host || !plan_out) {
    set_error("evf_plan_create: null argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  const bool mel = (cfg->spec_type == EVF_SPEC_MEL || cfg->spec_type == EVF_SPEC_MEL_LIBROSA);
  if (cfg->spec_type < EVF_SPEC_MEL || cfg->spec_type > EVF_SPEC_RAW) {
    set_error("evf_plan_create: unknown spec_type");
    return EVF_ERR_UNSUPPORTED;
  }
  if (cfg->n_fft < 1 || cfg->hop_length < 1 || cfg->win_length < 1 || cfg->win_length > cfg->n_fft) {
    set_error("evf_plan_create: need n_fft >= 1, hop_length >= 1 and 1 <= win_length <= n_fft (torch.stft's own limits)");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (cfg->fft_path != EVF_FFT_AUTO && cfg->fft_path != EVF_FFT_GENERIC) {
    set_error("evf_plan_create: unknown fft_path");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (cfg->sample_format != EVF_SAMPLES_F32 && cfg->sample_format != EVF_SAMPLES_S16) {
    set_error("evf_plan_create: unknown sample_format");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (cfg->apply_log && !(cfg->log_clip >= 1.17549435e-38f)) {
    set_error("evf_plan_create: log_clip must be a normal positive float (the reference uses 1e-5)");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (mel && (!mel_fb_host || cfg->n_mels < 1)) {
    set_error("evf_plan_create: mel spec types need a filterbank and n_mels >= 1");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
    cudaGetLastError();
    set_error("evf_plan_create: no such CUDA device; libevfeat has no CPU path");
    return EVF_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  EVF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("evf_plan_create: device is not sm_100 (Blackwell B200); libevfeat is built for sm_100a only");
    return EVF_ERR_NO_DEVICE;
  }
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");

  evf_plan* p = new (std::nothrow) evf_plan();
  if (!p) return EVF_ERR_OUT_OF_MEMORY;
  p->cfg = *cfg;
  p->device = device;
  // the warp-per-FFT kernel covers n_fft 1024, 512 and 256 (any hop <= n_fft) and 2048 (even hops); everything else --
  // and whatever does not fit its shared-memory carve-up or its triangular-bank tables -- runs in the any-size kernel
  const bool fast_shape = cfg->fft_path == EVF_FFT_AUTO && cfg->hop_length <= cfg->n_fft &&
                          (cfg->n_fft == 1024 || cfg->n_fft == 512 || cfg->n_fft == 256 ||
                           (cfg->n_fft == 2048 && (cfg->hop_length & 1) == 0));
  p->mode = !fast_shape ? MODE_GENERIC
                        : (cfg->n_fft == 1024 ? MODE_PACK2
                                              : (cfg->n_fft == 512 ? MODE_PACK2_512
                                                                   : (cfg->n_fft == 256 ? MODE_PACK2_256 : MODE_HALF)));
  // n_fft = R * 1024 (R = 3, 4: the "output" transform of a configuration with a sampling-rate change) with a hop
  // divisible by R: R phase-stream transforms of 1024 points in the warp kernel + a combine (evfeat_decimated.cu)
  if (cfg->fft_path == EVF_FFT_AUTO && (cfg->n_fft == 3072 || cfg->n_fft == 4096) &&
      cfg->hop_length % (cfg->n_fft / 1024) == 0 && cfg->hop_length <= cfg->n_fft)
    p->mode = MODE_DECIMATED;
  p->n_freq = cfg->n_fft / 2 + 1;
  p->warps = kMaxWarps;  // kCtasPerSm CTAs per SM
  p->frames_per_tile = p->warps * mode_frames_per_warp(p->mode);  // n_fft 1024: 8 jobs of two frames
  p->num_sms = prop.multiProcessorCount;
  p->row_floats = mel ? cfg->n_mels : (cfg->spec_type == EVF_SPEC_RAW ? 2 * p->n_freq : p->n_freq);

  PlanTables t;
  bool triangular = true;
  int rc = EVF_OK;
  if (mel) {
    rc = compress_filterbank(mel_fb_host, p->n_freq, cfg->n_mels, &t);
    if (rc == EVF_OK && p->mode != MODE_GENERIC && p->mode != MODE_DECIMATED) rc = build_walk_tables(cfg->n_mels, p->mode == MODE_HALF ? 8 : 16, &t);
    if (rc != EVF_OK) {  // not a bank of adjacent triangular filters (or too many of them): dense projection
      triangular = false;
      p->mode = MODE_GENERIC;
      rc = EVF_OK;
    }
    p->k_used = t.k_used;
  }
  if (p->mode == MODE_DECIMATED) {
    p->dec_r = cfg->n_fft / kFftSize;
    p->dec_hop = cfg->hop_length / p->dec_r;
    PlanTables none;
    p->dec_smem_bytes = features_smem_bytes(MODE_PACK2, EVF_SPEC_RAW, p->warps, p->dec_hop, kFftSize, none, &p->dec_carve);
    if (p->dec_smem_bytes < 0) p->mode = MODE_GENERIC;
  } else if (p->mode != MODE_GENERIC) {
    // does the carve-up fit?  (large hops: two or even one input tile of (frames - 1) * hop + n_fft samples do not)
    FeatParams probe{};
    if (features_smem_bytes(p->mode, cfg->spec_type, p->warps, cfg->hop_length, cfg->n_fft, t, &probe) < 0)
      p->mode = MODE_GENERIC;
  }
  if (p->mode == MODE_GENERIC) {
    rc = create_generic_plan(p, window_host, triangular ? nullptr : mel_fb_host, t);
    if (rc != EVF_OK) {
      free_plan_tables(p);
      delete p;
      return rc;
    }
    p->smem_bytes = p->gen_smem_bytes;
    p->frames_per_tile = p->gen_frames_per_tile;
    *plan_out = p;
    return EVF_OK;
  }
  if (p->mode == MODE_DECIMATED) {
    // the any-size plan rides along: its kernels run the backward, and its tables (melw, kstart, jk) serve the combine
    rc = create_generic_plan(p, window_host, nullptr, t);
    const int R = p->dec_r;
    std::vector<float> wins((size_t)R * kFftSize);
    for (int r = 0; r < R; ++r)
      for (int q = 0; q < 16; ++q)
        for (int lane = 0; lane < 32; ++lane) {  // pair layout of the 1024-point kernel, window w_r[m] = w[R m + r] / 2
          wins[(size_t)r * kFftSize + (q * 32 + lane) * 2 + 0] = 0.5f * window_host[R * (32 * q + lane) + r];
          wins[(size_t)r * kFftSize + (q * 32 + lane) * 2 + 1] = 0.5f * window_host[R * (32 * (q + 16) + lane) + r];
        }
    std::vector<float2> wcomb(513);
    std::vector<float4> tw4(kFftSize / 2);
    const double two_pi_d = 6.283185307179586476925286766559;
    for (int k = 0; k <= 512; ++k) {
      const double ang = -two_pi_d * (double)k / (double)cfg->n_fft;
      wcomb[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    for (int n = 0; n < 16; ++n)
      for (int lane = 0; lane < 32; ++lane) {
        float c[2], sn[2];
        for (int h = 0; h < 2; ++h) {
          const double ang = -two_pi_d * (double)((lane * (n + 16 * h)) % kFftSize) / (double)kFftSize;
          c[h] = (float)std::cos(ang);
          sn[h] = (float)std::sin(ang);
        }
        tw4[n * 32 + lane] = make_float4(c[0], sn[0], c[1], sn[1]);
      }
    if (rc == EVF_OK) rc = features_configure(MODE_PACK2, EVF_SPEC_RAW, EVF_SAMPLES_F32, p->dec_smem_bytes);
    if (rc == EVF_OK) rc = upload(wins, &p->d_dec_windows);
    if (rc == EVF_OK) rc = upload(wcomb, &p->d_dec_wcomb);
    if (rc == EVF_OK) rc = upload(tw4, &p->d_tw4);
    if (rc == EVF_OK) {  // the scratch of a run is allocated stream-ordered: keep it in the pool between runs
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
    }
    if (rc != EVF_OK) {
      free_plan_tables(p);
      delete p;
      return rc;
    }
    p->frames_per_tile = 2 * p->warps;  // tiles of the phase streams: 8 jobs of two frames
    p->smem_bytes = p->dec_smem_bytes;
    *plan_out = p;
    return EVF_OK;
  }
  const int n_pts = cfg->n_fft < kFftSize ? cfg->n_fft : kFftSize;  // points of one packed job
  const int rows = n_pts / 32;                                       // its first-pass rows = lanes per job afterwards
  t.window.resize(cfg->n_fft);
  if (mode_is_pack2(p->mode)) {
    // pair layout for LDS.64: [r][lane] = {w[32 * r + lane], w[32 * (r + rows / 2) + lane]}, r < rows / 2: the two
    // inputs of one first-stage butterfly of the first pass (evfeat_fft.cuh, win_head); 0.5 is exact
    for (int r = 0; r < rows / 2; ++r)
      for (int lane = 0; lane < 32; ++lane) {
        t.window[(r * 32 + lane) * 2 + 0] = 0.5f * window_host[32 * r + lane];
        t.window[(r * 32 + lane) * 2 + 1] = 0.5f * window_host[32 * (r + rows / 2) + lane];
      }
  } else {
    for (int i = 0; i < cfg->n_fft; ++i) t.window[i] = 0.5f * window_host[i];  // exact scaling
  }
  // four-step twiddles W_N^(n2 * k1), N = points of a job, applied after the transpose (lane = (job, k1)) inside the
  // first butterfly stage of the second pass: entry [n][lane] = {t[n], t[n + 16]}, n < 16
  t.tw4.resize(kFftSize / 2);
  const double two_pi = 6.283185307179586476925286766559;
  for (int n = 0; n < 16; ++n) {
    for (int lane = 0; lane < 32; ++lane) {
      float c[2], sn[2];
      for (int h = 0; h < 2; ++h) {
        const int n2 = n + 16 * h;
        const double ang = -two_pi * (double)(((lane % rows) * n2) % n_pts) / (double)n_pts;
        c[h] = (float)std::cos(ang);
        sn[h] = (float)std::sin(ang);
      }
      t.tw4[n * 32 + lane] = make_float4(c[0], sn[0], c[1], sn[1]);
    }
  }
  if (p->mode == MODE_HALF) {
    t.wpost.resize(513);
    for (int k = 0; k <= 512; ++k) {
      const double ang = two_pi * (double)k / 2048.0;
      t.wpost[k] = make_float2((float)std::cos(ang), (float)(-std::sin(ang)));
    }
  }
  p->smem_bytes = features_smem_bytes(p->mode, cfg->spec_type, p->warps, cfg->hop_length, cfg->n_fft, t, &p->carve);
  if (p->smem_bytes < 0) {
    delete p;
    set_error("evf_plan_create: this n_fft / hop_length / n_mels combination needs more than 227 KB of shared memory per CTA");
    return EVF_ERR_UNSUPPORTED;
  }
  rc = features_configure(p->mode, cfg->spec_type, cfg->sample_format, p->smem_bytes);
  if (rc == EVF_OK) rc = upload(t.window, &p->d_window);
  if (rc == EVF_OK) rc = upload(t.tw4, &p->d_tw4);
  if (rc == EVF_OK) rc = upload(t.wpost, &p->d_wpost);
  if (rc == EVF_OK) rc = upload(t.wtab, &p->d_wtab);
  if (rc == EVF_OK) rc = upload(t.gtab, &p->d_gtab);
  if (rc == EVF_OK) rc = upload(t.ltab, &p->d_ltab);
  if (rc == EVF_OK && mel) {
    rc = upload(t.melw, &p->d_melw);
    std::vector<int> jk(t.jk.begin(), t.jk.begin() + t.k_used);
    if (rc == EVF_OK) rc = upload(jk, &p->d_jk);
  }
  // n_fft 256: the backward runs in the any-size kernels (own tables, own tiles: evf_batch_create); n_fft 512 has the
  // packed-job layout in the warp backward as well
  if (rc == EVF_OK && p->mode == MODE_PACK2_256) rc = create_generic_plan(p, window_host, nullptr, t);
  if (rc != EVF_OK) {
    free_plan_tables(p);
    delete p;
    return rc;
  }
  *plan_out = p;
  return EVF_OK;
}

int evf_plan_destroy(evf_plan* plan) {
  if (!plan) return EVF_OK;
  DeviceGuard guard(plan->device);
  free_plan_tables(plan);
  delete plan;
  return EVF_OK;
}

int evf_plan_row_floats(const evf_plan* plan, int32_t* row_floats_out) {
  if (!plan || !row_floats_out) {
    set_error("evf_plan_row_floats: null argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  *row_floats_out = plan->row_floats;
  return EVF_OK;
}

int64_t evf_plan_num_frames(const evf_plan* plan, int64_t n_samples) {
  if (!plan || n_samples < 0) return -1;
  return frames_of(plan->cfg, n_samples);
}

int evf_batch_create(const evf_plan* plan, const int64_t* sample_offsets_host, int32_t n_utts,
                     evf_batch** batch_out) {
  if (batch_out) *batch_out = nullptr;
  if (!plan || !batch_out || n_utts < 0 || (n_utts > 0 && !sample_offsets_host)) {
    set_error("evf_batch_create: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  const int hop = plan->cfg.hop_length, n_fft = plan->cfg.n_fft, fr = plan->frames_per_tile;
  std::vector<long long> s_off(n_utts + 1, 0), f_off(n_utts + 1, 0);
  std::vector<TileDesc> tiles, gen_tiles;
  std::vector<int> tile_start(n_utts + 1, 0);
  // frames a warp (any-size kernel: a team) loads together: the span of a tile covers whole groups
  const bool dec = plan->mode == MODE_DECIMATED;
  const int fpj = (plan->mode == MODE_GENERIC || dec) ? 2 : mode_frames_per_warp(plan->mode);
  const bool want_gen_tiles = plan->mode != MODE_GENERIC && plan->has_gen;
  std::vector<long long> stream_off(dec ? n_utts + 1 : 0, 0);
  for (int b = 0; b < n_utts; ++b) {
    tile_start[b] = (int)tiles.size();
    const int64_t L = sample_offsets_host[b + 1] - sample_offsets_host[b];
    if (L <= n_fft / 2) {
      char buf[256];
      snprintf(buf, sizeof(buf),
               "utterance %d has %lld samples; reflect padding needs more than n_fft/2 = %d "
               "(torch.stft raises for the same input)", b, (long long)L, n_fft / 2);
      set_error(buf);
      return EVF_ERR_SHORT_INPUT;
    }
    const int64_t T = frames_of(plan->cfg, L);
    if (T > 0x7fffffff - fr) {
      set_error("evf_batch_create: utterance too long");
      return EVF_ERR_INVALID_ARGUMENT;
    }
    if (L > 0x7fffffff - 4 * n_fft) {
      set_error("evf_batch_create: utterance too long");
      return EVF_ERR_INVALID_ARGUMENT;
    }
    f_off[b + 1] = f_off[b] + T;
    if (dec) {
      // tiles of the utterance's phase streams (the same list serves every stream: the planes have one layout):
      // 1024-point frames, hop / R apart, no padding (the streams are cut from the padded signal)
      const int hs = plan->dec_hop;
      const long long need = T > 0 ? (T - 1) * hs + kFftSize : 0;
      const long long Ls = (need + 3) & ~3ll;
      stream_off[b + 1] = stream_off[b] + Ls;
      for (int64_t f0 = 0; f0 < T; f0 += fr) {
        TileDesc d;
        d.s_off = stream_off[b];
        d.out_frame0 = f_off[b] + f0;
        d.L = (int)Ls;
        d.start = (int)(f0 * hs);
        d.nvalid = (int)((T - f0 < fr) ? (T - f0) : fr);
        const int njobs = (d.nvalid + 1) / 2;
        d.span = (njobs * 2 - 1) * hs + kFftSize;
        tiles.push_back(d);
      }
    } else
    for (int64_t f0 = 0; f0 < T; f0 += fr) {
      TileDesc d;
      d.s_off = sample_offsets_host[b];
      d.out_frame0 = f_off[b] + f0;
      d.L = (int)L;
      d.start = (int)(f0 * hop) - n_fft / 2;
      d.nvalid = (int)((T - f0 < fr) ? (T - f0) : fr);
      const int njobs = (d.nvalid + fpj - 1) / fpj;
      d.span = (njobs * fpj - 1) * hop + n_fft;
      tiles.push_back(d);
    }
    if (want_gen_tiles) {
      const int gfr = plan->gen_frames_per_tile;
      for (int64_t f0 = 0; f0 < T; f0 += gfr) {
        TileDesc d;
        d.s_off = sample_offsets_host[b];
        d.out_frame0 = f_off[b] + f0;
        d.L = (int)L;
        d.start = (int)(f0 * hop) - n_fft / 2;
        d.nvalid = (int)((T - f0 < gfr) ? (T - f0) : gfr);
        d.span = (((d.nvalid + 1) / 2) * 2 - 1) * hop + n_fft;
        gen_tiles.push_back(d);
      }
    }
  }
  tile_start[n_utts] = (int)tiles.size();
  for (int b = 0; b <= n_utts; ++b) s_off[b] = n_utts ? sample_offsets_host[b] : 0;

  DeviceGuard guard(plan->device);
  evf_batch* bt = new (std::nothrow) evf_batch();
  if (!bt) return EVF_ERR_OUT_OF_MEMORY;
  bt->device = plan->device;
  bt->n_utts = n_utts;
  bt->n_tiles = (int)tiles.size();
  bt->total_frames = f_off[n_utts];
  for (int b = 0; b < n_utts; ++b) {
    const int64_t L = sample_offsets_host[b + 1] - sample_offsets_host[b];
    bt->max_len = L > bt->max_len ? L : bt->max_len;
  }
  bt->frame_off.assign(f_off.begin(), f_off.end());
  bt->tile_start.swap(tile_start);
  // one device block, one copy: [sample offsets | frame offsets | tile descriptors]
  const size_t off_bytes = (size_t)(n_utts + 1) * sizeof(long long);
  const size_t tile_bytes = tiles.size() * sizeof(TileDesc);
  const size_t gen_tile_bytes = gen_tiles.size() * sizeof(TileDesc);
  std::vector<unsigned char> blob(2 * off_bytes + tile_bytes + gen_tile_bytes);
  memcpy(blob.data(), s_off.data(), off_bytes);
  memcpy(blob.data() + off_bytes, f_off.data(), off_bytes);
  if (tile_bytes) memcpy(blob.data() + 2 * off_bytes, tiles.data(), tile_bytes);
  if (gen_tile_bytes) memcpy(blob.data() + 2 * off_bytes + tile_bytes, gen_tiles.data(), gen_tile_bytes);
  bt->n_gen_tiles = (int)gen_tiles.size();
  unsigned char* d_blob = nullptr;
  int rc = upload(blob, &d_blob);
  if (rc != EVF_OK) {
    evf_batch_destroy(bt);
    return rc;
  }
  bt->d_sample_off = reinterpret_cast<long long*>(d_blob);  // owns the block
  bt->d_frame_off = reinterpret_cast<long long*>(d_blob + off_bytes);
  bt->d_tiles = reinterpret_cast<evf::TileDesc*>(d_blob + 2 * off_bytes);  // 16 * (n_utts + 1) bytes in: 16-byte aligned
  bt->d_gen_tiles = reinterpret_cast<evf::TileDesc*>(d_blob + 2 * off_bytes + tile_bytes);
  if (dec) {
    bt->dec_stream_off = stream_off;
    rc = upload(stream_off, &bt->d_dec_stream_off);
    if (rc != EVF_OK) {
      evf_batch_destroy(bt);
      return rc;
    }
  }
  *batch_out = bt;
  return EVF_OK;
}

int evf_batch_destroy(evf_batch* batch) {
  if (!batch) return EVF_OK;
  DeviceGuard guard(batch->device);
  cudaFree(batch->d_sample_off);  // the single block
  cudaFree(batch->d_dec_stream_off);
  delete batch;
  return EVF_OK;
}

int evf_batch_total_frames(const evf_batch* batch, int64_t* total_frames_out) {
  if (!batch || !total_frames_out) {
    set_error("evf_batch_total_frames: null argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  *total_frames_out = batch->total_frames;
  return EVF_OK;
}

int evf_batch_frame_offsets(const evf_batch* batch, int64_t* frame_offsets_host_out) {
  if (!batch || !frame_offsets_host_out) {
    set_error("evf_batch_frame_offsets: null argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  memcpy(frame_offsets_host_out, batch->frame_off.data(), batch->frame_off.size() * sizeof(int64_t));
  return EVF_OK;
}

int evf_batch_frame_offsets_dev(const evf_batch* batch, const int64_t** frame_offsets_dev_out) {
  if (!batch || !frame_offsets_dev_out) {
    set_error("evf_batch_frame_offsets_dev: null argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  *frame_offsets_dev_out = reinterpret_cast<const int64_t*>(batch->d_frame_off);
  return EVF_OK;
}

static int features_run_tiles(const evf_plan* plan, const evf_batch* batch, int tile_begin, int tile_end,
                              const void* samples_dev, float* spec_out_dev, float* energy_out_dev, void* stream) {
  if (tile_end <= tile_begin) return EVF_OK;
  DeviceGuard guard(plan->device);
  if (plan->mode == MODE_GENERIC) {
    GenParams g = plan->gen;
    g.samples = samples_dev;
    g.tiles = batch->d_tiles + tile_begin;
    g.n_tiles = tile_end - tile_begin;
    g.spec_out = spec_out_dev;
    g.energy_out = (plan->cfg.spec_type == EVF_SPEC_RAW) ? nullptr : energy_out_dev;
    const int cap = plan->num_sms * plan->gen_grid_per_sm;
    return generic_launch(plan->cfg.spec_type, plan->cfg.sample_format, g, g.n_tiles < cap ? g.n_tiles : cap,
                          plan->smem_bytes, static_cast<cudaStream_t>(stream));
  }
  FeatParams p = plan->carve;
  p.samples = samples_dev;
  p.tiles = batch->d_tiles + tile_begin;
  p.n_tiles = tile_end - tile_begin;
  p.spec_out = spec_out_dev;
  p.energy_out = (plan->cfg.spec_type == EVF_SPEC_RAW) ? nullptr : energy_out_dev;
  p.window = plan->d_window;
  p.tw4 = plan->d_tw4;
  p.wpost = plan->d_wpost;
  p.wtab = plan->d_wtab;
  p.gtab = plan->d_gtab;
  p.ltab = plan->d_ltab;
  p.hop = plan->cfg.hop_length;
  p.n_mels = plan->cfg.n_mels;
  p.n_freq = plan->n_freq;
  p.k_used = plan->k_used;
  p.row_floats = plan->row_floats;
  p.apply_log = (plan->cfg.spec_type == EVF_SPEC_RAW) ? 0 : plan->cfg.apply_log;
  p.log_clip = plan->cfg.log_clip;
  const int grid = p.n_tiles < plan->num_sms * kCtasPerSm ? p.n_tiles : plan->num_sms * kCtasPerSm;
  return features_launch(plan->mode, plan->cfg.spec_type, plan->cfg.sample_format, p, grid, plan->smem_bytes,
                         static_cast<cudaStream_t>(stream));
}

// MODE_DECIMATED: utterances [u0, u1) in chunks whose scratch (the phase streams and R half spectra per frame, about
// R * 4 KB per frame) stays below ~1 GiB; the scratch is allocated and freed in stream order.
static int features_run_decimated(const evf_plan* plan, const evf_batch* batch, int u0, int u1, const void* samples_dev,
                                  float* spec_out_dev, float* energy_out_dev, void* stream) {
  DeviceGuard guard(plan->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int R = plan->dec_r;
  const long long row_raw = 2 * (kFftSize / 2 + 1);  // 1026 floats: Y_r[0..512] as (re, im)
  const long long budget_frames = (1ll << 30) / (R * row_raw * 4);
  const std::vector<long long>& so = batch->dec_stream_off;
  const std::vector<int64_t>& fo = batch->frame_off;
  int c0 = u0;
  while (c0 < u1) {
    int c1 = c0 + 1;
    while (c1 < u1 && fo[c1 + 1] - fo[c0] <= budget_frames) ++c1;
    const long long n_f = fo[c1] - fo[c0];
    const long long plane_len = so[c1] - so[c0];
    if (n_f > 0) {
      long long max_ls = 0;
      for (int b = c0; b < c1; ++b) max_ls = so[b + 1] - so[b] > max_ls ? so[b + 1] - so[b] : max_ls;
      float* planes = nullptr;
      float* raw = nullptr;
      EVF_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&planes), (size_t)R * plane_len * sizeof(float), st));
      cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&raw), (size_t)R * n_f * row_raw * sizeof(float), st);
      if (e != cudaSuccess) {
        cudaFreeAsync(planes, st);
        return cuda_fail(e, "cudaMallocAsync (scratch of the decimated transform)");
      }
      int rc = decimated_deinterleave(samples_dev, plan->cfg.sample_format, batch->d_sample_off, batch->d_dec_stream_off,
                                      c0, c1 - c0, so[c0], plane_len, max_ls, R, plan->cfg.n_fft, planes, st);
      const int t0 = batch->tile_start[c0], t1 = batch->tile_start[c1];
      for (int r = 0; r < R && rc == EVF_OK; ++r) {
        FeatParams p = plan->dec_carve;
        p.samples = planes + (long long)r * plane_len - so[c0];   // the tiles carry absolute stream offsets
        p.tiles = batch->d_tiles + t0;
        p.n_tiles = t1 - t0;
        p.spec_out = raw + (long long)r * n_f * row_raw - fo[c0] * row_raw;  // ... and absolute frame indices
        p.energy_out = nullptr;
        p.window = plan->d_dec_windows + (size_t)r * kFftSize;
        p.tw4 = plan->d_tw4;
        p.hop = plan->dec_hop;
        p.n_freq = kFftSize / 2 + 1;
        p.row_floats = (int)row_raw;
        p.apply_log = 0;
        p.log_clip = 0.f;
        const int grid = p.n_tiles < plan->num_sms * kCtasPerSm ? p.n_tiles : plan->num_sms * kCtasPerSm;
        rc = features_launch(MODE_PACK2, EVF_SPEC_RAW, EVF_SAMPLES_F32, p, grid, plan->dec_smem_bytes, st);
      }
      if (rc == EVF_OK)
        rc = decimated_combine(R, plan->cfg.spec_type, raw, n_f * row_raw, n_f,
                               spec_out_dev + fo[c0] * (long long)plan->row_floats,
                               energy_out_dev ? energy_out_dev + fo[c0] : nullptr, plan->d_dec_wcomb, plan->d_melw,
                               plan->d_kstart, plan->cfg.n_mels, plan->n_freq, plan->k_used, plan->row_floats, plan->cfg.apply_log,
                               plan->cfg.log_clip, plan->num_sms, st);
      cudaFreeAsync(raw, st);
      cudaFreeAsync(planes, st);
      if (rc != EVF_OK) return rc;
    }
    c0 = c1;
  }
  return EVF_OK;
}

int evf_features_run(const evf_plan* plan, const evf_batch* batch, const void* samples_dev,
                     float* spec_out_dev, float* energy_out_dev, void* stream) {
  if (!plan || !batch) {
    set_error("evf_features_run: null plan or batch");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (batch->device != plan->device) {
    set_error("evf_features_run: plan and batch live on different devices");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (batch->n_tiles == 0) return EVF_OK;
  if (!samples_dev || !spec_out_dev) {
    set_error("evf_features_run: null sample or output pointer");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (plan->mode == MODE_DECIMATED)
    return features_run_decimated(plan, batch, 0, batch->n_utts, samples_dev, spec_out_dev, energy_out_dev, stream);
  return features_run_tiles(plan, batch, 0, batch->n_tiles, samples_dev, spec_out_dev, energy_out_dev, stream);
}

int evf_features_run_range(const evf_plan* plan, const evf_batch* batch, int32_t utt_begin, int32_t utt_end,
                           const void* samples_base_dev, float* spec_base_dev, float* energy_base_dev,
                           void* stream) {
  if (!plan || !batch) {
    set_error("evf_features_run_range: null plan or batch");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (batch->device != plan->device) {
    set_error("evf_features_run_range: plan and batch live on different devices");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (utt_begin < 0 || utt_end > batch->n_utts || utt_begin > utt_end) {
    set_error("evf_features_run_range: utterance range outside the batch");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (utt_begin == utt_end) return EVF_OK;
  if (!samples_base_dev || !spec_base_dev) {
    set_error("evf_features_run_range: null sample or output pointer");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (plan->mode == MODE_DECIMATED)
    return features_run_decimated(plan, batch, utt_begin, utt_end, samples_base_dev, spec_base_dev, energy_base_dev,
                                  stream);
  return features_run_tiles(plan, batch, batch->tile_start[utt_begin], batch->tile_start[utt_end],
                            samples_base_dev, spec_base_dev, energy_base_dev, stream);
}

int evf_features_ragged(const evf_plan* plan, const void* samples_dev,
                        const int64_t* sample_offsets_host, int32_t n_utts, float* spec_out_dev,
                        float* energy_out_dev, int64_t* frame_offsets_host_out, void* stream) {
  evf_batch* bt = nullptr;
  int rc = evf_batch_create(plan, sample_offsets_host, n_utts, &bt);
  if (rc != EVF_OK) return rc;
  if (frame_offsets_host_out) evf_batch_frame_offsets(bt, frame_offsets_host_out);
  rc = evf_features_run(plan, bt, samples_dev, spec_out_dev, energy_out_dev, stream);
  // The kernel reads the batch tables asynchronously: wait before freeing them.
  if (rc == EVF_OK) {
    DeviceGuard guard(plan->device);
    cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
  }
  evf_batch_destroy(bt);
  return rc;
}

int evf_energy_from_spec(const float* spec_dev, int64_t n_frames, int32_t row_floats,
                         float* energy_out_dev, void* stream) {
  if (n_frames < 0 || row_floats < 1 || (n_frames > 0 && (!spec_dev || !energy_out_dev))) {
    set_error("evf_energy_from_spec: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  return launch_energy_from_spec(spec_dev, n_frames, row_floats, energy_out_dev,
                                 static_cast<cudaStream_t>(stream));
}

int evf_log_compress(const float* in_dev, float* out_dev, int64_t n, float c, float clip_val,
                     void* stream) {
  if (n < 0 || (n > 0 && (!in_dev || !out_dev))) {
    set_error("evf_log_compress: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  return launch_log_compress(in_dev, out_dev, n, c, clip_val, static_cast<cudaStream_t>(stream));
}

int evf_pitch_fill_unvoiced(const double* pitch_dev, const int64_t* offsets_dev, int32_t n_utts, float* out_dev,
                            void* stream) {
  if (n_utts < 0 || (n_utts > 0 && (!pitch_dev || !offsets_dev || !out_dev))) {
    set_error("evf_pitch_fill_unvoiced: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  return launch_pitch_fill_unvoiced(pitch_dev, offsets_dev, n_utts, out_dev, static_cast<cudaStream_t>(stream));
}

int evf_audio_gate_mask(const float* lkfs_dev, float gate_lkfs, float* values_dev, const int64_t* value_offsets_dev,
                        int32_t n_utts, int32_t* keep_out_dev, void* stream) {
  if (n_utts < 0 || (n_utts > 0 && (!lkfs_dev || (values_dev && !value_offsets_dev)))) {
    set_error("evf_audio_gate_mask: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  return launch_gate_mask(lkfs_dev, gate_lkfs, values_dev, value_offsets_dev, n_utts, keep_out_dev,
                          static_cast<cudaStream_t>(stream));
}

int evf_log_compress_backward(const float* in_dev, const float* grad_out_dev, float* grad_in_dev, int64_t n,
                              float clip_val, void* stream) {
  if (n < 0 || (n > 0 && (!in_dev || !grad_out_dev || !grad_in_dev))) {
    set_error("evf_log_compress_backward: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  return launch_log_compress_backward(in_dev, grad_out_dev, grad_in_dev, n, clip_val,
                                      static_cast<cudaStream_t>(stream));
}

int64_t evf_features_backward_scratch_floats(const evf_plan* plan, const evf_batch* batch) {
  if (!plan || !batch) return -1;
  return batch->total_frames * (int64_t)plan->cfg.n_fft;
}

int evf_features_backward(const evf_plan* plan, const evf_batch* batch, const float* samples_dev,
                          const float* grad_spec_dev, float* scratch_dev, float* grad_samples_dev, void* stream) {
  if (plan && plan->cfg.apply_log) {
    set_error("evf_features_backward: needs a linear-domain plan (apply_log = 0; the log has its own backward, "
              "evf_log_compress_backward, or is folded in by evf_features_backward_ex)");
    return EVF_ERR_UNSUPPORTED;
  }
  return evf_features_backward_ex(plan, batch, samples_dev, grad_spec_dev, EVF_GRAD_FRAME_MAJOR, nullptr, scratch_dev,
                                  grad_samples_dev, stream);
}

int evf_features_backward_ex(const evf_plan* plan, const evf_batch* batch, const float* samples_dev,
                             const float* grad_spec_dev, int32_t grad_layout, const float* log_spec_dev,
                             float* scratch_dev, float* grad_samples_dev, void* stream) {
  if (!plan || !batch || batch->device != plan->device) {
    set_error("evf_features_backward: invalid plan / batch");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (plan->cfg.sample_format != EVF_SAMPLES_F32 || plan->cfg.spec_type == EVF_SPEC_RAW) {
    set_error("evf_features_backward: needs float32 samples and a real spec_type");
    return EVF_ERR_UNSUPPORTED;
  }
  if (grad_layout != EVF_GRAD_FRAME_MAJOR && grad_layout != EVF_GRAD_BIN_MAJOR) {
    set_error("evf_features_backward_ex: unknown grad_layout");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (plan->cfg.apply_log && !log_spec_dev) {
    set_error("evf_features_backward_ex: a plan with apply_log = 1 needs the forward's log output (log_spec_dev)");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  if (batch->n_tiles == 0) return EVF_OK;
  if (!samples_dev || !grad_spec_dev || !scratch_dev || !grad_samples_dev) {
    set_error("evf_features_backward: null pointer");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  GradSrc gs{};
  gs.grad = grad_spec_dev;
  gs.log_spec = log_spec_dev;
  gs.log_clip = plan->cfg.log_clip;
  gs.bin_major = (grad_layout == EVF_GRAD_BIN_MAJOR);
  gs.keep_last = plan->cfg.keep_last_frame;
  gs.row = plan->row_floats;
  DeviceGuard guard(plan->device);
  if (plan->mode == MODE_GENERIC || plan->has_gen) {  // any n_fft / hop / mel basis: the shared-memory mixed-radix kernels
    GenParams g = plan->gen;
    g.samples = samples_dev;
    g.tiles = (plan->mode == MODE_GENERIC) ? batch->d_tiles : batch->d_gen_tiles;
    g.n_tiles = (plan->mode == MODE_GENERIC) ? batch->n_tiles : batch->n_gen_tiles;
    g.apply_log = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int cap = plan->num_sms * plan->gen_grid_per_sm;
    int rc = generic_backward_launch(plan->cfg.spec_type, g, gs, scratch_dev, plan->d_jk, plan->k_used,
                                     g.n_tiles < cap ? g.n_tiles : cap, plan->gen_smem_bytes, st);
    if (rc != EVF_OK) return rc;
    return overlap_add_launch(scratch_dev, batch->d_sample_off, batch->d_frame_off, batch->n_utts, batch->max_len,
                              plan->cfg.n_fft, plan->cfg.hop_length, grad_samples_dev, st);
  }
  BwdParams p{};
  p.samples = samples_dev;
  p.tiles = batch->d_tiles;
  p.n_tiles = batch->n_tiles;
  p.gs = gs;
  p.frame_grad = scratch_dev;
  p.grad_samples = grad_samples_dev;
  p.window = plan->d_window;
  p.tw4 = plan->d_tw4;
  p.wpost = plan->d_wpost;
  p.n_fft = plan->cfg.n_fft;
  p.melw = plan->d_melw;
  p.jk = plan->d_jk;
  p.sample_off = batch->d_sample_off;
  p.frame_off = batch->d_frame_off;
  p.n_utts = batch->n_utts;
  p.max_len = batch->max_len;
  p.hop = plan->cfg.hop_length;
  p.n_mels = plan->cfg.n_mels;
  p.k_used = plan->k_used;
  p.row_floats = plan->row_floats;
  p.spec_type = plan->cfg.spec_type;
  return features_backward_launch(p, static_cast<cudaStream_t>(stream));
}

int evf_segment_mean(const float* values_dev, const int64_t* value_offsets_dev,
                     const int64_t* durations_dev, const int64_t* phone_offsets_dev,
                     int32_t n_utts, float* out_dev, void* stream) {
  if (n_utts < 0 || (n_utts > 0 && (!value_offsets_dev || !phone_offsets_dev))) {
    set_error("evf_segment_mean: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  return launch_segment_mean(values_dev, value_offsets_dev, durations_dev, phone_offsets_dev,
                             n_utts, out_dev, static_cast<cudaStream_t>(stream));
}

int evf_stats_partial(const float* values_dev, int64_t n, double* out5_dev, int32_t accumulate,
                      void* stream) {
  if (n < 0 || !out5_dev || (n > 0 && !values_dev)) {
    set_error("evf_stats_partial: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  return launch_stats_partial(values_dev, n, out5_dev, accumulate, static_cast<cudaStream_t>(stream));
}

int evf_normalize_inplace(float* values_dev, int64_t n, float mean, float std, void* stream) {
  if (n < 0 || (n > 0 && !values_dev)) {
    set_error("evf_normalize_inplace: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  return launch_normalize(values_dev, n, mean, std, static_cast<cudaStream_t>(stream));
}

int evf_normalize_by_stats(float* values_dev, int64_t n, const double* stats5_dev, void* stream) {
  if (n < 0 || !stats5_dev || (n > 0 && !values_dev)) {
    set_error("evf_normalize_by_stats: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  return launch_normalize_by_stats(values_dev, n, stats5_dev, static_cast<cudaStream_t>(stream));
}

int evf_stats_merge(const double* parts_dev, int32_t n_parts, int32_t stride_doubles, double* out5_dev,
                    void* stream) {
  if (n_parts < 1 || stride_doubles < 5 || !parts_dev || !out5_dev) {
    set_error("evf_stats_merge: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  return launch_stats_merge(parts_dev, n_parts, stride_doubles, out5_dev, static_cast<cudaStream_t>(stream));
}

int evf_normalize_by_gathered_stats(float* values_dev, int64_t n, const double* parts_dev, int32_t n_parts,
                                    int32_t stride_doubles, void* stream) {
  if (n < 0 || n_parts < 1 || stride_doubles < 5 || !parts_dev || (n > 0 && !values_dev)) {
    set_error("evf_normalize_by_gathered_stats: invalid argument");
    return EVF_ERR_INVALID_ARGUMENT;
  }
  return launch_normalize_by_gathered(values_dev, n, parts_dev, n_parts, stride_doubles,
                                      static_cast<cudaStream_t>(stream));
}

}  // extern "C"
