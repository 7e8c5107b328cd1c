// In-register radix-2 DIF DFT-32 used by the warp-level 1024-point complex FFT.
//
// One warp computes one 1024-point complex FFT as 32 x 32 (four-step): every lane holds 32
// complex values in registers, does a 32-point DFT on them, the warp transposes through
// shared memory (with the inter-pass twiddle applied), and every lane does a second
// 32-point DFT.  All loops here are fully unrolled with compile-time indices so the
// arrays live in registers and the DFT-32 twiddles fold into immediates.
#pragma once
#include <cuda_runtime.h>

namespace evf {

__host__ __device__ constexpr int bitrev5(int x) {
  return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}

// cos(2*pi*k/32), sin(2*pi*k/32) for k = 0..8, correctly rounded from fp64.
__device__ __forceinline__ constexpr float cos32(int k) {
  switch (k) {
    case 0: return 1.0f;
    case 1: return 0.98078528040323043f;
    case 2: return 0.92387953251128674f;
    case 3: return 0.83146961230254524f;
    case 4: return 0.70710678118654757f;
    case 5: return 0.55557023301960218f;
    case 6: return 0.38268343236508978f;
    case 7: return 0.19509032201612825f;
    default: return 0.0f;
  }
}
__device__ __forceinline__ constexpr float sin32(int k) { return cos32(8 - k); }

// (r, i) *= exp(-2*pi*i*K/32) for K in [0, 16)
template <int K>
__device__ __forceinline__ void mul_w32(float& r, float& i) {
  if constexpr (K == 0) {
    return;
  } else if constexpr (K == 8) {  // -i
    float t = r;
    r = i;
    i = -t;
  } else if constexpr (K == 4) {  // (1 - i)/sqrt2
    float a = (r + i) * 0.70710678118654757f;
    float b = (i - r) * 0.70710678118654757f;
    r = a;
    i = b;
  } else if constexpr (K == 12) {  // (-1 - i)/sqrt2
    float a = (i - r) * 0.70710678118654757f;
    float b = -(r + i) * 0.70710678118654757f;
    r = a;
    i = b;
  } else if constexpr (K < 8) {
    constexpr float c = cos32(K), s = sin32(K);  // w = c - i s
    float a = fmaf(i, s, r * c);
    float b = fmaf(-r, s, i * c);
    r = a;
    i = b;
  } else {  // 8 < K < 16: w = -sin32(K-8) - i cos32(K-8)
    constexpr float c = cos32(K - 8), s = sin32(K - 8);
    float a = fmaf(i, c, -r * s);
    float b = fmaf(-r, c, -i * s);
    r = a;
    i = b;
  }
}

template <int HALF, int G, int J>
__device__ __forceinline__ void bfly(float (&re)[32], float (&im)[32]) {
  constexpr int a = G + J, b = G + J + HALF;
  float sr = re[a] + re[b], si = im[a] + im[b];
  float dr = re[a] - re[b], di = im[a] - im[b];
  mul_w32<J*(16 / HALF)>(dr, di);
  re[a] = sr;
  im[a] = si;
  re[b] = dr;
  im[b] = di;
}

template <int HALF, int G, int J>
struct BflyLoop {
  __device__ __forceinline__ static void run(float (&re)[32], float (&im)[32]) {
    bfly<HALF, G, J>(re, im);
    if constexpr (J + 1 < HALF) BflyLoop<HALF, G, J + 1>::run(re, im);
  }
};
template <int HALF, int G>
struct GroupLoop {
  __device__ __forceinline__ static void run(float (&re)[32], float (&im)[32]) {
    BflyLoop<HALF, G, 0>::run(re, im);
    if constexpr (G + 2 * HALF < 32) GroupLoop<HALF, G + 2 * HALF>::run(re, im);
  }
};

// Forward DFT-32, decimation in frequency, in place.  X[k] ends up at index bitrev5(k).
__device__ __forceinline__ void dft32_dif(float (&re)[32], float (&im)[32]) {
  GroupLoop<16, 0>::run(re, im);
  GroupLoop<8, 0>::run(re, im);
  GroupLoop<4, 0>::run(re, im);
  GroupLoop<2, 0>::run(re, im);
  GroupLoop<1, 0>::run(re, im);
}

}  // namespace evf
