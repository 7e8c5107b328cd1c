// In-register radix-2 DIT DFT-32 used by the warp-level 1024-point complex FFT.
//
// One warp computes one 1024-point complex FFT as 32 x 32 (four-step): every lane holds 32
// complex values in registers, does a 32-point DFT on them, the warp transposes through
// shared memory, and every lane does a second 32-point DFT whose first stage carries the
// inter-pass twiddle.  All loops here are fully unrolled with compile-time indices so the
// arrays live in registers and the DFT-32 twiddles fold into immediates.
//
// Instruction diet (the kernel is issue bound):
//   * decimation in time, so every non-trivial butterfly is  a' = a + w*b ;  b' = 2a - a'
//     = 6 FFMA instead of the 4 FADD + 2 FMUL + 2 FFMA of the decimation-in-frequency form;
//   * the first stage (twiddle 1) is fused with whatever scales its inputs: the window in the
//     first pass (3 instead of 4 instructions per real pair), the four-step twiddle in the
//     second pass (10 instead of 12 per complex pair).
// The functions are __host__ __device__ so tests/test_fft_host.py can run the exact index
// arithmetic on the CPU against a double-precision DFT.
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#define EVF_HD __host__ __device__ __forceinline__

namespace evf {

__host__ __device__ constexpr int bitrev5(int x) {
  return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}

// cos(2*pi*j/32) and sin(2*pi*j/32), j = 0..16, correctly rounded from fp64.
EVF_HD constexpr float cos32(int j) {
  switch (j) {
    case 0: return 1.0f;
    case 1: return 0.98078528040323043f;
    case 2: return 0.92387953251128674f;
    case 3: return 0.83146961230254524f;
    case 4: return 0.70710678118654757f;
    case 5: return 0.55557023301960218f;
    case 6: return 0.38268343236508978f;
    case 7: return 0.19509032201612825f;
    case 8: return 0.0f;
    default: return -cos32(16 - j);  // 8 < j <= 16
  }
}
EVF_HD constexpr float sin32(int j) { return (j <= 8) ? cos32(8 - j) : cos32(j - 8); }

// The butterflies are written over a value type V with the operations below (V = float).
EVF_HD float v_add(float a, float b) { return a + b; }
EVF_HD float v_sub(float a, float b) { return a - b; }
EVF_HD float v_neg(float a) { return -a; }
EVF_HD float v_mul(float a, float s) { return a * s; }                 // a * s
EVF_HD float v_fma(float a, float s, float c) { return fmaf(a, s, c); }  // a * s + c, s scalar

// One DIT butterfly on (a, b) with w = exp(-2*pi*i*J/32):  a' = a + w b,  b' = a - w b.
template <int J, typename V>
EVF_HD void bfly_dit(V& ar, V& ai, V& br, V& bi) {
  if constexpr (J == 0) {
    const V xr = v_add(ar, br), xi = v_add(ai, bi);
    br = v_sub(ar, br);
    bi = v_sub(ai, bi);
    ar = xr;
    ai = xi;
  } else if constexpr (J == 8) {  // w = -i:  w b = (bi, -br)
    const V xr = v_add(ar, bi), xi = v_sub(ai, br);
    const V yr = v_sub(ar, bi), yi = v_add(ai, br);
    ar = xr;
    ai = xi;
    br = yr;
    bi = yi;
  } else {  // w = c - i s:  w b = (br c + bi s) + i (bi c - br s)
    constexpr float c = cos32(J), s = sin32(J);
    const V xr = v_fma(br, c, v_fma(bi, s, ar));
    const V xi = v_fma(bi, c, v_fma(br, -s, ai));
    br = v_fma(ar, 2.0f, v_neg(xr));
    bi = v_fma(ai, 2.0f, v_neg(xi));
    ar = xr;
    ai = xi;
  }
}

template <int HALF, int G, int J>
struct DitJ {
  template <typename V, int N>
  EVF_HD static void run(V (&re)[N], V (&im)[N]) {
    bfly_dit<J*(16 / HALF)>(re[G + J], im[G + J], re[G + J + HALF], im[G + J + HALF]);
    if constexpr (J + 1 < HALF) DitJ<HALF, G, J + 1>::run(re, im);
  }
};
template <int HALF, int G>
struct DitG {
  template <typename V, int N>
  EVF_HD static void run(V (&re)[N], V (&im)[N]) {
    DitJ<HALF, G, 0>::run(re, im);
    if constexpr (G + 2 * HALF < N) DitG<HALF, G + 2 * HALF>::run(re, im);
  }
};

// Stages 2..5 of the forward DIT DFT-32 (butterfly spans 2, 4, 8, 16), in place.
// In : index i holds the output of the span-1 stage for the pair it belongs to, where the
//      span-1 stage combined x[bitrev5(i & ~1)] (= some n < 16) and x[n + 16].
// Out: index k holds X[k] (natural order).
template <typename V>
EVF_HD void dft32_dit_tail(V (&re)[32], V (&im)[32]) {
  DitG<2, 0>::run(re, im);
  DitG<4, 0>::run(re, im);
  DitG<8, 0>::run(re, im);
  DitG<16, 0>::run(re, im);
}

// The same for J = 1, 2, 4 independent DFT-(32 / J) on the register blocks [j * 32 / J, (j + 1) * 32 / J): a DIT stage
// of span H acts on every aligned block of 2H registers alike (its twiddles depend on the position inside the block
// only), so J smaller transforms side by side are the DFT-32 without its last log2(J) stages.
// In : block j, index i holds the span-1 output of the pair of rows (bitrev(i & ~1), + 16 / J) of transform j.
template <int J, typename V>
EVF_HD void dft32_dit_tail_jobs(V (&re)[32], V (&im)[32]) {
  static_assert(J == 1 || J == 2 || J == 4, "1, 2 or 4 transforms per register file");
  DitG<2, 0>::run(re, im);
  DitG<4, 0>::run(re, im);
  if constexpr (J <= 2) DitG<8, 0>::run(re, im);
  if constexpr (J == 1) DitG<16, 0>::run(re, im);
}

// bit reversal of the low `bits` bits
__host__ __device__ constexpr int bitrev_n(int x, int bits) {
  int r = 0;
  for (int b = 0; b < bits; ++b) r |= ((x >> b) & 1) << (bits - 1 - b);
  return r;
}

// Plain first stage (span 1, twiddle 1) for callers without anything to fuse into it.
// In: index i holds x[bitrev5(i)].
template <typename V>
EVF_HD void dft32_dit_head(V (&re)[32], V (&im)[32]) { DitG<1, 0>::run(re, im); }

// First stage fused with a real scaling of the inputs (the analysis window): the pair at
// indices (i, i + 1), i even, combines rows n = bitrev5(i) < 16 and n + 16:
//   out[i] = wa*xa + wb*xb ,  out[i + 1] = wa*xa - wb*xb          (3 instructions)
template <typename V, typename W>
EVF_HD void win_head(V& lo, V& hi, V xa, W wa, V xb, W wb) {
  const V t = v_mul(xa, wa);
  lo = v_fma(xb, wb, t);
  hi = v_fma(xb, v_neg(wb), t);
}

// First stage fused with a complex scaling (the four-step twiddles ta, tb = (cos, sin) of a
// negative angle, i.e. t = c + i s with s <= 0 for the forward transform):
//   A = ta*xa ;  out[i] = A + tb*xb ;  out[i + 1] = 2A - out[i]      (10 instructions)
template <bool kUnitA, typename V, typename W>
EVF_HD void tw_head(V& lor, V& loi, V& hir, V& hii, V xar, V xai, W tac, W tas, V xbr, V xbi, W tbc, W tbs) {
  V Ar = xar, Ai = xai;
  if constexpr (!kUnitA) {
    Ar = v_fma(xai, v_neg(tas), v_mul(xar, tac));
    Ai = v_fma(xar, tas, v_mul(xai, tac));
  }
  const V pr = v_fma(xbr, tbc, v_fma(xbi, v_neg(tbs), Ar));
  const V pi = v_fma(xbr, tbs, v_fma(xbi, tbc, Ai));
  hir = v_fma(Ar, 2.0f, v_neg(pr));
  hii = v_fma(Ai, 2.0f, v_neg(pi));
  lor = pr;
  loi = pi;
}

}  // namespace evf
