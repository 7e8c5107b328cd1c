// Host-side check of the index arithmetic of the warp-level 1024-point FFT (evfeat_fft.cuh):
// the 32 lanes of a warp are simulated one after the other, the shared-memory transpose is a
// plain array, and the result is compared with a double-precision DFT.  Test infrastructure
// (built and run by tests/test_fft_host.py with nvcc; no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "evfeat_fft.cuh"

using namespace evf;

// J packed transforms of N = 1024 / J points in one simulated warp (n_fft 512: J = 2, 256: J = 4): the index arithmetic
// of features_kernel's J > 1 branch -- register block jj = job jj in the first pass, lane (jj, k1) afterwards.
template <int J>
int check_jobs() {
  constexpr int N = 1024 / J, R1 = 32 / J, HR = R1 / 2, LB = (R1 == 16) ? 3 : 2;
  std::vector<double> xr(J * N), xi(J * N), w(N);
  srand(99 + J);
  for (int n = 0; n < J * N; ++n) {
    xr[n] = rand() / (double)RAND_MAX * 2 - 1;
    xi[n] = rand() / (double)RAND_MAX * 2 - 1;
  }
  for (int n = 0; n < N; ++n) w[n] = 0.5 - 0.5 * std::cos(2 * M_PI * n / N);
  static float Yr[32][32], Yi[32][32];  // [register][lane]
  for (int lane = 0; lane < 32; ++lane) {
    float re[32], im[32];
    for (int jj = 0; jj < J; ++jj)
      for (int r = 0; r < HR; ++r) {
        const int i = jj * R1 + 2 * bitrev_n(r, LB);
        const int na = 32 * r + lane, nb = 32 * (r + HR) + lane;
        win_head(re[i], re[i + 1], (float)xr[jj * N + na], (float)w[na], (float)xr[jj * N + nb], (float)w[nb]);
        win_head(im[i], im[i + 1], (float)xi[jj * N + na], (float)w[na], (float)xi[jj * N + nb], (float)w[nb]);
      }
    dft32_dit_tail_jobs<J>(re, im);
    for (int p = 0; p < 32; ++p) {
      Yr[p][lane] = re[p];
      Yi[p][lane] = im[p];
    }
  }
  std::vector<double> Zr(J * N), Zi(J * N);
  for (int lane = 0; lane < 32; ++lane) {
    const int jj = lane / R1, k1 = lane % R1;
    float re[32], im[32], tr[32], ti[32];
    for (int n2 = 0; n2 < 32; ++n2) {
      tr[n2] = Yr[lane][n2];
      ti[n2] = Yi[lane][n2];
    }
    for (int n = 0; n < 16; ++n) {
      float c[2], s[2];
      for (int h = 0; h < 2; ++h) {
        const double ang = -2 * M_PI * ((k1 * (n + 16 * h)) % N) / N;
        c[h] = (float)std::cos(ang);
        s[h] = (float)std::sin(ang);
      }
      const int i = bitrev5(n);
      if (n == 0)
        tw_head<true>(re[i], im[i], re[i + 1], im[i + 1], tr[n], ti[n], c[0], s[0], tr[n + 16], ti[n + 16], c[1], s[1]);
      else
        tw_head<false>(re[i], im[i], re[i + 1], im[i + 1], tr[n], ti[n], c[0], s[0], tr[n + 16], ti[n + 16], c[1], s[1]);
    }
    dft32_dit_tail(re, im);
    for (int k2 = 0; k2 < 32; ++k2) {
      Zr[jj * N + k1 + R1 * k2] = re[k2];
      Zi[jj * N + k1 + R1 * k2] = im[k2];
    }
  }
  double worst = 0, scale = 0;
  for (int jj = 0; jj < J; ++jj)
    for (int k = 0; k < N; ++k) {
      double ar = 0, ai = 0;
      for (int n = 0; n < N; ++n) {
        const double a = -2 * M_PI * ((long long)k * n % N) / N;
        const double vr = xr[jj * N + n] * w[n], vi = xi[jj * N + n] * w[n];
        ar += vr * std::cos(a) - vi * std::sin(a);
        ai += vr * std::sin(a) + vi * std::cos(a);
      }
      worst = std::fmax(worst, std::fmax(std::fabs(ar - Zr[jj * N + k]), std::fabs(ai - Zi[jj * N + k])));
      scale = std::fmax(scale, std::hypot(ar, ai));
    }
  printf("%d x fft%d max abs err %.3e (max |Z| %.3f)\n", J, N, worst, scale);
  return worst < 2e-4 * (scale / 30 + 1) ? 0 : 1;
}

// The backward's second transform for two packed jobs (evfeat_backward.cu): after the first transform lane (j, k1)
// holds elements k1 + 16 q of ITS job; the inverse (run through the forward code) wants lane n2 to hold rows n1 of BOTH
// jobs.  Element k1 + 16 (2 n1 + b) is parked at position b * 16 + bitrev4(n1); the values with b != job change lane
// halves.  Checked: the transform of two arbitrary 512-point complex inputs laid out that way is their DFT.
int check_relayout_two_jobs() {
  constexpr int J = 2, N = 512, R1 = 16;
  std::vector<double> cr(J * N), ci(J * N);
  srand(4242);
  for (int n = 0; n < J * N; ++n) {
    cr[n] = rand() / (double)RAND_MAX * 2 - 1;
    ci[n] = rand() / (double)RAND_MAX * 2 - 1;
  }
  static float re[32][32], im[32][32];  // [lane][position]
  auto pos = [](int q) { return (q & 1) * 16 + bitrev_n(q >> 1, 4); };
  for (int lane = 0; lane < 32; ++lane) {
    const int j = lane / R1, k1 = lane % R1;
    for (int q = 0; q < 32; ++q) {
      re[lane][pos(q)] = (float)cr[j * N + k1 + R1 * q];
      im[lane][pos(q)] = (float)ci[j * N + k1 + R1 * q];
    }
  }
  for (int n1 = 0; n1 < 16; ++n1) {  // the exchange (a __shfl_xor by 16 on the device)
    const int p0 = bitrev_n(n1, 4), p1 = 16 + p0;
    float sr[32], si[32];
    for (int lane = 0; lane < 32; ++lane) {
      const bool upper = lane >= 16;
      sr[lane] = upper ? re[lane][p0] : re[lane][p1];
      si[lane] = upper ? im[lane][p0] : im[lane][p1];
    }
    for (int lane = 0; lane < 32; ++lane) {
      const bool upper = lane >= 16;
      (upper ? re[lane][p0] : re[lane][p1]) = sr[lane ^ 16];
      (upper ? im[lane][p0] : im[lane][p1]) = si[lane ^ 16];
    }
  }
  static float Yr[32][32], Yi[32][32];
  for (int lane = 0; lane < 32; ++lane) {
    dft32_dit_head(re[lane], im[lane]);
    dft32_dit_tail_jobs<J>(re[lane], im[lane]);
    for (int p = 0; p < 32; ++p) {
      Yr[p][lane] = re[lane][p];
      Yi[p][lane] = im[lane][p];
    }
  }
  double worst = 0, scale = 0;
  for (int lane = 0; lane < 32; ++lane) {
    const int jj = lane / R1, k1 = lane % R1;
    float xr[32], xi[32], tr[32], ti[32];
    for (int n2 = 0; n2 < 32; ++n2) {
      tr[n2] = Yr[lane][n2];
      ti[n2] = Yi[lane][n2];
    }
    for (int n = 0; n < 16; ++n) {
      float c[2], sn[2];
      for (int h = 0; h < 2; ++h) {
        const double ang = -2 * M_PI * ((k1 * (n + 16 * h)) % N) / N;
        c[h] = (float)std::cos(ang);
        sn[h] = (float)std::sin(ang);
      }
      const int i = bitrev5(n);
      if (n == 0)
        tw_head<true>(xr[i], xi[i], xr[i + 1], xi[i + 1], tr[n], ti[n], c[0], sn[0], tr[n + 16], ti[n + 16], c[1], sn[1]);
      else
        tw_head<false>(xr[i], xi[i], xr[i + 1], xi[i + 1], tr[n], ti[n], c[0], sn[0], tr[n + 16], ti[n + 16], c[1], sn[1]);
    }
    dft32_dit_tail(xr, xi);
    for (int k2 = 0; k2 < 32; ++k2) {
      const int k = k1 + R1 * k2;
      double ar = 0, ai = 0;
      for (int n = 0; n < N; ++n) {
        const double a = -2 * M_PI * ((long long)k * n % N) / N;
        ar += cr[jj * N + n] * std::cos(a) - ci[jj * N + n] * std::sin(a);
        ai += cr[jj * N + n] * std::sin(a) + ci[jj * N + n] * std::cos(a);
      }
      worst = std::fmax(worst, std::fmax(std::fabs(ar - xr[k2]), std::fabs(ai - xi[k2])));
      scale = std::fmax(scale, std::hypot(ar, ai));
    }
  }
  printf("re-layout + 2 x fft512 max abs err %.3e (max |Z| %.3f)\n", worst, scale);
  return worst < 2e-4 * (scale / 30 + 1) ? 0 : 1;
}

int main() {
  if (check_jobs<2>() || check_jobs<4>() || check_relayout_two_jobs()) return 1;
  const int N = 1024;
  std::vector<double> xr(N), xi(N), w(N);
  srand(1234);
  for (int n = 0; n < N; ++n) {
    xr[n] = rand() / (double)RAND_MAX * 2 - 1;
    xi[n] = rand() / (double)RAND_MAX * 2 - 1;
    w[n] = 0.5 - 0.5 * std::cos(2 * M_PI * n / N);
  }
  // pass 1, lane = n2: rows n1 of z[32*n1 + n2], window fused into the first stage
  static float Yr[32][32], Yi[32][32];  // [k1][n2]
  for (int lane = 0; lane < 32; ++lane) {
    float re[32], im[32];
    for (int r = 0; r < 16; ++r) {
      const int i = bitrev5(r);
      const int na = 32 * r + lane, nb = 32 * (r + 16) + lane;
      win_head(re[i], re[i + 1], (float)xr[na], (float)w[na], (float)xr[nb], (float)w[nb]);
      win_head(im[i], im[i + 1], (float)xi[na], (float)w[na], (float)xi[nb], (float)w[nb]);
    }
    dft32_dit_tail(re, im);
    for (int p = 0; p < 32; ++p) {
      Yr[p][lane] = re[p];
      Yi[p][lane] = im[p];
    }
  }
  // pass 2, lane = k1
  std::vector<double> Zr(N), Zi(N);
  for (int lane = 0; lane < 32; ++lane) {
    float re[32], im[32], tr[32], ti[32];
    for (int n2 = 0; n2 < 32; ++n2) {
      tr[n2] = Yr[lane][n2];
      ti[n2] = Yi[lane][n2];
    }
    for (int n = 0; n < 16; ++n) {
      float c[2], s[2];
      for (int h = 0; h < 2; ++h) {
        const double ang = -2 * M_PI * ((lane * (n + 16 * h)) % N) / N;
        c[h] = (float)std::cos(ang);
        s[h] = (float)std::sin(ang);
      }
      const int i = bitrev5(n);
      if (n == 0)
        tw_head<true>(re[i], im[i], re[i + 1], im[i + 1], tr[n], ti[n], c[0], s[0], tr[n + 16], ti[n + 16], c[1], s[1]);
      else
        tw_head<false>(re[i], im[i], re[i + 1], im[i + 1], tr[n], ti[n], c[0], s[0], tr[n + 16], ti[n + 16], c[1], s[1]);
    }
    dft32_dit_tail(re, im);
    for (int k2 = 0; k2 < 32; ++k2) {
      Zr[lane + 32 * k2] = re[k2];
      Zi[lane + 32 * k2] = im[k2];
    }
  }
  // plain DFT-32 through head + tail as well
  double worst32 = 0;
  {
    float re[32], im[32];
    for (int i = 0; i < 32; ++i) {
      re[i] = (float)xr[bitrev5(i)];
      im[i] = (float)xi[bitrev5(i)];
    }
    dft32_dit_head(re, im);
    dft32_dit_tail(re, im);
    for (int k = 0; k < 32; ++k) {
      double ar = 0, ai = 0;
      for (int n = 0; n < 32; ++n) {
        const double a = -2 * M_PI * k * n / 32;
        ar += xr[n] * std::cos(a) - xi[n] * std::sin(a);
        ai += xr[n] * std::sin(a) + xi[n] * std::cos(a);
      }
      worst32 = std::fmax(worst32, std::fmax(std::fabs(ar - re[k]), std::fabs(ai - im[k])));
    }
  }
  double worst = 0, scale = 0;
  for (int k = 0; k < N; ++k) {
    double ar = 0, ai = 0;
    for (int n = 0; n < N; ++n) {
      const double a = -2 * M_PI * ((long long)k * n % N) / N;
      const double vr = xr[n] * w[n], vi = xi[n] * w[n];
      ar += vr * std::cos(a) - vi * std::sin(a);
      ai += vr * std::sin(a) + vi * std::cos(a);
    }
    worst = std::fmax(worst, std::fmax(std::fabs(ar - Zr[k]), std::fabs(ai - Zi[k])));
    scale = std::fmax(scale, std::hypot(ar, ai));
  }
  printf("dft32 max abs err %.3e ; fft1024 max abs err %.3e (max |Z| %.3f)\n", worst32, worst, scale);
  return (worst32 < 2e-5 && worst < 2e-4 * (scale / 30 + 1)) ? 0 : 1;
}
