// CPU self-test of the register FFT building blocks in csrc/fft500.cuh: runs the exact pass-A / pass-B / untangle
// index logic of the frontend kernel with "threads" emulated by loops and compares with a direct O(N^2) DFT in double.
//   nvcc -O2 -std=c++17 -o /tmp/fft_selftest tools/fft_selftest.cu && /tmp/fft_selftest
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../yolo-inspired-audio-activity-detection_b200/csrc/fft500.cuh"

using namespace yad;

int main() {
  const int N = 1000;
  std::vector<double> x(N);
  srand(7);
  for (int i = 0; i < N; ++i) x[i] = (rand() / (double)RAND_MAX - 0.5) * 2.0 + 0.3 * sin(0.37 * i);
  // reference: power spectrum of the 1000-point real DFT, bins 0..500
  std::vector<double> pref(501);
  for (int k = 0; k <= 500; ++k) {
    double re = 0, im = 0;
    for (int n = 0; n < N; ++n) {
      const double a = -2.0 * M_PI * (double)((long long)k * n % N) / N;
      re += x[n] * cos(a);
      im += x[n] * sin(a);
    }
    pref[k] = re * re + im * im;
  }
  // twiddle tables as the kernel has them
  std::vector<cf32> tw1000(1000);
  for (int k = 0; k < 1000; ++k) tw1000[k] = cmake((float)cos(-2.0 * M_PI * k / 1000.0), (float)sin(-2.0 * M_PI * k / 1000.0));
  std::vector<cf32> twA(25 * 20);   // [reg r][n2] = W_500^(n2 * k1(r))
  for (int r = 0; r < 25; ++r)
    for (int n2 = 0; n2 < 20; ++n2) twA[r * 20 + n2] = tw1000[(2 * n2 * passA_k1_of_reg(r)) % 1000];

  std::vector<cf32> z(FFT_NZ), Y(FFT_Y_STRIDE), Z(FFT_NZ);
  for (int n = 0; n < FFT_NZ; ++n) z[n] = cmake((float)x[2 * n], (float)x[2 * n + 1]);
  // pass A
  for (int n2 = 0; n2 < 20; ++n2) {
    cf32 v[25];
    for (int n1 = 0; n1 < 25; ++n1) v[n1] = z[20 * n1 + n2];
    dft25(v);
    for (int r = 0; r < 25; ++r) {
      const cf32 t = twA[r * 20 + n2];
      Y[passA_k1_of_reg(r) * FFT_Y_PITCH + n2] = cmulc(v[r], t.x, t.y);
    }
  }
  // pass B
  for (int k1 = 0; k1 < 25; ++k1) {
    cf32 v[20];
    for (int n2 = 0; n2 < 20; ++n2) v[n2] = Y[k1 * FFT_Y_PITCH + n2];
    dft20(v);
    for (int r = 0; r < 20; ++r) Z[k1 + 25 * passB_k2_of_reg(r)] = v[r];
  }
  // check the complex FFT itself
  double maxc = 0;
  for (int k = 0; k < 500; ++k) {
    double re = 0, im = 0;
    for (int n = 0; n < 500; ++n) {
      const double a = -2.0 * M_PI * (double)((long long)k * n % 500) / 500;
      re += x[2 * n] * cos(a) - x[2 * n + 1] * sin(a);
      im += x[2 * n] * sin(a) + x[2 * n + 1] * cos(a);
    }
    maxc = fmax(maxc, fmax(fabs(re - Z[k].x), fabs(im - Z[k].y)));
  }
  // untangle + power (same arithmetic as the kernel)
  double maxrel = 0;
  for (int k = 0; k <= 250; ++k) {
    const cf32 zk = Z[k], zq = Z[(500 - k) % 500];
    const cf32 zn = cmake(zq.x, -zq.y);
    const cf32 E = cmake(0.5f * (zk.x + zn.x), 0.5f * (zk.y + zn.y));
    const cf32 D = cmake(zk.x - zn.x, zk.y - zn.y);
    const cf32 Od = cmake(0.5f * D.y, -0.5f * D.x);
    const cf32 Tt = cmulc(Od, tw1000[k].x, tw1000[k].y);
    const float pr = E.x + Tt.x, pi = E.y + Tt.y, qr = E.x - Tt.x, qi = E.y - Tt.y;
    const double Pk = (double)pr * pr + (double)pi * pi, Pq = (double)qr * qr + (double)qi * qi;
    maxrel = fmax(maxrel, fabs(Pk - pref[k]) / (pref[k] + 1e-3));
    maxrel = fmax(maxrel, fabs(Pq - pref[500 - k]) / (pref[500 - k] + 1e-3));
  }
  printf("complex FFT max abs err %.3e ; power max rel err %.3e\n", maxc, maxrel);
  const bool ok = maxc < 2e-4 && maxrel < 2e-5;
  printf(ok ? "OK\n" : "FAIL\n");
  return ok ? 0 : 1;
}
