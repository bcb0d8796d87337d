// Host-only check (no GPU): the register-level pieces of the one-warp 512-point FFT (s2k_fft16.cuh) run for 32 emulated
// lanes -- same index maps, same butterflies, same twiddle recurrences as the device code, with arrays standing in for
// the shared-memory exchange and the shuffle -- against a long-double DFT.  Built and run by tests/test_boundary.py.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../s2kit_b200/csrc/s2k_fft16.cuh"

using namespace s2k;

int main() {
    const int N = F16_N;
    std::vector<double> xr(N), xi(N);
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&]() {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        return (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
    };
    for (int i = 0; i < N; ++i) { xr[i] = rnd(); xi[i] = rnd(); }

    // dft16 alone
    double maxe16 = 0;
    {
        double a[16], b[16];
        for (int e = 0; e < 16; ++e) { a[e] = xr[e]; b[e] = xi[e]; }
        f16_dft16(a, b);
        for (int k = 0; k < 16; ++k) {
            long double sr = 0, si = 0;
            for (int e = 0; e < 16; ++e) {
                long double ang = -2.0L * M_PIl * (long double)(e * k) / 16.0L;
                sr += xr[e] * cosl(ang) - xi[e] * sinl(ang);
                si += xr[e] * sinl(ang) + xi[e] * cosl(ang);
            }
            maxe16 = std::fmax(maxe16, std::fmax(std::fabs((double)(sr - a[k])), std::fabs((double)(si - b[k]))));
        }
    }

    double regr[32][16], regi[32][16];
    std::vector<double> exr(F16_EX_ELEMS, 0.0), exi(F16_EX_ELEMS, 0.0);
    for (int t = 0; t < 32; ++t) {
        for (int e = 0; e < 16; ++e) { regr[t][e] = xr[f16_in_index(t, e)]; regi[t][e] = xi[f16_in_index(t, e)]; }
        const double ang = 2.0 * M_PI * t / N;
        f16_phase1(regr[t], regi[t], std::cos(ang), -std::sin(ang));
        for (int k1 = 0; k1 < 16; ++k1) { exr[f16_ex_write(t, k1)] = regr[t][k1]; exi[f16_ex_write(t, k1)] = regi[t][k1]; }
    }
    for (int L = 0; L < 32; ++L) {
        for (int j = 0; j < 16; ++j) { regr[L][j] = exr[f16_ex_read(L, j)]; regi[L][j] = exi[f16_ex_read(L, j)]; }
        f16_dft16(regr[L], regi[L]);
    }
    double pr[32][8], pi_[32][8];
    for (int L = 0; L < 32; ++L) {  // what lane L receives from lane L ^ 16
        const int P = L ^ 16, hs = P >> 4;
        for (int qi = 0; qi < 8; ++qi) { pr[L][qi] = hs ? regr[P][qi] : regr[P][8 + qi]; pi_[L][qi] = hs ? regi[P][qi] : regi[P][8 + qi]; }
    }
    std::vector<double> Xr(N, 0.0), Xi(N, 0.0);
    std::vector<int> seen(N, 0);
    for (int L = 0; L < 32; ++L) {
        f16_phase3(regr[L], regi[L], pr[L], pi_[L], L >> 4);
        for (int o = 0; o < 16; ++o) { const int k = f16_out_index(L, o); Xr[k] = regr[L][o]; Xi[k] = regi[L][o]; ++seen[k]; }
    }
    double maxe = 0, scale = 0;
    int bad_cover = 0;
    for (int k = 0; k < N; ++k) {
        if (seen[k] != 1) ++bad_cover;
        long double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            long double ang = -2.0L * M_PIl * (long double)((long long)n * k % N) / (long double)N;
            sr += xr[n] * cosl(ang) - xi[n] * sinl(ang);
            si += xr[n] * sinl(ang) + xi[n] * cosl(ang);
        }
        maxe = std::fmax(maxe, std::fmax(std::fabs((double)(sr - Xr[k])), std::fabs((double)(si - Xi[k]))));
        scale = std::fmax(scale, std::fmax(std::fabs((double)sr), std::fabs((double)si)));
    }
    std::printf("dft16 max abs err %.3e; fft512 max abs err %.3e (scale %.1f); outputs not covered exactly once: %d\n", maxe16,
                maxe, scale, bad_cover);
    return (maxe16 < 1e-13 && maxe < 1e-11 && bad_cover == 0) ? 0 : 1;
}
