// Host-only check (no GPU) of s2k_dct16.cuh: the forward DCT pair on the one-warp 512-point FFT for 32 emulated lanes --
// the same index maps (in-place exchange slots, Z[N-k] shuffle sources / offered registers, panel slots), butterflies
// and rotation constants as the device code, with arrays standing in for shared memory and the shuffles -- against the
// long-double definition of the orthonormal DCT-II (FFTW REDFT10 + the scaling of seminaive.c:170-176).
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../s2kit_b200/csrc/s2k_dct16.cuh"

using namespace s2k;

int main() {
    const int N = 512, B = 256, CS = 132, NC = 32, PS = NC * CS + 8;
    std::vector<double> a(N), b(N);
    unsigned long long s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        return (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
    };
    for (int i = 0; i < N; ++i) { a[i] = rnd(); b[i] = rnd(); }
    // rotation constants
    double cerr = 0;
    for (int qi = 0; qi < 8; ++qi) {
        double qr, qim;
        d16_quarter_rot(1.0, 0.0, qi, qr, qim);
        cerr = std::fmax(cerr, std::fmax(std::fabs(qr - std::cos(M_PI * qi / 64.0)), std::fabs(qim - std::sin(M_PI * qi / 64.0))));
    }
    // bank check of the in-place exchange: the 16 lanes of a half-warp read distinct 8-byte bank pairs
    int bank_conflicts = 0;
    for (int j = 0; j < 16; ++j)
        for (int hw = 0; hw < 2; ++hw) {
            int seen[16] = {0};
            for (int l = 16 * hw; l < 16 * hw + 16; ++l) bank_conflicts += seen[d16_ex_read(l, j, PS) & 15]++;
        }
    // coverage of the exchange: every (k1, t) written once, read once, inside the two column copies
    std::vector<int> cover(2 * PS, 0);
    int ex_bad = 0;
    for (int t = 0; t < 32; ++t)
        for (int k1 = 0; k1 < 16; ++k1) {
            const int at = d16_ex_write(t, k1, PS);
            if (!((at >= 0 && at < 2 * CS) || (at >= PS && at < PS + 2 * CS))) ++ex_bad;
            cover[at]++;
        }
    for (int l = 0; l < 32; ++l)
        for (int j = 0; j < 16; ++j) {
            // lane (k1, h) must receive A[t = h + 2j][k1]
            if (d16_ex_read(l, j, PS) != d16_ex_write((l >> 4) + 2 * j, l & 15, PS)) ++ex_bad;
        }
    for (int v : cover)
        if (v > 1) ++ex_bad;

    double regr[32][16], regi[32][16];
    std::vector<double> panel(2 * PS, 1e300);  // the warp's column pair: parity-0 copy at 0, parity-1 copy at PS
    for (int t = 0; t < 32; ++t) {
        for (int e = 0; e < 16; ++e) {
            const int p = t + 32 * e, j = (p < B) ? 2 * p : 2 * (N - 1 - p) + 1;
            regr[t][e] = a[j];
            regi[t][e] = b[j];
        }
        const double ang = 2.0 * M_PI * t / N;
        f16_phase1(regr[t], regi[t], std::cos(ang), -std::sin(ang));
    }
    for (int part = 0; part < 2; ++part) {
        for (int t = 0; t < 32; ++t)
            for (int k1 = 0; k1 < 16; ++k1) panel[d16_ex_write(t, k1, PS)] = part ? regi[t][k1] : regr[t][k1];
        double tmp[32][16];
        for (int l = 0; l < 32; ++l)
            for (int j = 0; j < 16; ++j) tmp[l][j] = panel[d16_ex_read(l, j, PS)];
        for (int l = 0; l < 32; ++l)
            for (int j = 0; j < 16; ++j) (part ? regi : regr)[l][j] = tmp[l][j];
    }
    for (int l = 0; l < 32; ++l) f16_dft16(regr[l], regi[l]);
    {
        double pr[32][8], pi_[32][8];
        for (int l = 0; l < 32; ++l) {
            const int P = l ^ 16, hs = P >> 4;
            for (int qi = 0; qi < 8; ++qi) {
                pr[l][qi] = hs ? regr[P][qi] : regr[P][8 + qi];
                pi_[l][qi] = hs ? regi[P][qi] : regi[P][8 + qi];
            }
        }
        for (int l = 0; l < 32; ++l) f16_phase3(regr[l], regi[l], pr[l], pi_[l], l >> 4);
    }
    const double s_all = 1.0 / std::sqrt(2.0 * N);
    for (int qi = 0; qi < 8; ++qi) {
        double offr[32], offi[32];
        for (int l = 0; l < 32; ++l) {
            offr[l] = regr[l][d16_offer_reg(l, qi)];
            offi[l] = regi[l][d16_offer_reg(l, qi)];
        }
        for (int l = 0; l < 32; ++l) {
            const int k1 = l & 15, h = l >> 4, kk = k1 + 16 * (qi + 8 * h), src = d16_src_lane(l, qi);
            double br = offr[src], bi = offi[src];
            const double ar = regr[l][2 * qi], ai = regi[l][2 * qi];
            if (kk == 0) { br = ar; bi = ai; }
            const double q0r = std::cos(M_PI * (k1 + 128 * h) / (2.0 * N)), q0i = std::sin(M_PI * (k1 + 128 * h) / (2.0 * N));
            double qr, qim, y1, y2;
            d16_quarter_rot(q0r, q0i, qi, qr, qim);
            d16_separate(ar, ai, br, bi, qr, qim, kk, s_all, y1, y2);
            // same addresses as the device code: out = col0 + (k1 & 1) * ps + (k1 >> 1) + 64 h; out[8 qi], out[8 qi + CS]
            const int at = (k1 & 1) * PS + (k1 >> 1) + 64 * h + 8 * qi;
            if (at != d16_panel_slot(kk, PS)) ++ex_bad;
            panel[at] = y1;
            panel[at + CS] = y2;
        }
    }
    double maxe = 0, scale = 0;
    for (int k = 0; k < B; ++k) {
        long double y1 = 0, y2 = 0;
        for (int j = 0; j < N; ++j) {
            const long double c = cosl(M_PIl * (2.0L * j + 1.0L) * k / (2.0L * N));
            y1 += a[j] * c;
            y2 += b[j] * c;
        }
        y1 *= 2.0L; y2 *= 2.0L;
        if (k == 0) { y1 *= M_SQRT1_2l; y2 *= M_SQRT1_2l; }
        y1 /= sqrtl(2.0L * N); y2 /= sqrtl(2.0L * N);
        const int at = d16_panel_slot(k, PS);
        maxe = std::fmax(maxe, std::fmax(std::fabs((double)y1 - panel[at]), std::fabs((double)y2 - panel[at + CS])));
        scale = std::fmax(scale, std::fabs((double)y1));
    }
    std::printf("rotation constants err %.2e; exchange bank conflicts %d; exchange index errors %d; dct16 max abs err %.3e "
                "(scale %.2f)\n", cerr, bank_conflicts, ex_bad, maxe, scale);
    const bool ok = cerr < 1e-16 && bank_conflicts == 0 && ex_bad == 0 && maxe < 1e-13;
    std::printf("dct16 check: %s\n", ok ? "OK" : "FAILED");
    return ok ? 0 : 1;
}
