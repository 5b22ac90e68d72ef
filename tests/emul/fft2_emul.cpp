// CPU emulation of the FFT v2 device code (comms-rs_b200/csrc/fft2_core.cuh): every pass is run for
// every thread of a frame in turn (a pass boundary = __syncthreads), with ping-pong shared-memory
// buffers, and the result is compared with an O(N^2) DFT in double.  Also reports the worst
// shared-memory bank-conflict degree per half-warp (8-byte words, 16 banks of 8 bytes).
// Build: g++ -O2 -std=c++17 -I/usr/local/cuda/include fft2_emul.cpp -o fft2_emul
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "../../comms-rs_b200/csrc/fft2_core.cuh"

using namespace cb::fft2;

struct LogSm {
    float2 *base;
    std::vector<std::vector<int>> *log;  // [thread] -> sequence of padded indices
    int tid;
    float2 ld(int i) const { (*log)[tid].push_back(pad16(i)); return base[pad16(i)]; }
    void st(int i, float2 v) const { (*log)[tid].push_back(pad16(i)); base[pad16(i)] = v; }
};

static int worst_conflict(const std::vector<std::vector<int>> &log, int T)
{
    int worst = 1;
    if (log.empty() || log[0].empty()) return 0;
    const size_t nacc = log[0].size();
    for (size_t a = 0; a < nacc; ++a)
        for (int h0 = 0; h0 < T; h0 += 16) {
            std::map<int, std::vector<int>> banks;
            for (int t = h0; t < h0 + 16 && t < T; ++t) {
                const int w = log[t][a];
                auto &v = banks[w % 16];
                bool seen = false;
                for (int x : v) seen |= (x == w);
                if (!seen) v.push_back(w);
            }
            for (auto &kv : banks) worst = std::max(worst, (int)kv.second.size());
        }
    return worst;
}

template <int LOG2N, bool INV>
static void make_tw(std::vector<float2> &tw)
{
    using PL = Plan<LOG2N>;
    tw.assign(PL::TW_TOTAL > 0 ? PL::TW_TOTAL : 1, make_float2(0, 0));
    const long double sgn = INV ? 2.0L : -2.0L, pi = 3.14159265358979323846264338327950288L;
    for (int p = 1; p < PL::P16; ++p) {
        const int Ns = 1 << (4 * p);
        for (int s = 0; s < Ns; ++s) {
            const long double a = sgn * pi * s / (16.0L * Ns);
            tw[PL::tw_off(p) + s] = make_float2((float)cosl(a), (float)sinl(a));
        }
    }
    if (PL::REM) {
        const int n = PL::N >> PL::REM;
        for (int s = 0; s < n; ++s) {
            const long double a = sgn * pi * s / PL::N;
            tw[PL::tw_off(PL::P16) + s] = make_float2((float)cosl(a), (float)sinl(a));
        }
    }
}

template <int LOG2N, bool INV, int PASS>
static void emul_pass(const std::vector<float2> &x, std::vector<float2> &out, std::vector<float2> *buf,
                      const std::vector<float2> &tw, int &worst)
{
    using PL = Plan<LOG2N>;
    if constexpr (PASS < PL::PASSES) {
        std::vector<std::vector<int>> lin(PL::T), lout(PL::T);
        for (int j = 0; j < PL::T; ++j) {
            auto gld = [&](int i) { return x[i]; };
            auto gst = [&](int i, float2 v) { out[i] = v; };
            LogSm sin{buf[(PASS + 1) & 1].data(), &lin, j}, sout{buf[PASS & 1].data(), &lout, j};
            auto mid = [] {};
            run_pass<LOG2N, INV, PASS>(j, tw.data(), gld, gst, sin, sout, mid);
        }
        // loads and stores are logged separately (different buffers)
        worst = std::max(worst, worst_conflict(lin, PL::T));
        worst = std::max(worst, worst_conflict(lout, PL::T));
        emul_pass<LOG2N, INV, PASS + 1>(x, out, buf, tw, worst);
    }
}

template <int LOG2N, bool INV>
static int check()
{
    using PL = Plan<LOG2N>;
    const int N = PL::N;
    std::vector<float2> x(N), out(N), tw;
    srand(1234 + LOG2N);
    for (auto &v : x) v = make_float2(rand() / (float)RAND_MAX * 2 - 1, rand() / (float)RAND_MAX * 2 - 1);
    make_tw<LOG2N, INV>(tw);
    std::vector<float2> buf[2] = {std::vector<float2>(PL::PADN + 16), std::vector<float2>(PL::PADN + 16)};
    int worst = 0;
    emul_pass<LOG2N, INV, 0>(x, out, buf, tw, worst);
    // reference: direct DFT in double on a subset of bins (all bins for N <= 2048)
    const int step = N <= 2048 ? 1 : 7;
    double num = 0, den = 0;
    const double sgn = INV ? 1.0 : -1.0;
    for (int k = 0; k < N; k += step) {
        std::complex<double> acc = 0;
        for (int n = 0; n < N; ++n) {
            const double a = sgn * 2.0 * M_PI * (double)(((long long)k * n) % N) / N;
            acc += std::complex<double>(x[n].x, x[n].y) * std::complex<double>(cos(a), sin(a));
        }
        const std::complex<double> got(out[k].x, out[k].y);
        num += std::norm(got - acc);
        den += std::norm(acc);
    }
    const double rel = sqrt(num / den);
    printf("fft2 emul N=%5d inv=%d passes=%d rel_l2=%.3e worst_bank_conflict=%d %s\n", N, (int)INV, PL::PASSES, rel,
           worst, rel < 2e-6 ? "OK" : "FAIL");
    return rel < 2e-6 ? 0 : 1;
}

int main()
{
    int bad = 0;
    bad += check<4, false>();
    bad += check<5, false>();
    bad += check<6, true>();
    bad += check<7, false>();
    bad += check<8, false>();
    bad += check<8, true>();
    bad += check<9, false>();
    bad += check<10, false>();
    bad += check<10, true>();
    bad += check<11, false>();
    bad += check<12, false>();
    bad += check<12, true>();
    bad += check<13, false>();
    bad += check<14, true>();  // the largest one-CTA size (4 passes: 16 16 16 4)
    return bad;
}
