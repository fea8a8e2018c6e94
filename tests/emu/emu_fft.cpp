// tests/emu/emu_fft.cpp -- host emulation of the device FFT building blocks (irbaboon_b200/csrc/irb_fft.cuh).
// TEST INFRASTRUCTURE: compiles the very same header with g++; "threads" are loop iterations and the
// CTA barriers of fft_run are loop boundaries.  Lets the index math of the Stockham passes, the packed
// real split/merge and the MAC be checked on a machine without a GPU.
#include <cstring>
#include <vector>
#include "../../irbaboon_b200/csrc/irb_fft.cuh"

using namespace irb;

template <int M, bool INV, int PASS = 0, int PS = 1>
static void emu_run(std::vector<float2>& regs, float2* srow, const float2* W) {
    constexpr int TPF = M / kPts;
    constexpr int R = pass_radix(M, PASS);
    for (int t = 0; t < TPF; ++t) fft_pass<M, R, PS, INV, PASS>(&regs[t * kPts], t, W);
    if constexpr (PS * R < M) {
        for (int t = 0; t < TPF; ++t) fft_scatter<M, R, PS>(&regs[t * kPts], t, srow);
        for (int t = 0; t < TPF; ++t) fft_gather_sw<M>(&regs[t * kPts], t, srow);
        emu_run<M, INV, PASS + 1, PS * R>(regs, srow, W);
    }
}

// x: len <= 2M real samples (zero padded) -> packed spectrum of M complex
template <int M>
static void fwd(const float* x, int len, const float2* W, float2* packed) {
    constexpr int TPF = M / kPts;
    std::vector<float2> regs(M), s(M);
    for (int t = 0; t < TPF; ++t)
        for (int j = 0; j < kPts; ++j) {
            const int m = 2 * (t + j * TPF);
            regs[t * kPts + j] = make_float2(m < len ? x[m] : 0.f, m + 1 < len ? x[m + 1] : 0.f);
        }
    emu_run<M, false>(regs, s.data(), W);
    for (int t = 0; t < TPF; ++t)
        for (int j = 0; j < kPts; ++j) s[t + j * TPF] = regs[t * kPts + j];
    for (int k = 0; k < M; ++k) packed[k] = real_split(s[k], s[(M - k) & (M - 1)], W[k], k);
}
// packed spectrum -> 2M real samples (scaled 1/N)
template <int M>
static void inv(const float2* packed, const float2* W, float* y) {
    constexpr int TPF = M / kPts;
    std::vector<float2> regs(M), s(packed, packed + M);
    for (int t = 0; t < TPF; ++t)
        for (int j = 0; j < kPts; ++j) {
            const int k = t + j * TPF;
            regs[t * kPts + j] = real_merge(s[k], s[(M - k) & (M - 1)], W[k], k);
        }
    emu_run<M, true>(regs, s.data(), W);
    const float scale = 1.0f / (float) (2 * M);
    for (int t = 0; t < TPF; ++t)
        for (int j = 0; j < kPts; ++j) {
            const int n = t + j * TPF;
            y[2 * n] = regs[t * kPts + j].x * scale;
            y[2 * n + 1] = regs[t * kPts + j].y * scale;
        }
}

static void make_w(int M, std::vector<float2>& W) {
    W.resize((size_t) fft_table_size(M));
    fft_build_table(M, W.data());
}

#define DISPATCH(M_, CALL)                         \
    switch (M_) {                                  \
        case 16: { constexpr int MM = 16; CALL; break; }     \
        case 32: { constexpr int MM = 32; CALL; break; }     \
        case 64: { constexpr int MM = 64; CALL; break; }     \
        case 128: { constexpr int MM = 128; CALL; break; }   \
        case 256: { constexpr int MM = 256; CALL; break; }   \
        case 512: { constexpr int MM = 512; CALL; break; }   \
        case 1024: { constexpr int MM = 1024; CALL; break; } \
        case 2048: { constexpr int MM = 2048; CALL; break; } \
        default: return -1;                        \
    }

// Shared-memory wavefronts of the exchange of every pass, counted from the addresses the device code itself produces:
// each "thread" scatters a tag (its id and register number) through fft_scatter / reads through fft_gather_sw, and the
// 8-byte accesses of each half-warp are grouped by bank pair (16 of them).  out[2*pass] = scatter wavefronts,
// out[2*pass+1] = gather wavefronts, both divided by the conflict-free count (1.0 = no conflicts).
template <int M, int PASS = 0, int PS = 1>
static void emu_conflicts(double* out) {
    constexpr int TPF = M / kPts;
    constexpr int R = pass_radix(M, PASS);
    if constexpr (PS * R < M) {
        std::vector<float2> s(M);
        long long wf_s = 0, wf_g = 0, ideal = 0;
        for (int w0 = 0; w0 < TPF; w0 += 16) {                     // one half-warp at a time
            const int nt = TPF - w0 < 16 ? TPF - w0 : 16;
            for (int reg = 0; reg < kPts; ++reg) {
                int cnt_s[16] = {0}, cnt_g[16] = {0};
                for (int l = 0; l < nt; ++l) {
                    const int t = w0 + l;
                    // scatter: find where register `reg` of thread t lands by scattering a one-hot tag
                    float2 v[kPts];
                    for (int j = 0; j < kPts; ++j) v[j] = make_float2(j == reg ? 1.f : 0.f, 0.f);
                    for (auto& e : s) e = make_float2(0.f, 0.f);
                    fft_scatter<M, R, PS>(v, t, s.data());
                    for (int e = 0; e < M; ++e) if (s[e].x == 1.f) cnt_s[e & 15]++;
                    cnt_g[fft_sw(t + reg * TPF) & 15]++;
                }
                int ms = 0, mg = 0;
                for (int b = 0; b < 16; ++b) { ms = cnt_s[b] > ms ? cnt_s[b] : ms; mg = cnt_g[b] > mg ? cnt_g[b] : mg; }
                wf_s += ms; wf_g += mg; ideal += 1;
            }
        }
        out[2 * PASS] = (double) wf_s / (double) ideal;
        out[2 * PASS + 1] = (double) wf_g / (double) ideal;
        emu_conflicts<M, PASS + 1, PS * R>(out);
    }
}

extern "C" {
int emu_exchange_conflicts(int M, double* out) {
    DISPATCH(M, emu_conflicts<MM>(out));
    return 0;
}
int emu_real_forward(int M, const float* x, int len, float* packed) {
    std::vector<float2> W;
    make_w(M, W);
    DISPATCH(M, fwd<MM>(x, len, W.data(), (float2*) packed));
    return 0;
}
int emu_real_inverse(int M, const float* packed, float* y) {
    std::vector<float2> W;
    make_w(M, W);
    DISPATCH(M, inv<MM>((const float2*) packed, W.data(), y));
    return 0;
}
// acc += x * h on packed spectra (bin 0 = two real bins)
void emu_mac(int M, float* acc, const float* x, const float* h) {
    float2* a = (float2*) acc; const float2* xx = (const float2*) x; const float2* hh = (const float2*) h;
    cmac_packed0(a[0], xx[0], hh[0]);
    for (int k = 1; k < M; ++k) cmac(a[k], xx[k], hh[k]);
}
}
