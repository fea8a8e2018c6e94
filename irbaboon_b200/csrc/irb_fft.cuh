// irb_fft.cuh -- block real FFT / IFFT building blocks for the partitioned-convolution engine.
//
// Replaces the third-party transform the reference calls at fp/convolution.cpp:75-77,123,144,206 and
// Source/PluginProcessor.cpp:73-75,435,459,504 (juce::dsp::FFT::performRealOnlyForwardTransform /
// performRealOnlyInverseTransform on a zero-padded 2N-float buffer).  Instead of a full N-point complex
// transform of {x[i],0} this computes the N = 2M point REAL transform through one M-point complex
// Stockham FFT (8 points per thread, radix-8/4/2 passes through shared memory) plus a split/merge pass,
// and stores the N/2+1 useful bins PACKED into M complex values: bin 0 = {Re X[0], Re X[M]} (the
// reference's own export format, fp/ir.cpp:123-124).
//
// Everything here is written against an explicit thread index and plain pointers so that the same code
// compiles for the host (tests/emu) where "threads" are loop iterations and barriers are loop boundaries.
#pragma once

#if defined(__CUDACC__)
#define IRB_HD __host__ __device__ __forceinline__
#define IRB_CX __host__ __device__ constexpr
#else
#define IRB_HD inline
#define IRB_CX constexpr
#include <cmath>
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#endif

namespace irb {

constexpr int kPts = 8;                      // complex points held by one thread

// radix of pass `idx` of the M-point transform: 8 while at least 8 remain, then the remainder (4 or 2)
IRB_CX int pass_radix(int M, int idx) {
    int rem = M;
    for (int i = 0; i < idx; ++i) rem /= 8;
    return rem >= 8 ? 8 : rem;
}

IRB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
IRB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
IRB_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
IRB_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> IRB_HD float2 rot90(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
// multiply by exp(-+ i pi/4) and exp(-+ 3i pi/4)
template <bool INV> IRB_HD float2 rot45(float2 a) {
    const float h = 0.70710678118654752440f;
    return INV ? make_float2(h * (a.x - a.y), h * (a.x + a.y)) : make_float2(h * (a.x + a.y), h * (a.y - a.x));
}
template <bool INV> IRB_HD float2 rot135(float2 a) {
    const float h = 0.70710678118654752440f;
    return INV ? make_float2(-h * (a.x + a.y), h * (a.x - a.y)) : make_float2(h * (a.y - a.x), -h * (a.x + a.y));
}

IRB_HD void bfly2(float2& a, float2& b) { float2 t = a; a = cadd(t, b); b = csub(t, b); }

// natural-order in-place DFTs on registers a[0], a[S], a[2S] ... (S = register stride)
template <bool INV, int S> IRB_HD void dft2(float2* a) { bfly2(a[0], a[S]); }
template <bool INV, int S> IRB_HD void dft4(float2* a) {
    bfly2(a[0], a[2 * S]); bfly2(a[S], a[3 * S]);
    a[3 * S] = rot90<INV>(a[3 * S]);
    bfly2(a[0], a[S]); bfly2(a[2 * S], a[3 * S]);
    float2 t = a[S]; a[S] = a[2 * S]; a[2 * S] = t;          // bit-reversed -> natural
}
template <bool INV, int S> IRB_HD void dft8(float2* a) {
    bfly2(a[0], a[4 * S]); bfly2(a[S], a[5 * S]); bfly2(a[2 * S], a[6 * S]); bfly2(a[3 * S], a[7 * S]);
    a[5 * S] = rot45<INV>(a[5 * S]); a[6 * S] = rot90<INV>(a[6 * S]); a[7 * S] = rot135<INV>(a[7 * S]);
    bfly2(a[0], a[2 * S]); bfly2(a[S], a[3 * S]); a[3 * S] = rot90<INV>(a[3 * S]);
    bfly2(a[0], a[S]); bfly2(a[2 * S], a[3 * S]);
    bfly2(a[4 * S], a[6 * S]); bfly2(a[5 * S], a[7 * S]); a[7 * S] = rot90<INV>(a[7 * S]);
    bfly2(a[4 * S], a[5 * S]); bfly2(a[6 * S], a[7 * S]);
    // outputs sit bit-reversed: X0=a0 X4=a1 X2=a2 X6=a3 X1=a4 X5=a5 X3=a6 X7=a7
    float2 t;
    t = a[S]; a[S] = a[4 * S]; a[4 * S] = t;
    t = a[3 * S]; a[3 * S] = a[6 * S]; a[6 * S] = t;
}

// The table W of an M-point transform: first the N = 2M roots W[k] = exp(-2 pi i k / N) (split / merge of the real
// transform), then, for every pass after the first, the twiddles of that pass laid out PER THREAD: entry
// [slot*TPF + t], slot = b*(R-1) + r-1, is the factor of register b + r*NB of thread t.  Consecutive threads read
// consecutive entries, so a warp's twiddle load is one contiguous 256-byte line instead of up to 28 scattered ones
// (the scattered lookups into the root table were what kept the block FFT kernels busy: ncu, profiles/).
template <bool INV> IRB_HD float2 root(const float2* __restrict__ W, int k) {
#if defined(__CUDA_ARCH__)
    float2 w = __ldg(W + k);
#else
    float2 w = W[k];
#endif
    return INV ? cconj(w) : w;
}
IRB_CX int fft_pass_slots() { return kPts - 1; }                 // entries per thread and pass (radix 8: 7; 4: 2*3; 2: 4*1)
IRB_CX int fft_num_passes(int M) { int n = 0; for (int rem = M; rem > 1; rem /= (rem >= 8 ? 8 : rem)) ++n; return n; }
IRB_CX int fft_table_size(int M) { return 2 * M + (fft_num_passes(M) - 1) * fft_pass_slots() * (M / kPts); }
// host: fill W[fft_table_size(M)], every value computed in double and rounded once (as the reference FFT's tables are)
inline void fft_build_table(int M, float2* W) {
    const double two_pi = 6.283185307179586476925286766559;
    const int N = 2 * M, TPF = M / kPts;
    for (int k = 0; k < N; ++k) {
        const double a = -two_pi * (double) k / (double) N;
        W[k].x = (float) cos(a); W[k].y = (float) sin(a);
    }
    int PS = 1;
    for (int pass = 0; PS < M; ++pass) {
        const int R = pass_radix(M, pass), NB = kPts / R;
        if (pass > 0) {
            float2* T = W + N + (pass - 1) * fft_pass_slots() * TPF;
            for (int b = 0; b < NB; ++b)
                for (int r = 1; r < R; ++r)
                    for (int t = 0; t < TPF; ++t) {
                        const int k = (t + b * TPF) & (PS - 1);
                        const double a = -two_pi * (double) r * (double) k / (double) (PS * R);
                        float2& w = T[(b * (R - 1) + r - 1) * TPF + t];
                        w.x = (float) cos(a); w.y = (float) sin(a);
                    }
        }
        PS *= R;
    }
}

// One Stockham pass on the 8 registers of thread t (of M/8 per row).  The thread holds v[j] = x[t + j*M/8].
// With NB = 8/R butterflies per thread, butterfly b works on v[b + r*NB], r < R; its index is i = t + b*M/8,
// k = i mod PS, and after the pass its r-th output is element (i-k)*R + k + r*PS of the next array.
template <int M, int R, int PS, bool INV, int PASS>
IRB_HD void fft_pass(float2* v, int t, const float2* __restrict__ W) {
    constexpr int NB = kPts / R;
    constexpr int TPF = M / kPts;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (PS > 1) {
            const float2* T = W + 2 * M + (PASS - 1) * fft_pass_slots() * TPF + t;
#pragma unroll
            for (int r = 1; r < R; ++r) v[b + r * NB] = cmul(v[b + r * NB], root<INV>(T, (b * (R - 1) + r - 1) * TPF));
        }
        if (R == 8) dft8<INV, NB>(v + b);
        else if (R == 4) dft4<INV, NB>(v + b);
        else dft2<INV, NB>(v + b);
    }
}
// Bank swizzle of the exchange buffer.  A pass scatters with stride R (pass 0: element 8t + r), which on 8-byte elements
// puts the 16 threads of a half-warp on 2 (pass 0) or 8 (pass 1) of the 16 bank pairs.  XOR-ing the low four index bits
// with bits 4..6 and bit 6 of the index keeps aligned groups of 16 together (so the gather of 16 consecutive elements
// stays conflict-free) and spreads every scatter of every pass over all 16 bank pairs (checked exhaustively for
// M = 128 .. 2048 by tests/test_emu_fft.py::test_exchange_is_bank_conflict_free).
IRB_HD int fft_sw(int e) { return e ^ (((e >> 4) & 7) | (((e >> 6) & 1) << 3)); }

template <int M, int R, int PS>
IRB_HD void fft_scatter(const float2* v, int t, float2* srow) {
    constexpr int NB = kPts / R;
    constexpr int TPF = M / kPts;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int i = t + b * TPF;
        const int k = i & (PS - 1);
        const int base = (i - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) srow[fft_sw(base + r * PS)] = v[b + r * NB];
    }
}
// natural-order row -> registers (the layout a caller stages), and the same from the swizzled exchange buffer
template <int M>
IRB_HD void fft_gather(float2* v, int t, const float2* srow) {
#pragma unroll
    for (int j = 0; j < kPts; ++j) v[j] = srow[t + j * (M / kPts)];
}
template <int M>
IRB_HD void fft_gather_sw(float2* v, int t, const float2* srow) {
#pragma unroll
    for (int j = 0; j < kPts; ++j) v[j] = srow[fft_sw(t + j * (M / kPts))];
}

// ---- real <-> packed-half-complex split / merge ---------------------------------------------------
// Forward: Z = FFT_M(z), z[n] = x[2n] + i x[2n+1].  X[k] for one k in [1, M) from Z[k] and Z[M-k];
// k == 0 gives the packed pair {X[0], X[M]}.
IRB_HD float2 real_split(float2 zk, float2 zmk, float2 wk, int k) {
    if (k == 0) return make_float2(zk.x + zk.y, zk.x - zk.y);
    const float2 e = make_float2(0.5f * (zk.x + zmk.x), 0.5f * (zk.y - zmk.y));       // (Z[k] + conj Z[M-k]) / 2
    const float2 d = make_float2(0.5f * (zk.x - zmk.x), 0.5f * (zk.y + zmk.y));       // (Z[k] - conj Z[M-k]) / 2
    const float2 o = make_float2(d.y, -d.x);                                           // d / i
    return cadd(e, cmul(wk, o));
}
// Inverse: Z[k] = (X[k] + conj X[M-k]) + i conj(W[k]) (X[k] - conj X[M-k]); the caller scales by 1/N at the end.
// k == 0 takes the packed pair.
IRB_HD float2 real_merge(float2 xk, float2 xmk, float2 wk, int k) {
    if (k == 0) return make_float2(xk.x + xk.y, xk.x - xk.y);
    const float2 e = make_float2(xk.x + xmk.x, xk.y - xmk.y);
    const float2 d = make_float2(xk.x - xmk.x, xk.y + xmk.y);
    const float2 o = cmul(cconj(wk), d);
    return make_float2(e.x - o.y, e.y + o.x);
}

// per-bin multiply-accumulate of two packed spectra values; bin 0 holds two real bins (DC, Nyquist).
// Follows tools::complexMul (fp/tools.cpp:44-52) + the accumulate at fp/convolution.cpp:193-195.
IRB_HD void cmac(float2& acc, float2 x, float2 h) {
    acc.x = fmaf(x.x, h.x, fmaf(-x.y, h.y, acc.x));
    acc.y = fmaf(x.y, h.x, fmaf(x.x, h.y, acc.y));
}
IRB_HD void cmac_packed0(float2& acc, float2 x, float2 h) {
    acc.x = fmaf(x.x, h.x, acc.x);
    acc.y = fmaf(x.y, h.y, acc.y);
}

}  // namespace irb
