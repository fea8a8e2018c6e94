// irb_spectral.cuh -- sm_100a kernels of the single large real FFT path: fp::convolution::convolveNonPeriodic
// (fp/convolution.cpp:246-347), deconvolve (:351-403), averagingFilter (:406-546), tools::fftTransform /
// fftInvTransform (fp/tools.cpp:321-369).  These replace the third-party juce::dsp::FFT calls at
// fp/convolution.cpp:286-288,308,315,336 and fp/tools.cpp:331-335,359-363 for N = 2^4 ... 2^21.
//
// An N-point real transform is one M = N/2 point complex FFT of z[n] = x[2n] + i x[2n+1] plus a split pass.
// M <= 2048 is one shared-memory Stockham FFT per line.  Larger M = M1*M2 is the four-step scheme in two
// kernels launches of the same kernel: M2 column FFTs of length M1 (strided lines, tiled so every global access
// is a >= 64-byte segment) with the inter-pass twiddle exp(-2 pi i n2 k1 / M) computed in double, then M1 row
// FFTs of length M2 written back in NATURAL bin order, so every per-bin stage is a plain elementwise kernel.
#pragma once
#include "irb_kernels.cuh"

namespace irb {

template <int L> struct LineTile {
    static constexpr int TPF = L / kPts;                                   // threads per line
    static constexpr int G = kThreads / TPF;                               // lines transformed concurrently
    static constexpr int C = L >= 2048 ? 4 : (G > 8 ? G : 8);              // lines per CTA tile
    static constexpr int PITCH = L + 2;                                    // float2 per line in shared memory (bank spread)
    static constexpr size_t SMEM = sizeof(float2) * (size_t) C * PITCH;
};

struct LineArgs {
    const void* in;               // complex float2, or real float when in_real_len >= 0
    float2* out;
    long long in_elem_stride, in_line_stride, in_batch_stride;      // float2 units (complex view)
    long long out_elem_stride, out_line_stride, out_batch_stride;
    int n_lines;                  // lines per batch item (grid.y = batch)
    int in_real_len;              // >= 0: input is real, this many valid floats per batch item, zero beyond
    int tw_M;                     // > 0: multiply output element k of line l by exp(-+2 pi i l k / tw_M)
    float scale;                  // applied on store
    const float2* W;              // the 2L roots of unity of the line length
};

template <int L, bool INV>
static __global__ void __launch_bounds__(kThreads) k_line_fft(const LineArgs a) {
    using T = LineTile<L>;
    extern __shared__ __align__(16) float2 s_lines[];
    const int tid = threadIdx.x;
    const int line0 = blockIdx.x * T::C;
    const long long ib = (long long) blockIdx.y * a.in_batch_stride, ob = (long long) blockIdx.y * a.out_batch_stride;
    const bool in_contig = a.in_elem_stride == 1, out_contig = a.out_elem_stride == 1;

    for (int idx = tid; idx < T::C * L; idx += kThreads) {
        const int c = in_contig ? idx / L : idx % T::C;
        const int j = in_contig ? idx % L : idx / T::C;
        const int line = line0 + c;
        float2 v = make_float2(0.f, 0.f);
        if (line < a.n_lines) {
            const long long off = (long long) line * a.in_line_stride + (long long) j * a.in_elem_stride;
            if (a.in_real_len < 0) v = reinterpret_cast<const float2*>(a.in)[ib + off];
            else {
                const float* r = reinterpret_cast<const float*>(a.in) + 2 * ib;
                const long long e = 2 * off;
                if (e + 1 < a.in_real_len) v = *reinterpret_cast<const float2*>(r + e);
                else if (e < a.in_real_len) v.x = r[e];
            }
        }
        s_lines[c * T::PITCH + j] = v;
    }
    bar_compute();
    {
        const int t = tid % T::TPF;
        for (int c = tid / T::TPF; c < T::C; c += T::G) {
            float2* srow = s_lines + c * T::PITCH;
            float2 v[kPts];
            fft_gather<L>(v, t, srow);
            fft_run<L, INV>(v, t, srow, a.W);
            const int line = line0 + c;
#pragma unroll
            for (int j = 0; j < kPts; ++j) {
                const int k = t + j * T::TPF;
                if (a.tw_M > 0) {
                    double sn, cs;
                    sincospi(2.0 * (double) ((long long) line * k % a.tw_M) / (double) a.tw_M, &sn, &cs);
                    const float2 w = make_float2((float) cs, INV ? (float) sn : (float) -sn);
                    v[j] = cmul(v[j], w);
                }
                srow[k] = v[j];              // each thread rewrites exactly the elements it gathered last
            }
        }
    }
    bar_compute();
    for (int idx = tid; idx < T::C * L; idx += kThreads) {
        const int c = out_contig ? idx / L : idx % T::C;
        const int j = out_contig ? idx % L : idx / T::C;
        const int line = line0 + c;
        if (line < a.n_lines) {
            float2 v = s_lines[c * T::PITCH + j];
            v.x *= a.scale; v.y *= a.scale;
            a.out[ob + (long long) line * a.out_line_stride + (long long) j * a.out_elem_stride] = v;
        }
    }
}

// ---- elementwise stages on natural-order arrays; grid.y = batch item ------------------------------------
// W2M: the N = 2M roots exp(-2 pi i k / N) (table for M <= 2048, else computed in double)
__device__ __forceinline__ float2 root_N(const float2* __restrict__ W, int M, int k) {
    if (W) return __ldg(W + k);
    double sn, cs;
    sincospi((double) k / (double) M, &sn, &cs);
    return make_float2((float) cs, (float) -sn);
}

// Z (M complex) -> interleaved spectrum S: bins 0..M, and with `mirror` the conjugate bins M+1..2M-1 as the
// full complex transform of a real signal has them (2N floats per channel, N = 2M)
static __global__ void k_spec_split(const float2* Z, long long z_stride, float2* S, long long s_stride, int M, int mirror, const float2* W) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > M) return;
    const float2* z = Z + blockIdx.y * z_stride;
    float2* s = S + blockIdx.y * s_stride;
    float2 x;
    if (k == 0) x = make_float2(z[0].x + z[0].y, 0.f);
    else if (k == M) x = make_float2(z[0].x - z[0].y, 0.f);
    else x = real_split(z[k], z[M - k], root_N(W, M, k), k);
    s[k] = x;
    if (mirror && k > 0 && k < M) s[2 * M - k] = make_float2(x.x, -x.y);
}
// interleaved spectrum S (bins 0..M read) -> Z' (M complex) ready for the inverse complex FFT
static __global__ void k_spec_merge(const float2* S, long long s_stride, float2* Z, long long z_stride, int M, const float2* W) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const float2* s = S + blockIdx.y * s_stride;
    float2* z = Z + blockIdx.y * z_stride;
    if (k == 0) {
        // the reference's inverse reads bins 0 and M as complex values; their imaginary parts enter the rebuilt
        // spectrum but cancel in a real output except through x[0] and the alternating term: keep the real parts
        z[0] = make_float2(s[0].x + s[M].x, s[0].x - s[M].x);
    } else z[k] = real_merge(s[k], s[M - k], root_N(W, M, k), k);
}
// tools::complexMul (fp/tools.cpp:44-52) / complexDivCartesian (:72-86) per bin
__device__ __forceinline__ float2 bin_mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.y * b.x + a.x * b.y); }
__device__ __forceinline__ float2 bin_div(float2 a, float2 b) {
    if (b.x == 0.0f && b.y == 0.0f) return a;
    const float den = b.x * b.x + b.y * b.y;
    return make_float2((a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den);
}
// S[k] = op(A[k], B[k]) on bins 0..M of interleaved spectra (B broadcast over the batch when b_stride == 0)
template <bool DIV>
static __global__ void k_spec_binop(float2* A, long long a_stride, const float2* Bs, long long b_stride, int M) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > M) return;
    float2* pa = A + blockIdx.y * a_stride;
    const float2 b = Bs[blockIdx.y * b_stride + k];
    pa[k] = DIV ? bin_div(pa[k], b) : bin_mul(pa[k], b);
}
// fused split -> per-bin op -> merge on the pair (k, M-k): Za op Zb -> Z' without materialising the spectra
template <bool DIV>
static __global__ void k_spec_fused(const float2* Za, long long a_stride, const float2* Zb, long long b_stride, float2* Zo, long long o_stride, int M,
                             const float2* W) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > M / 2) return;
    const float2* za = Za + blockIdx.y * a_stride;
    const float2* zb = Zb + blockIdx.y * b_stride;
    float2* zo = Zo + blockIdx.y * o_stride;
    if (k == 0) {
        const float2 a0 = make_float2(za[0].x + za[0].y, 0.f), aM = make_float2(za[0].x - za[0].y, 0.f);
        const float2 b0 = make_float2(zb[0].x + zb[0].y, 0.f), bM = make_float2(zb[0].x - zb[0].y, 0.f);
        const float2 q0 = DIV ? bin_div(a0, b0) : bin_mul(a0, b0), qM = DIV ? bin_div(aM, bM) : bin_mul(aM, bM);
        zo[0] = make_float2(q0.x + qM.x, q0.x - qM.x);
        return;
    }
    const float2 w = root_N(W, M, k), wm = root_N(W, M, M - k);
    const float2 ak = za[k], am = za[M - k], bk = zb[k], bm = zb[M - k];
    const float2 A = real_split(ak, am, w, k), Am = real_split(am, ak, wm, M - k);
    const float2 B = real_split(bk, bm, w, k), Bm = real_split(bm, bk, wm, M - k);
    const float2 Q = DIV ? bin_div(A, B) : bin_mul(A, B), Qm = DIV ? bin_div(Am, Bm) : bin_mul(Am, Bm);
    zo[k] = real_merge(Q, Qm, w, k);
    if (2 * k != M) zo[M - k] = real_merge(Qm, Q, wm, M - k);
}

// tools::fftTransform(formatAmplPhase = true): bins 0..M -> {amplitude, phase} (fp/tools.cpp:337-342,222-231)
static __global__ void k_spec_ampl_phase(float2* S, long long s_stride, int M) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > M) return;
    float2* s = S + blockIdx.y * s_stride;
    const float2 v = s[k];
    s[k] = make_float2((float) sqrt((double) v.x * (double) v.x + (double) v.y * (double) v.y), atan2f(v.y, v.x));
}

// ---- averagingFilter (fp/convolution.cpp:406-546), one pass = three kernels ------------------------------
__device__ __forceinline__ float round_1e16(float x) {                  // tools::roundTo1TenQuadrillionth, fp/tools.cpp:212-218
    if (!signbit(x) && (double) x < 1e-16) return 1e-16f;
    if (signbit(x) && (double) x > -1e-16) return -1e-16f;
    return x;
}
__device__ __forceinline__ float round_to_zero(float x, float thr) {   // tools::roundToZero, fp/tools.cpp:199-209
    if (!signbit(x) && x < thr) x = 0.0f;
    if (signbit(x) && x > -thr) x = 0.0f;
    return x;
}
// (1) per bin k <= M: la[k] = log(clamp(|S[k]|)) (or the plain amplitude for the linear average) and the window
// [lo, hi] of the bin, with the reference's double-precision operation order (fp/convolution.cpp:451-458):
// binFreq = k*freqPerBin; lo = round((binFreq / c) / freqPerBin); hi = round((binFreq * c) / freqPerBin)
static __global__ void k_avg_prepare(const float2* S, long long s_stride, float* la, long long la_stride, int* lo, int* hi, int M, int log_avg,
                              double freq_per_bin, double c_side) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > M) return;
    const float2 v = S[blockIdx.y * s_stride + k];
    const float ampl = (float) sqrt((double) v.x * (double) v.x + (double) v.y * (double) v.y);     // tools::binAmpl
    la[blockIdx.y * la_stride + k] = log_avg ? logf(round_1e16(ampl)) : ampl;
    if (blockIdx.y == 0) {
        const double f = (double) k * freq_per_bin;
        lo[k] = (int) round((f / c_side) / freq_per_bin);
        hi[k] = (int) round((f * c_side) / freq_per_bin);
    }
}
// (2) the running window sum in the reference's exact sequential float order (fp/convolution.cpp:482-505): one warp
// per spectrum; lane 0 carries the dependent add chain while all lanes stage la[] and the window edges through
// shared memory in coalesced chunks.  Bins past Nyquist (k > M) hold amplitude 0 -> clamp -> log(1e-16) and
// count up to bin 2M-1 (the "addBin < fftSize" test).  rs[k] receives the raw running sum.
static __global__ void __launch_bounds__(32) k_avg_scan(const float* la, long long la_stride, const int* lo, const int* hi, float* rs, long long rs_stride,
                                                 int M) {
    constexpr int CH = 1024;
    __shared__ float s_sub[CH], s_add[CH], s_out[CH];
    __shared__ int s_lo[CH], s_hi[CH];
    const float* a = la + blockIdx.x * la_stride;
    float* out = rs + blockIdx.x * rs_stride;
    const int lane = threadIdx.x;
    const float log_floor = logf(1e-16f);
    float running = 0.0f;
    int prevLo = 0, prevHi = -1;
    int sub_base = 0, add_base = 0;
    auto stage = [&](float* dst, int base) {
        __syncwarp();
        for (int i = lane; i < CH; i += 32) { const int k = base + i; dst[i] = k <= M ? a[k] : log_floor; }
        __syncwarp();
    };
    stage(s_sub, 0);
    stage(s_add, 0);
    for (int k0 = 0; k0 <= M; k0 += CH) {
        __syncwarp();
        for (int i = lane; i < CH; i += 32) { const int k = k0 + i; s_lo[i] = k <= M ? lo[k] : 0; s_hi[i] = k <= M ? hi[k] : 0; }
        __syncwarp();
        const int kend = min(CH, M + 1 - k0);
        for (int i = 0; i < kend; ++i) {
            const int l = s_lo[i], h = s_hi[i];
            for (int b = prevLo; b < l; ++b) {
                if (b - sub_base >= CH) { sub_base = b; stage(s_sub, sub_base); }
                if (lane == 0) running -= s_sub[b - sub_base];
            }
            for (int b = prevHi + 1; b <= h; ++b) {
                if (b - add_base >= CH) { add_base = b; stage(s_add, add_base); }
                if (b < 2 * M && lane == 0) running += s_add[b - add_base];
            }
            if (lane == 0) s_out[i] = running;
            prevLo = l; prevHi = h;
        }
        __syncwarp();
        for (int i = lane; i < kend; i += 32) out[k0 + i] = s_out[i];
    }
}
// linear average: a fresh ascending sum per bin (fp/convolution.cpp:508-514) -- no sequential dependence
static __global__ void k_avg_linear_sum(const float* la, long long la_stride, const int* lo, const int* hi, float* rs, long long rs_stride, int M) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > M) return;
    const float* a = la + blockIdx.y * la_stride;
    float sum = 0.0f;
    for (int b = lo[k]; b <= hi[k]; ++b) sum += (b <= M ? a[b] : 0.0f);
    rs[blockIdx.y * rs_stride + k] = sum;
}
// (3) new amplitude = exp(sum / window length), bins rebuilt from it and the original phase (fp/convolution.cpp:518-543)
static __global__ void k_avg_apply(float2* S, long long s_stride, const float* rs, long long rs_stride, const int* lo, const int* hi, int M, int log_avg,
                            int include_phase, int include_ampl) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > M) return;
    float2* s = S + blockIdx.y * s_stride;
    const double len = (double) (hi[k] - lo[k]) + 1.0;
    float ampl = (float) ((double) rs[blockIdx.y * rs_stride + k] / len);
    if (log_avg) ampl = expf(ampl);
    ampl = round_1e16(ampl);
    const float re = round_to_zero(s[k].x, 1e-11f), im = round_to_zero(s[k].y, 1e-11f);
    float phase = atan2f(im, re);
    if (!include_ampl) ampl = 1.0f;
    if (!include_phase) phase = 0.0f;
    s[k] = make_float2(ampl * cosf(phase), ampl * sinf(phase));
}

// out[i] = in[(i + h1) mod n] : ir::shifteroo (fp/ir.cpp:85-103), h1 = ceil(n/2)
static __global__ void k_shifteroo(const float* in, float* out, int n, long long stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int h2 = n / 2, h1 = n - h2;
    out[blockIdx.y * stride + i] = in[blockIdx.y * stride + (i < h2 ? i + h1 : i - h2)];
}

// fp::ExpSineSweep::generate / generateInv (fp/ExpSineSweep.cpp:26-41,59-79), FP64:
// sweep[i] = g sin(K (exp(i / L) - 1)); inverse = reversed sweep times k^(i+1)
static __global__ void k_ess(double* out, int n, double g, double K, double L, int inverse, double kdecay) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = inverse ? n - 1 - i : i;
    double v = g * sin(K * (exp((double) j / L) - 1.0));
    if (inverse) v *= pow(kdecay, (double) (i + 1));
    out[i] = v;
}

}  // namespace irb
