/*
 * oracle/irb_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of IRBaboon's partitioned-convolution hot path, used as the
 * checker for the CUDA implementation (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).
 * The product (irbaboon_b200/) never includes, links or calls anything in this file.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Parity status: the reference ships no tests or golden vectors.  This restatement is pinned against the
 * reference's OWN object code (oracle/_ref/libirb_ref.so = fp/*.cpp compiled unmodified) in
 * tests/test_oracle.py, and against golden vectors generated from that build (tests/golden/).  The one
 * piece that cannot be pinned is the third-party FFT: JUCE 6.0.1 juce_dsp (dsp::FFT) is not vendored by the
 * reference; orc_fft_* restates the published behaviour of its built-in fallback engine (mixed radix-4/2
 * decimation in time, float32 butterflies, double-generated twiddles, inverse scaled by 1/N) and is
 * cross-checked against a float64 FFT.
 *
 * Buffers are planar, channel-contiguous float arrays: x[ch * L + i].
 * Compile with -ffp-contract=off so the arithmetic matches the reference's non-FMA x86-64 build.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------------
 * FFT: juce::dsp::FFT contract as used at convolution.cpp:75-77,123,144,206; tools.cpp:331-335,359-363;
 * PluginProcessor.cpp:73-75,435,459,504.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int n, nstages;
    int radix[32], length[32];
    float *twf, *twi;            /* interleaved {cos, sin}; forward uses -2*pi/n, inverse +2*pi/n */
} orc_fft;

static void orc_fft_init(orc_fft* f, int n) {
    f->n = n;
    f->twf = (float*) malloc(sizeof(float) * 2 * (size_t) n);
    f->twi = (float*) malloc(sizeof(float) * 2 * (size_t) n);
    for (int i = 0; i < n; ++i) {
        double pf = -2.0 * M_PI / (double) n, pi_ = 2.0 * M_PI / (double) n;
        f->twf[2 * i] = (float) cos(i * pf);  f->twf[2 * i + 1] = (float) sin(i * pf);
        f->twi[2 * i] = (float) cos(i * pi_); f->twi[2 * i + 1] = (float) sin(i * pi_);
    }
    f->nstages = 0;
    int rem = n;
    while (rem > 1) {
        int r = (rem % 4 == 0) ? 4 : 2;
        rem /= r;
        f->radix[f->nstages] = r;
        f->length[f->nstages] = rem;
        f->nstages++;
    }
}
static void orc_fft_free(orc_fft* f) { free(f->twf); free(f->twi); }

static void orc_bfly2(const float* tw, float* d, int stride, int len) {
    for (int i = 0; i < len; ++i) {
        float wr = tw[2 * (size_t) i * stride], wi = tw[2 * (size_t) i * stride + 1];
        float* a = d + 2 * i; float* b = d + 2 * (i + len);
        float sr = b[0] * wr - b[1] * wi, si = b[0] * wi + b[1] * wr;
        b[0] = a[0] - sr; b[1] = a[1] - si;
        a[0] += sr; a[1] += si;
    }
}
static void orc_bfly4(const float* tw, float* d, int stride, int len, int inverse) {
    for (int i = 0; i < len; ++i) {
        const float* w1 = tw + 2 * (size_t) i * stride;
        const float* w2 = tw + 2 * (size_t) i * stride * 2;
        const float* w3 = tw + 2 * (size_t) i * stride * 3;
        float* d0 = d + 2 * i; float* d1 = d + 2 * (i + len); float* d2 = d + 2 * (i + 2 * len); float* d3 = d + 2 * (i + 3 * len);
        float s0r = d1[0] * w1[0] - d1[1] * w1[1], s0i = d1[0] * w1[1] + d1[1] * w1[0];
        float s1r = d2[0] * w2[0] - d2[1] * w2[1], s1i = d2[0] * w2[1] + d2[1] * w2[0];
        float s2r = d3[0] * w3[0] - d3[1] * w3[1], s2i = d3[0] * w3[1] + d3[1] * w3[0];
        float s3r = s0r + s2r, s3i = s0i + s2i;
        float s4r = s0r - s2r, s4i = s0i - s2i;
        float s5r = d0[0] - s1r, s5i = d0[1] - s1i;
        d0[0] += s1r; d0[1] += s1i;
        d2[0] = d0[0] - s3r; d2[1] = d0[1] - s3i;
        d0[0] += s3r; d0[1] += s3i;
        if (inverse) {
            d1[0] = s5r - s4i; d1[1] = s5i + s4r;
            d3[0] = s5r + s4i; d3[1] = s5i - s4r;
        } else {
            d1[0] = s5r + s4i; d1[1] = s5i - s4r;
            d3[0] = s5r - s4i; d3[1] = s5i + s4r;
        }
    }
}
static void orc_fft_rec(const orc_fft* f, int inverse, const float* in, float* out, int stride, int level) {
    int r = f->radix[level], len = f->length[level];
    if (len == 1) {
        for (int i = 0; i < r; ++i) { out[2 * i] = in[2 * (size_t) i * stride]; out[2 * i + 1] = in[2 * (size_t) i * stride + 1]; }
    } else {
        for (int i = 0; i < r; ++i) orc_fft_rec(f, inverse, in + 2 * (size_t) i * stride, out + 2 * (size_t) i * len, stride * r, level + 1);
    }
    const float* tw = inverse ? f->twi : f->twf;
    if (r == 2) orc_bfly2(tw, out, stride, len); else orc_bfly4(tw, out, stride, len, inverse);
}
/* performRealOnlyForwardTransform(d, true): n reals in a 2n-float buffer -> n interleaved complex bins */
static void orc_fft_real_forward(const orc_fft* f, float* d) {
    int n = f->n;
    if (n == 1) return;
    float* s = (float*) malloc(sizeof(float) * 2 * (size_t) n);
    for (int i = 0; i < n; ++i) { s[2 * i] = d[i]; s[2 * i + 1] = 0.0f; }
    orc_fft_rec(f, 0, s, d, 1, 0);
    free(s);
}
/* performRealOnlyInverseTransform(d): reads bins 0..n/2, rebuilds the rest, n reals (scaled 1/n) to d[0..n) */
static void orc_fft_real_inverse(const orc_fft* f, float* d) {
    int n = f->n;
    if (n == 1) return;
    for (int i = n >> 1; i < n; ++i) { d[2 * i] = d[2 * (n - i)]; d[2 * i + 1] = -d[2 * (n - i) + 1]; }
    float* s = (float*) malloc(sizeof(float) * 2 * (size_t) n);
    orc_fft_rec(f, 1, d, s, 1, 0);
    const float scale = 1.0f / (float) n;
    for (int i = 0; i < 2 * n; ++i) s[i] *= scale;
    for (int i = 0; i < n; ++i) { d[i] = s[2 * i]; d[i + n] = s[2 * i + 1]; }
    free(s);
}

/* exported raw transforms (for kernel-level parity tests): buf is 2n floats, in place */
void orc_real_forward(float* buf, int n) { orc_fft f; orc_fft_init(&f, n); orc_fft_real_forward(&f, buf); orc_fft_free(&f); }
void orc_real_inverse(float* buf, int n) { orc_fft f; orc_fft_init(&f, n); orc_fft_real_inverse(&f, buf); orc_fft_free(&f); }

/* ------------------------------------------------------------------------------------------------
 * tools.cpp scalar primitives
 * ---------------------------------------------------------------------------------------------- */
/* tools.cpp:44-52 */
void orc_complex_mul(float* a, float* b, float c, float d) {
    float re = (*a) * c - (*b) * d;
    float im = (*b) * c + (*a) * d;
    *a = re; *b = im;
}
/* tools.cpp:72-86 (denominator 0+0i: leave the numerator untouched) */
void orc_complex_div_cartesian(float* a, float* b, float c, float d) {
    if (c == 0.0 && d == 0.0) return;
    float re = ((*a) * c + (*b) * d) / (c * c + d * d);
    float im = ((*b) * c - (*a) * d) / (c * c + d * d);
    *a = re; *b = im;
}
/* tools.cpp:199-209 */
static void orc_round_to_zero(float* x, float thr) {
    if (thr < 0.0f) return;
    if (!signbit(*x) && *x < thr) *x = 0.0f;
    if (signbit(*x) && *x > -thr) *x = 0.0f;
}
/* tools.cpp:213-218 (comparison against the double literal 1e-16) */
static void orc_round_1e16(float* x) {
    if (!signbit(*x) && (double) *x < 1e-16) *x = (float) 1e-16;
    if (signbit(*x) && (double) *x > -1e-16) *x = (float) -1e-16;
}
/* tools.cpp:222-231: amplitude through double pow/sqrt, phase through float atan2 */
static float orc_bin_ampl(const float* bin) { return (float) sqrt(pow((double) bin[0], 2.0) + pow((double) bin[1], 2.0)); }
static float orc_bin_phase(const float* bin) { return atan2f(bin[1], bin[0]); }
/* tools.cpp:184-196 */
int orc_next_pow2(int x) {
    if (x != 0 && (x & (x - 1)) == 0) return x;
    int r = 1;
    while (r <= x) r *= 2;
    return r;
}
/* tools.cpp:14-30 */
static void orc_sum_to_mono(float* p0, float* p1, int n) {
    for (int i = 0; i < n; ++i) { p0[i] += p1[i]; p0[i] /= 2.0f; p1[i] = 0.0f; }
}

/* ------------------------------------------------------------------------------------------------
 * convolvePeriodic -- convolution.cpp:14-242 (offline uniformly partitioned overlap-add)
 * out: chx * (Lx+Lh-1) floats.  Returns Lx+Lh-1.
 * ---------------------------------------------------------------------------------------------- */
enum { ORC_UNKNOWN, ORC_IRM_AM, ORC_IRM_AS, ORC_IRS_AM, ORC_IRS_AS };   /* convolution.hpp:18-24 */

static int orc_layout(int chIR, int chAudio) {                              /* convolution.cpp:28-37 */
    if (chIR == 1 && chAudio == 1) return ORC_IRM_AM;
    if (chIR == 1 && chAudio == 2) return ORC_IRM_AS;
    if (chIR == 2 && chAudio == 1) return ORC_IRS_AM;
    if (chIR == 2 && chAudio == 2) return ORC_IRS_AS;
    return ORC_UNKNOWN;
}

int orc_convolve_periodic(const float* x, int chx, int Lx, const float* h, int chh, int Lh, int B, float* out) {
    const int Lo = Lx + Lh - 1;
    memset(out, 0, sizeof(float) * (size_t) chx * (size_t) Lo);                      /* :20-21 */
    const int layout = orc_layout(chh, chx);
    if (layout == ORC_UNKNOWN) return Lo;                                            /* :39-42 */

    int N = 1;
    while (N < 2 * B - 1) N *= 2;                                                    /* :45-48 */
    const int fbs = 2 * N;                                                           /* :49 */
    const int P = (int) ceilf((float) Lh / (float) B);                               /* :52 */

    float* irFft = (float*) calloc((size_t) P * chh * fbs, sizeof(float));           /* :55-60 */
    float* auFft = (float*) calloc((size_t) P * chx * fbs, sizeof(float));           /* :63-68 */
    float* conv = (float*) calloc((size_t) chx * fbs, sizeof(float));                /* :80-81 */
    float* inplace = (float*) calloc((size_t) fbs, sizeof(float));
    float* overlap = (float*) calloc((size_t) chx * B, sizeof(float));               /* :86-87 */
    orc_fft fft; orc_fft_init(&fft, N);

    int irLoaded = 0, auIndex = 0, k = 0, endAudio = 0;
    const int irFwdChMax = (layout == ORC_IRS_AS) ? 2 : 1;                           /* :94-101 */

    do {
        if (k < P) {                                                                 /* :106-125 lazy IR partition load */
            int n = ((k + 1) * B <= Lh) ? B : Lh - k * B;
            for (int c = 0; c < chh; ++c) memcpy(irFft + ((size_t) k * chh + c) * fbs, h + (size_t) c * Lh + (size_t) k * B, sizeof(float) * (size_t) n);
            if (layout == ORC_IRS_AM) orc_sum_to_mono(irFft + ((size_t) k * chh) * fbs, irFft + ((size_t) k * chh + 1) * fbs, fbs);
            for (int c = 0; c < irFwdChMax; ++c) orc_fft_real_forward(&fft, irFft + ((size_t) k * chh + c) * fbs);
            irLoaded++;
        }
        if (k * B <= Lx) {                                                           /* :128-149 (note <=) */
            auIndex = k % P;
            memset(auFft + (size_t) auIndex * chx * fbs, 0, sizeof(float) * (size_t) chx * fbs);
            int n = ((k + 1) * B <= Lx) ? B : Lx - k * B;
            for (int c = 0; c < chx; ++c) {
                float* dst = auFft + ((size_t) auIndex * chx + c) * fbs;
                if (n > 0) memcpy(dst, x + (size_t) c * Lx + (size_t) k * B, sizeof(float) * (size_t) n);
                orc_fft_real_forward(&fft, dst);
            }
            if ((k + 1) * B > Lx) endAudio++;
        } else {
            endAudio++;                                                              /* :151-153 tail */
        }

        memset(conv, 0, sizeof(float) * (size_t) chx * fbs);                         /* :158 */
        for (int c = 0; c < chx; ++c) {                                              /* :160-215 */
            float* cr = conv + (size_t) c * fbs;
            float* ov = overlap + (size_t) c * B;
            int p = endAudio - 1 > 0 ? endAudio - 1 : 0;                             /* :166 */
            int a = auIndex;                                                         /* :167 */
            while (p < irLoaded) {                                                   /* :171 */
                memcpy(inplace, auFft + ((size_t) a * chx + c) * fbs, sizeof(float) * (size_t) fbs);   /* :173-174 */
                int irCh = (layout == ORC_IRM_AM || layout == ORC_IRS_AS) ? c : 0;   /* :176-180 */
                const float* ir = irFft + ((size_t) p * chh + irCh) * fbs;
                for (int i = 0; i <= N; i += 2) orc_complex_mul(inplace + i, inplace + i + 1, ir[i], ir[i + 1]);   /* :184-189 */
                for (int i = 0; i < fbs; ++i) cr[i] += inplace[i];                   /* :193-195 */
                p++; a--;
                if (a < 0) a = P - 1;                                                /* :198-201 */
            }
            orc_fft_real_inverse(&fft, cr);                                          /* :206 */
            for (int i = 0; i < B; ++i) { cr[i] += ov[i]; ov[i] = cr[B + i]; }       /* :210-213 */
        }
        {
            int n = ((k + 1) * B <= Lo) ? B : Lo - k * B;                            /* :219-230 */
            for (int c = 0; c < chx; ++c) if (n > 0) memcpy(out + (size_t) c * Lo + (size_t) k * B, conv + (size_t) c * fbs, sizeof(float) * (size_t) n);
        }
        k++;
    } while (endAudio < P);                                                          /* :233 -- last overlap never flushed */

    orc_fft_free(&fft);
    free(irFft); free(auFft); free(conv); free(inplace); free(overlap);
    return Lo;
}

/* number of block iterations convolvePeriodic performs (convolution.cpp:104-233): floor(Lx/B) + P */
int orc_periodic_iterations(int Lx, int Lh, int B) {
    int P = (int) ceilf((float) Lh / (float) B);
    return Lx / B + P;
}

/* ------------------------------------------------------------------------------------------------
 * convolveNonPeriodic -- convolution.cpp:246-347 (single FFT of length N >= Lx+Lh-1)
 * out: chx * (Lx+Lh-1) floats; unknown layout returns a cleared chx*Lx buffer (:271-275) -> returns Lx.
 * ---------------------------------------------------------------------------------------------- */
int orc_convolve_nonperiodic(const float* x, int chx, int Lx, const float* h, int chh, int Lh, float* out) {
    const int Lo = Lx + Lh - 1;
    const int layout = orc_layout(chh, chx);
    if (layout == ORC_UNKNOWN) { memset(out, 0, sizeof(float) * (size_t) chx * (size_t) Lx); return Lx; }
    int N = 1;
    while (N < Lo) N *= 2;                                                           /* :278-281 */
    const int fbs = 2 * N;
    float* a = (float*) calloc((size_t) chx * fbs, sizeof(float));                   /* :291 setSize keeps + zero pads */
    float* b = (float*) calloc((size_t) chh * fbs, sizeof(float));                   /* :293 */
    for (int c = 0; c < chx; ++c) memcpy(a + (size_t) c * fbs, x + (size_t) c * Lx, sizeof(float) * (size_t) Lx);
    for (int c = 0; c < chh; ++c) memcpy(b + (size_t) c * fbs, h + (size_t) c * Lh, sizeof(float) * (size_t) Lh);
    orc_fft fft; orc_fft_init(&fft, N);
    int irFwdChMax = (layout == ORC_IRS_AS) ? 2 : 1;
    if (layout == ORC_IRS_AM) orc_sum_to_mono(b, b + fbs, fbs);                      /* :301 */
    for (int c = 0; c < irFwdChMax; ++c) orc_fft_real_forward(&fft, b + (size_t) c * fbs);   /* :307-308 */
    for (int c = 0; c < chx; ++c) {                                                  /* :311-337 */
        float* au = a + (size_t) c * fbs;
        orc_fft_real_forward(&fft, au);
        int irCh = (layout == ORC_IRM_AM || layout == ORC_IRS_AS) ? c : 0;
        const float* ir = b + (size_t) irCh * fbs;
        for (int i = 0; i <= N; i += 2) orc_complex_mul(au + i, au + i + 1, ir[i], ir[i + 1]);
        orc_fft_real_inverse(&fft, au);
        memcpy(out + (size_t) c * Lo, au, sizeof(float) * (size_t) Lo);              /* :340-344 */
    }
    orc_fft_free(&fft);
    free(a); free(b);
    return Lo;
}

/* ------------------------------------------------------------------------------------------------
 * tools::fftTransform / fftInvTransform -- tools.cpp:321-369
 * ---------------------------------------------------------------------------------------------- */
/* out: ch * 2N floats, N = nextPow2(L).  Only channel 0 of the input is copied (tools.cpp:328). */
int orc_fft_transform(const float* x, int ch, int L, float* out) {
    int N = orc_next_pow2(L);
    int fbs = 2 * N;
    memset(out, 0, sizeof(float) * (size_t) ch * fbs);
    memcpy(out, x, sizeof(float) * (size_t) L);
    orc_fft fft; orc_fft_init(&fft, N);
    for (int c = 0; c < ch; ++c) orc_fft_real_forward(&fft, out + (size_t) c * fbs);
    orc_fft_free(&fft);
    return fbs;
}
/* in: ch * fftSize floats; out: ch * fftSize/2 floats */
int orc_fft_inv_transform(const float* x, int ch, int fftSize, float* out) {
    int N = fftSize / 2;
    float* t = (float*) malloc(sizeof(float) * (size_t) fftSize);
    orc_fft fft; orc_fft_init(&fft, N);
    for (int c = 0; c < ch; ++c) {
        memcpy(t, x + (size_t) c * fftSize, sizeof(float) * (size_t) fftSize);
        orc_fft_real_inverse(&fft, t);
        memcpy(out + (size_t) c * N, t, sizeof(float) * (size_t) N);
    }
    orc_fft_free(&fft);
    free(t);
    return N;
}

/* ------------------------------------------------------------------------------------------------
 * averagingFilter -- convolution.cpp:406-546.  buf: ch * fftSize floats (interleaved spectrum), in place.
 * NOTE the last two flags are INCLUDE flags in the body although convolution.hpp:44 names them nullify*.
 * ---------------------------------------------------------------------------------------------- */
void orc_averaging_filter(float* buf, int ch, int fftSize, double octaveFraction, double sampleRate, int logAvg,
                          int includePhase, int includeAmplitude) {
    if (!(fftSize != 0 && (fftSize & (fftSize - 1)) == 0)) return;                   /* :412-415 */
    float* oldA = (float*) calloc((size_t) ch * (fftSize / 2), sizeof(float));       /* :417-420 */
    float* newA = (float*) calloc((size_t) ch * fftSize, sizeof(float));
    int N = fftSize / 2;
    double fractPerSide = octaveFraction / 2.0;
    double nyquist = sampleRate / 2;
    double freqPerBin = nyquist / (double) (N / 2);                                  /* :425 */

    for (int c = 0; c < ch; ++c) {                                                   /* :429-436 */
        float* b = buf + (size_t) c * fftSize;
        float* o = oldA + (size_t) c * (fftSize / 2);
        for (int bin = 0; bin <= N; bin += 2) o[bin / 2] = orc_bin_ampl(b + bin);
    }
    for (int c = 0; c < ch; ++c) {                                                   /* :439-545 */
        float runningSum = 0.0f;
        int prevLower = 0, prevUpper = -2;
        float* o = oldA + (size_t) c * (fftSize / 2);
        float* nw = newA + (size_t) c * fftSize;
        for (int bin = 0; bin <= N; bin += 2) {
            double binFreq = (bin / 2) * freqPerBin;
            double lowerFreq = binFreq / pow(2.0, fractPerSide);
            int lowerBin = 2 * (int) round(lowerFreq / freqPerBin);
            double upperFreq = binFreq * pow(2.0, fractPerSide);
            int upperBin = 2 * (int) round(upperFreq / freqPerBin);
            double binRangeLength = (double) (upperBin - lowerBin) / 2.0 + 1.0;
            if (logAvg) {                                                            /* :482-505 sequential float running sum */
                int sub = prevLower;
                while (sub < lowerBin) {
                    float v = o[sub / 2];
                    orc_round_1e16(&v);
                    v = logf(v);
                    runningSum -= v;
                    sub += 2;
                }
                int add = prevUpper + 2;
                while (add <= upperBin) {
                    if (add < fftSize) {
                        float v = o[add / 2];
                        orc_round_1e16(&v);
                        v = logf(v);
                        runningSum += v;
                    }
                    add += 2;
                }
            } else {                                                                 /* :508-514 */
                runningSum = 0.0f;
                for (int i = lowerBin; i <= upperBin; i += 2) runningSum += o[i / 2];
            }
            float na = (float) ((double) runningSum / binRangeLength);               /* :518 */
            if (logAvg) na = expf(na);
            orc_round_1e16(&na);
            nw[bin] = na;
            prevLower = lowerBin;
            prevUpper = upperBin;
        }
        float* b = buf + (size_t) c * fftSize;                                       /* :530-543 */
        for (int bin = 0; bin <= N; bin += 2) {
            float ampl = nw[bin];
            orc_round_to_zero(b + bin, 1e-11f);
            orc_round_to_zero(b + bin + 1, 1e-11f);
            float phase = orc_bin_phase(b + bin);
            if (!includeAmplitude) ampl = 1.0f;
            if (!includePhase) phase = 0.0f;
            b[bin] = ampl * cosf(phase);
            b[bin + 1] = ampl * sinf(phase);
        }
    }
    free(oldA); free(newA);
}

/* ir::shifteroo -- ir.cpp:85-103 */
void orc_shifteroo(float* buf, int ch, int n) {
    if (n < 2) return;
    int h2 = n / 2, h1 = h2 + (n % 2 != 0 ? 1 : 0);
    float* t = (float*) malloc(sizeof(float) * (size_t) n);
    for (int c = 0; c < ch; ++c) {
        float* b = buf + (size_t) c * n;
        memcpy(t, b + h1, sizeof(float) * (size_t) h2);
        memcpy(t + h2, b, sizeof(float) * (size_t) h1);
        memcpy(b, t, sizeof(float) * (size_t) n);
    }
    free(t);
}

/* ------------------------------------------------------------------------------------------------
 * deconvolve -- convolution.cpp:351-403.  Channel 0 of each input is used (tools.cpp:328).
 * out: nextPow2(max(Ln, Ld)) floats; returns that length.
 * ---------------------------------------------------------------------------------------------- */
int orc_deconvolve(const float* num, int Ln, const float* den, int Ld, double sampleRate, int smoothing,
                   int includePhase, int includeAmplitude, float* out) {
    int L = Ln > Ld ? Ln : Ld;                                                       /* :358-362 */
    int N = orc_next_pow2(L);
    int fbs = 2 * N;
    float* a = (float*) calloc((size_t) fbs, sizeof(float));
    float* b = (float*) calloc((size_t) fbs, sizeof(float));
    memcpy(a, num, sizeof(float) * (size_t) Ln);
    memcpy(b, den, sizeof(float) * (size_t) Ld);
    orc_fft fft; orc_fft_init(&fft, N);
    orc_fft_real_forward(&fft, a);                                                   /* :365-366 */
    orc_fft_real_forward(&fft, b);
    for (int i = 0; i <= N; i += 2) orc_complex_div_cartesian(a + i, a + i + 1, b[i], b[i + 1]);   /* :370-381 */
    if (smoothing) {                                                                 /* :389-394 */
        float smoothPerAvg = 1.0 / 13.0;
        for (int i = 0; i < 3; ++i) orc_averaging_filter(a, 1, fbs, smoothPerAvg, sampleRate, 1, includePhase, includeAmplitude);
    }
    orc_fft_real_inverse(&fft, a);                                                   /* :397 */
    memcpy(out, a, sizeof(float) * (size_t) N);
    if (!includePhase) orc_shifteroo(out, 1, N);                                     /* :400 */
    orc_fft_free(&fft);
    free(a); free(b);
    return N;
}

/* ir::invertFilter -- ir.cpp:13-18 = deconvolve(generatePulse(len), buffer, sr) with default flags */
int orc_invert_filter(const float* x, int L, int sampleRate, float* out) {
    float* pulse = (float*) calloc((size_t) L, sizeof(float));
    pulse[0] = 1.0f;                                                                 /* tools.cpp:235-241 */
    int n = orc_deconvolve(pulse, L, x, L, (double) sampleRate, 1, 1, 1, out);
    free(pulse);
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * ExpSineSweep -- ExpSineSweep.cpp:26-41 (generate), :59-79 (generateInv), :212-220 (assignParameters)
 * mode 0: sweep; mode 1: inverse sweep.  out may be NULL to query the length (int) T.
 * ---------------------------------------------------------------------------------------------- */
int orc_ess(double dur, double sr, double f1, double f2, double dBGain, int mode, double* out) {
    double T = sr * dur;
    double w1 = f1 / sr * 2 * M_PI, w2 = f2 / sr * 2 * M_PI;
    double K = T * w1 / log(w2 / w1);
    double L = T / log(w2 / w1);
    int n = (int) T;
    if (!out) return n;
    double g = pow(10.0, dBGain / 20.0);                                             /* tools.cpp:94-96 */
    for (int i = 0; i < n; ++i) out[i] = g * sin(K * (exp((double) i / L) - 1.0));
    if (mode == 1) {
        for (int i = 0; i < n / 2; ++i) { double t = out[i]; out[i] = out[n - 1 - i]; out[n - 1 - i] = t; }
        double k = pow(10.0, (-6.0 * log2(w2 / w1)) / 20.0 / T);
        double kit = k;
        for (int i = 0; i < n; ++i) { out[i] *= kit; kit *= k; }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * Real-time engine -- restatement of the convolution block of IRBaboonAudioProcessor::processBlock
 * (Source/PluginProcessor.cpp:403-562), its set-up (:57-82, prepareToPlay :164-234) and state
 * (PluginProcessor.h:194-217).  PluginProcessor.cpp cannot be compiled without JUCE's AudioProcessor,
 * so this part of the oracle is a restatement only; it is cross-checked against orc_convolve_periodic
 * (same arithmetic, KA2) in tests/test_oracle.py.
 *
 * Generalised in two places, both reducing to the reference when B=256 and the IR is 2048 taps:
 *   - B (processBlockSize) is a parameter (reference: const 256, PluginProcessor.h:197)
 *   - P follows the IR length given at creation (reference pins P to IRpulse, :225-226)
 * The IR is refreshed round-robin, one partition per processed block (:455-461), so a new IR set with
 * orc_rt_set_ir() fades in over P blocks exactly like the plugin's IRtoConvolve switch (:411-414).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int B, N, fbs, H, ch, P, R;            /* R = audio FDL ring size = max(P, inputArraySize) */
    int inArr, outArr;
    float *inBuf, *auFft, *irFft, *inplace, *overlap, *convRes, *outBuf;
    float* ir; int irLen;
    int inW, inSample, auW, auR, saved, irW, crW, crR, outW, outR, outBufSample, outSample;
    int blocksToProcess, blocksToOutput;
    orc_fft fft;
} orc_rt;

void* orc_rt_create(int B, int hostBlock, int channels, const float* ir, int irLen) {
    orc_rt* e = (orc_rt*) calloc(1, sizeof(orc_rt));
    e->B = B; e->H = hostBlock; e->ch = channels;
    e->N = 1; while (e->N < 2 * B - 1) e->N *= 2;                                    /* :61-65 */
    e->fbs = 2 * e->N;
    e->inArr = (int) ceilf((float) hostBlock / (float) B);                           /* :203 */
    e->outArr = orc_next_pow2((int) (ceilf((float) B / (float) hostBlock) + 1));     /* :205 */
    e->P = (int) ceilf((float) irLen / (float) B);                                   /* :225 */
    e->R = e->P > e->inArr ? e->P : e->inArr;                                        /* :229-230 */
    e->inBuf = (float*) calloc((size_t) e->inArr * channels * e->fbs, sizeof(float));
    e->convRes = (float*) calloc((size_t) e->inArr * channels * e->fbs, sizeof(float));
    e->outBuf = (float*) calloc((size_t) e->outArr * channels * hostBlock, sizeof(float));
    e->auFft = (float*) calloc((size_t) e->R * channels * e->fbs, sizeof(float));
    e->irFft = (float*) calloc((size_t) e->P * e->fbs, sizeof(float));
    e->inplace = (float*) calloc((size_t) e->fbs, sizeof(float));
    e->overlap = (float*) calloc((size_t) channels * B, sizeof(float));
    e->irLen = e->P * B;                                                             /* zero-extended so :457 never reads past the end */
    e->ir = (float*) calloc((size_t) e->irLen, sizeof(float));
    memcpy(e->ir, ir, sizeof(float) * (size_t) irLen);
    e->outR = 1;                                                                     /* :218 */
    e->auW = 0;
    e->auR = e->R - 1;                                                               /* :231-233 */
    e->saved = e->auR;
    orc_fft_init(&e->fft, e->N);
    return e;
}
void orc_rt_destroy(void* h) {
    orc_rt* e = (orc_rt*) h;
    orc_fft_free(&e->fft);
    free(e->inBuf); free(e->convRes); free(e->outBuf); free(e->auFft); free(e->irFft); free(e->inplace); free(e->overlap); free(e->ir);
    free(e);
}
/* switch IRtoConvolve (:411-414); partitions are picked up round-robin by the following blocks */
void orc_rt_set_ir(void* h, const float* ir, int irLen) {
    orc_rt* e = (orc_rt*) h;
    memset(e->ir, 0, sizeof(float) * (size_t) e->irLen);
    memcpy(e->ir, ir, sizeof(float) * (size_t) (irLen < e->irLen ? irLen : e->irLen));
}
/* one host callback: buffer is ch * n planar floats (n <= hostBlock given at creation), processed in place.
 * The output gain / limiter of :567-574 is NOT applied here (see orc_rt_post). */
void orc_rt_process(void* h, float* buffer, int n) {
    orc_rt* e = (orc_rt*) h;
    const int B = e->B, fbs = e->fbs, ch = e->ch, H = e->H, N = e->N;
    for (int s = 0; s < n; ++s) {                                                    /* :421-445 */
        for (int c = 0; c < ch; ++c) e->inBuf[((size_t) e->inW * ch + c) * fbs + e->inSample] = buffer[(size_t) c * n + s];
        e->inSample++;
        if (e->inSample >= B) {
            memcpy(e->auFft + (size_t) e->auW * ch * fbs, e->inBuf + (size_t) e->inW * ch * fbs, sizeof(float) * (size_t) ch * fbs);
            for (int c = 0; c < ch; ++c) orc_fft_real_forward(&e->fft, e->auFft + ((size_t) e->auW * ch + c) * fbs);
            e->auW = (e->auW + 1) % e->R;
            e->inW = (e->inW + 1) % e->inArr;
            memset(e->inBuf + (size_t) e->inW * ch * fbs, 0, sizeof(float) * (size_t) ch * fbs);
            e->inSample = 0;
            e->blocksToProcess++;
        }
    }
    while (e->blocksToProcess > 0) {                                                 /* :452-518 */
        float* irw = e->irFft + (size_t) e->irW * fbs;                               /* :455-461 round-robin IR refresh */
        memset(irw, 0, sizeof(float) * (size_t) fbs);
        memcpy(irw, e->ir + (size_t) e->irW * B, sizeof(float) * (size_t) B);
        orc_fft_real_forward(&e->fft, irw);
        e->irW = (e->irW + 1) % e->P;

        float* cres = e->convRes + (size_t) e->crW * ch * fbs;
        memset(cres, 0, sizeof(float) * (size_t) ch * fbs);                          /* :466 */
        e->saved = (e->saved + 1) % e->R;                                            /* :468-470 */
        for (int c = 0; c < ch; ++c) {                                               /* :472-512 */
            float* ov = e->overlap + (size_t) c * B;
            float* cr = cres + (size_t) c * fbs;
            int a = e->saved;
            for (int p = 0; p < e->P; ++p) {
                memcpy(e->inplace, e->auFft + ((size_t) a * ch + c) * fbs, sizeof(float) * (size_t) fbs);
                const float* ir = e->irFft + (size_t) p * fbs;                       /* IR channel 0 for every audio channel (:485) */
                for (int i = 0; i <= N; i += 2) orc_complex_mul(e->inplace + i, e->inplace + i + 1, ir[i], ir[i + 1]);
                for (int i = 0; i < fbs; ++i) cr[i] += e->inplace[i];
                a--; if (a < 0) a = e->R - 1;
            }
            orc_fft_real_inverse(&e->fft, cr);                                       /* :504 */
            for (int i = 0; i < B; ++i) { cr[i] += ov[i]; ov[i] = cr[B + i]; }       /* :507-510 */
        }
        e->crW = (e->crW + 1) % e->inArr;
        e->blocksToOutput++;
        e->blocksToProcess--;
    }
    while (e->blocksToOutput > 0) {                                                  /* :525-545 */
        const float* cres = e->convRes + (size_t) e->crR * ch * fbs;
        for (int s = 0; s < B; ++s) {
            if (e->outBufSample == 0) memset(e->outBuf + (size_t) e->outW * ch * H, 0, sizeof(float) * (size_t) ch * H);
            for (int c = 0; c < ch; ++c) e->outBuf[((size_t) e->outW * ch + c) * H + e->outBufSample] = cres[(size_t) c * fbs + s];
            e->outBufSample++;
            if (e->outBufSample >= H) { e->outW = (e->outW + 1) % e->outArr; e->outBufSample = 0; }
        }
        e->crR = (e->crR + 1) % e->inArr;
        e->blocksToOutput--;
    }
    for (int s = 0; s < n; ++s) {                                                    /* :552-560 */
        for (int c = 0; c < ch; ++c) buffer[(size_t) c * n + s] = e->outBuf[((size_t) e->outR * ch + c) * H + e->outSample];
        e->outSample++;
        if (e->outSample >= H) { e->outR = (e->outR + 1) % e->outArr; e->outSample = 0; }
    }
}
/* output gain and "makeshift limiter" -- PluginProcessor.cpp:567-574; buffer is ch * n planar floats */
void orc_rt_post(float* buffer, int ch, int n, float outputVolumedB) {
    float g = powf(10.0f, outputVolumedB / 20.0f);                                   /* tools.cpp:89-91 */
    for (int i = 0; i < ch * n; ++i) buffer[i] *= g;
    float mag = 0.0f;
    for (int i = 0; i < n; ++i) { float a = fabsf(buffer[i]); if (a > mag) mag = a; }   /* getMagnitude(0, 0, n): channel 0 only */
    float dB = (mag == 0.0f) ? -333.0f : 20.0f * logf(mag) / logf(10.0f);            /* tools.cpp:99-102 */
    if (dB > 0.0) {
        float m = 0.0f;
        for (int i = 0; i < ch * n; ++i) { float a = fabsf(buffer[i]); if (a > m) m = a; }
        float gain = powf(10.0f, 0.0f / 20.0f) / m;                                  /* tools.cpp:111-115 */
        for (int i = 0; i < ch * n; ++i) buffer[i] *= gain;
    }
}

/* ------------------------------------------------------------------------------------------------
 * deterministic synthetic inputs -- shared definition with irbaboon_b200/synth.py
 * ---------------------------------------------------------------------------------------------- */
static uint64_t orc_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
void orc_white_noise(uint64_t seed, uint64_t stream, int n, float* out) {
    for (int i = 0; i < n; ++i) {
        uint64_t u = orc_splitmix64(seed + (stream << 40) + (uint64_t) i);
        out[i] = (float) ((double) (u >> 40) * (1.0 / 16777216.0) * 2.0 - 1.0);
    }
}
