/*
 * oracle/ref_capi.cpp -- TEST INFRASTRUCTURE ONLY (checker + CPU baseline; never the product path).
 *
 * extern "C" wrapper that lets Python (ctypes) and bench.py drive the REFERENCE's own object code:
 * /root/reference/fp/{convolution,tools,ir,CircularBufferArray,ExpSineSweep}.cpp compiled unmodified
 * against oracle/juce_shim (see oracle/Makefile).  Output: oracle/_ref/libirb_ref.so (git-ignored).
 *
 * Buffers cross this boundary as planar, channel-contiguous float arrays: x[ch * L + i].
 */
#include <fp_include_all.hpp>

#include <atomic>
#include <chrono>
#include <thread>
#include <cstdint>

/* link stubs: ir.cpp:111 instantiates a ParallelBufferPrinter it never uses; ParallelBufferPrinter.cpp
 * (debug WAV/TSV dumper, out of scope) is not compiled. */
namespace fp {
ParallelBufferPrinter::ParallelBufferPrinter() : maxBufferLength(0) {}
ParallelBufferPrinter::~ParallelBufferPrinter() {}
ParallelBufferPrinter::PrintBuffer::PrintBuffer() : empty(true) {}
ParallelBufferPrinter::PrintBuffer::~PrintBuffer() {}
}

namespace {

AudioSampleBuffer toBuffer(const float* p, int ch, int n) {
    AudioSampleBuffer b(ch, n);
    for (int c = 0; c < ch; ++c) b.copyFrom(c, 0, p + (size_t) c * (size_t) n, n);
    return b;
}

void fromBuffer(const AudioSampleBuffer& b, float* out) {
    const int n = b.getNumSamples();
    for (int c = 0; c < b.getNumChannels(); ++c) memcpy(out + (size_t) c * (size_t) n, b.getReadPointer(c), sizeof(float) * (size_t) n);
}

/* deterministic generators shared with the product tests */
inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline float noiseSample(uint64_t seed, uint64_t stream, uint64_t i) {
    uint64_t u = splitmix64(seed + (stream << 40) + i);
    return (float) ((double) (u >> 40) * (1.0 / 16777216.0) * 2.0 - 1.0);
}

} // namespace

extern "C" {

int ref_convolve_periodic(const float* x, int chx, int Lx, const float* h, int chh, int Lh, int B, float* out) {
    AudioSampleBuffer a = toBuffer(x, chx, Lx), b = toBuffer(h, chh, Lh);
    AudioSampleBuffer r = fp::convolution::convolvePeriodic(a, b, B);
    fromBuffer(r, out);
    return r.getNumSamples();
}

/* out must hold chx * max(Lx, Lx+Lh-1) floats; returns samples per channel actually produced */
int ref_convolve_nonperiodic(const float* x, int chx, int Lx, const float* h, int chh, int Lh, float* out) {
    AudioSampleBuffer a = toBuffer(x, chx, Lx), b = toBuffer(h, chh, Lh);
    AudioSampleBuffer r = fp::convolution::convolveNonPeriodic(a, b);
    fromBuffer(r, out);
    return r.getNumSamples();
}

/* out must hold nextPow2(max(Ln, Ld)) floats; returns that length */
int ref_deconvolve(const float* num, int chn, int Ln, const float* den, int chd, int Ld, double sampleRate,
                   int smoothing, int includePhase, int includeAmplitude, float* out) {
    AudioSampleBuffer a = toBuffer(num, chn, Ln), b = toBuffer(den, chd, Ld);
    AudioSampleBuffer r = fp::convolution::deconvolve(&a, &b, sampleRate, smoothing != 0, includePhase != 0, includeAmplitude != 0);
    fromBuffer(r, out);
    return r.getNumSamples();
}

void ref_averaging_filter(float* buf, int ch, int fftSize, double octaveFraction, double sampleRate, int logAvg,
                          int includePhase, int includeAmplitude) {
    AudioSampleBuffer a = toBuffer(buf, ch, fftSize);
    fp::convolution::averagingFilter(&a, octaveFraction, sampleRate, logAvg != 0, includePhase != 0, includeAmplitude != 0);
    fromBuffer(a, buf);
}

/* out: ch * 2*nextPow2(L) floats; returns 2*N */
int ref_fft_transform(const float* x, int ch, int L, int formatAmplPhase, float* out) {
    AudioSampleBuffer a = toBuffer(x, ch, L);
    AudioSampleBuffer r = fp::tools::fftTransform(a, formatAmplPhase != 0);
    fromBuffer(r, out);
    return r.getNumSamples();
}

/* in: ch * fftSize floats; out: ch * fftSize/2 floats; returns fftSize/2 */
int ref_fft_inv_transform(const float* x, int ch, int fftSize, float* out) {
    AudioSampleBuffer a = toBuffer(x, ch, fftSize);
    AudioSampleBuffer r = fp::tools::fftInvTransform(a);
    fromBuffer(r, out);
    return r.getNumSamples();
}

void ref_shifteroo(float* buf, int ch, int n) {
    AudioSampleBuffer a = toBuffer(buf, ch, n);
    fp::ir::shifteroo(&a);
    fromBuffer(a, buf);
}

int ref_invert_filter(const float* x, int ch, int L, int sampleRate, float* out) {
    AudioSampleBuffer a = toBuffer(x, ch, L);
    AudioSampleBuffer r = fp::ir::invertFilter(a, sampleRate);
    fromBuffer(r, out);
    return r.getNumSamples();
}

int ref_ir_chop(const float* x, int ch, int L, int IRlength, float thresholdLeveldB, int consecutive, float* out) {
    AudioSampleBuffer a = toBuffer(x, ch, L);
    AudioSampleBuffer r = fp::ir::IRchop(a, IRlength, thresholdLeveldB, consecutive);
    fromBuffer(r, out);
    return r.getNumSamples();
}

/* out: (L/irPartSize + 1) * 2*irPartSize floats; returns that count.  (The reference also tries to
 * write /Users/flixor/Desktop/IR.txt, ir.cpp:137; the open fails silently here.) */
int ref_ir_to_real_fft_raw(const float* x, int L, int irPartSize, float* out) {
    AudioSampleBuffer a = toBuffer(x, 1, L);
    AudioSampleBuffer r = fp::ir::IRtoRealFFTRaw(a, irPartSize);
    fromBuffer(r, out);
    return r.getNumSamples();
}

void ref_sum_to_mono(float* buf, int ch, int n) {
    AudioSampleBuffer a = toBuffer(buf, ch, n);
    fp::tools::sumToMono(&a);
    fromBuffer(a, buf);
}

void ref_complex_mul(float* a, float* b, float c, float d) { fp::tools::complexMul(a, b, c, d); }
void ref_complex_div_cartesian(float* a, float* b, float c, float d) { fp::tools::complexDivCartesian(a, b, c, d); }
void ref_complex_div_polar(float* a, float* b, float c, float d) { fp::tools::complexDivPolar(a, b, c, d); }
float ref_bin_ampl(float* bin) { return fp::tools::binAmpl(bin); }
float ref_bin_phase(float* bin) { return fp::tools::binPhase(bin); }
void ref_round_to_zero(float* x, float thr) { fp::tools::roundToZero(x, thr); }
void ref_round_to_1e16(float* x) { fp::tools::roundTo1TenQuadrillionth(x); }
int ref_next_pow2(int x) { return fp::tools::nextPowerOfTwo(x); }
float ref_db_to_lin(float dB) { return fp::tools::dBToLin(dB); }
float ref_lin_to_db(float lin) { return fp::tools::linTodB(lin); }

void ref_generate_pulse(int n, int offset, float* out) {
    AudioSampleBuffer r = fp::tools::generatePulse(n, offset);
    fromBuffer(r, out);
}
void ref_linear_fade(float* buf, int ch, int n, int fadeIn, int start, int count) {
    AudioSampleBuffer a = toBuffer(buf, ch, n);
    fp::tools::linearFade(&a, fadeIn != 0, start, count);
    fromBuffer(a, buf);
}
void ref_normalize(float* buf, int ch, int n, float dBGoal) {
    AudioSampleBuffer a = toBuffer(buf, ch, n);
    fp::tools::normalize(&a, dBGoal, false);
    fromBuffer(a, buf);
}
void ref_sine_fill(float* buf, int ch, int n, float freq, float sr, float ampl) {
    AudioSampleBuffer a(ch, n);
    fp::tools::sineFill(&a, freq, sr, ampl);
    fromBuffer(a, buf);
}

/* ExpSineSweep: mode 0 = sweep, 1 = inverse sweep; fadeKind 0 none, 1 lin, 2 dB, 3 brickwall (applied to the sweep
 * before the inverse is derived, as a caller would).  out may be NULL to query the length. */
int ref_ess(double dur, double sr, double f1, double f2, double dBGain, int mode, int fadeKind, double fadeFreq, double* out) {
    fp::ExpSineSweep s;
    s.generate(dur, sr, f1, f2, dBGain);
    if (fadeKind == 1) s.linFadeout(fadeFreq);
    if (fadeKind == 2) s.dBFadeout(fadeFreq);
    if (fadeKind == 3) s.brickwallFadeout(fadeFreq);
    if (mode == 1) s.generateInv();
    AudioBuffer<double> b = mode == 1 ? s.getSweepInv() : s.getSweep();
    if (out) memcpy(out, b.getReadPointer(0), sizeof(double) * (size_t) b.getNumSamples());
    return b.getNumSamples();
}
int ref_ess_float(double dur, double sr, double f1, double f2, double dBGain, int mode, float* out) {
    fp::ExpSineSweep s;
    s.generate(dur, sr, f1, f2, dBGain);
    if (mode == 1) s.generateInv();
    AudioSampleBuffer b = mode == 1 ? s.getSweepInvFloat() : s.getSweepFloat();
    if (out) fromBuffer(b, out);
    return b.getNumSamples();
}
int ref_ess_index_at_freq(double freq, double dur, double sr, double f1, double f2) {
    fp::ExpSineSweep s;
    return s.getSampleIndexAtFreq(freq, dur, sr, f1, f2);
}
double ref_ess_freq_at_index(int idx, double dur, double sr, double f1, double f2) {
    fp::ExpSineSweep s;
    return s.getFreqAtSampleIndex(idx, dur, sr, f1, f2);
}

/* ---- CircularBufferArray, handle based ---- */
void* ref_cba_create(int buffers, int ch, int n) { return new fp::CircularBufferArray(buffers, ch, n); }
void ref_cba_destroy(void* h) { delete (fp::CircularBufferArray*) h; }
void ref_cba_clear_and_resize(void* h, int buffers, int ch, int n) { ((fp::CircularBufferArray*) h)->clearAndResize(buffers, ch, n); }
void ref_cba_change_array_size(void* h, int buffers) { ((fp::CircularBufferArray*) h)->changeArraySize(buffers); }
void ref_cba_write(void* h, const float* data) {           /* fills the write buffer (records lastWrittenIndex) */
    AudioSampleBuffer* b = ((fp::CircularBufferArray*) h)->getWriteBufferPtr();
    for (int c = 0; c < b->getNumChannels(); ++c) b->copyFrom(c, 0, data + (size_t) c * (size_t) b->getNumSamples(), b->getNumSamples());
}
void ref_cba_read(void* h, float* out) { fromBuffer(*((fp::CircularBufferArray*) h)->getReadBufferPtr(), out); }
void ref_cba_read_at(void* h, int idx, float* out) { fromBuffer(*((fp::CircularBufferArray*) h)->getBufferPtrAtIndex(idx), out); }
void ref_cba_incr_read(void* h) { ((fp::CircularBufferArray*) h)->incrReadIndex(); }
void ref_cba_decr_read(void* h) { ((fp::CircularBufferArray*) h)->decrReadIndex(); }
void ref_cba_incr_write(void* h) { ((fp::CircularBufferArray*) h)->incrWriteIndex(); }
int ref_cba_get_read_index(void* h) { return ((fp::CircularBufferArray*) h)->getReadIndex(); }
int ref_cba_get_write_index(void* h) { return ((fp::CircularBufferArray*) h)->getWriteIndex(); }
void ref_cba_set_read_index(void* h, int i) { ((fp::CircularBufferArray*) h)->setReadIndex(i); }
void ref_cba_set_write_index(void* h, int i) { ((fp::CircularBufferArray*) h)->setWriteIndex(i); }
int ref_cba_get_array_size(void* h) { return ((fp::CircularBufferArray*) h)->getArraySize(); }
int ref_cba_consolidate(void* h, int offset, float* out) {
    AudioSampleBuffer r = ((fp::CircularBufferArray*) h)->consolidate(offset);
    fromBuffer(r, out);
    return r.getNumSamples();
}

/* ---- CPU baseline: the reference's convolvePeriodic, one whole stream per call, `threads` workers ----
 * Streams are mono white noise (seed, stream id) of Lx samples; all share the IR `h` (Lh taps).
 * Returns wall seconds; *checksum receives the sum of all output samples (keeps the work observable). */
double ref_bench_convolve_periodic(int threads, int streams, int Lx, const float* h, int Lh, int B, uint64_t seed, double* checksum) {
    AudioSampleBuffer ir = toBuffer(h, 1, Lh);
    std::vector<AudioSampleBuffer> inputs;
    inputs.reserve((size_t) streams);
    for (int s = 0; s < streams; ++s) {
        AudioSampleBuffer x(1, Lx);
        float* p = x.getWritePointer(0);
        for (int i = 0; i < Lx; ++i) p[i] = noiseSample(seed, (uint64_t) s, (uint64_t) i);
        inputs.push_back(x);
    }
    std::vector<double> sums((size_t) streams, 0.0);
    std::atomic<int> next(0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&]() {
            AudioSampleBuffer myIr(ir);
            for (;;) {
                int s = next.fetch_add(1);
                if (s >= streams) break;
                AudioSampleBuffer r = fp::convolution::convolvePeriodic(inputs[(size_t) s], myIr, B);
                const float* p = r.getReadPointer(0);
                double acc = 0.0;
                for (int i = 0; i < r.getNumSamples(); ++i) acc += p[i];
                sums[(size_t) s] = acc;
            }
        });
    }
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    double total = 0.0;
    for (double v : sums) total += v;
    if (checksum) *checksum = total;
    return std::chrono::duration<double>(t1 - t0).count();
}

/* CPU baseline for the ESS path: `count` deconvolve(capture, sweep, sr, smoothing) calls over `threads` workers */
double ref_bench_deconvolve(int threads, int count, const float* captures, int L, const float* sweep, double sr, int smoothing, double* checksum) {
    AudioSampleBuffer den = toBuffer(sweep, 1, L);
    std::vector<double> sums((size_t) count, 0.0);
    std::atomic<int> next(0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&]() {
            AudioSampleBuffer myDen(den);
            for (;;) {
                int j = next.fetch_add(1);
                if (j >= count) break;
                AudioSampleBuffer num = toBuffer(captures + (size_t) j * (size_t) L, 1, L);
                AudioSampleBuffer r = fp::convolution::deconvolve(&num, &myDen, sr, smoothing != 0);
                const float* p = r.getReadPointer(0);
                double acc = 0.0;
                for (int i = 0; i < r.getNumSamples(); ++i) acc += p[i];
                sums[(size_t) j] = acc;
            }
        });
    }
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    double total = 0.0;
    for (double v : sums) total += v;
    if (checksum) *checksum = total;
    return std::chrono::duration<double>(t1 - t0).count();
}

int ref_hardware_threads() { return (int) std::thread::hardware_concurrency(); }

} // extern "C"
